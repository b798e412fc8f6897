from .xdeepfm import *  # noqa: F401,F403
