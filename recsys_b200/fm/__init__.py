from .fm import *  # noqa: F401,F403
