from .deepfm import *  # noqa: F401,F403
