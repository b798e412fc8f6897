from .dcn import *  # noqa: F401,F403
