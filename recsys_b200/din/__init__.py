from .din import *  # noqa: F401,F403
