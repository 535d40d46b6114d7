from wsl4mis_b200.utils.losses import *  # noqa: F401,F403
