from wsl4mis_b200.utils.ramps import *  # noqa: F401,F403
