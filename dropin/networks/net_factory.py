from wsl4mis_b200.networks.net_factory import *  # noqa: F401,F403
