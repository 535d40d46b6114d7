from wsl4mis_b200.networks.net_factory import *  # noqa: F401,F403
from wsl4mis_b200.networks.net_factory import net_factory  # noqa: F401
