from wsl4mis_b200.networks.pnet import *  # noqa: F401,F403
from wsl4mis_b200.networks.pnet import PNetBlock, ConcatBlock, OutPutBlock, PNet2D  # noqa: F401
