from wsl4mis_b200.dataloaders.dataset import *  # noqa: F401,F403
from wsl4mis_b200.dataloaders.dataset import BaseDataSets, RandomGenerator  # noqa: F401
