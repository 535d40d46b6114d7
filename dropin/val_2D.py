from wsl4mis_b200.val_2D import *  # noqa: F401,F403
from wsl4mis_b200.val_2D import test_single_volume, test_single_volume_cct, test_single_volume_ds, calculate_metric_percase  # noqa: F401
