from wsl4mis_b200.utils.gate_crf_loss import *  # noqa: F401,F403
