"""Scalar schedules used by the semi-/weakly-supervised step bodies (host side, no tensors involved).

Same names and values as the reference's ``utils/ramps.py`` (:19-41); used by ``engine.UAMTStep`` for the consistency
weight ``consistency * sigmoid_rampup(iter // 300, 200)`` and the uncertainty threshold
``(0.75 + 0.25 * sigmoid_rampup(iter, max_iter)) * ln 2`` (train_uncertainty_aware_mean_teacher_2D.py:73-75,185-186).
"""
from math import cos, exp, pi

__all__ = ["sigmoid_rampup", "linear_rampup", "cosine_rampdown"]


def _progress(step, length):
    """fraction of the ramp completed, clipped to [0, 1]"""
    return min(max(float(step), 0.0), float(length)) / float(length)


def sigmoid_rampup(current, rampup_length):
    """Gaussian-shaped warm-up exp(-5 (1 - t)^2) with t the clipped progress; 1 when there is no ramp."""
    return 1.0 if rampup_length == 0 else float(exp(-5.0 * (1.0 - _progress(current, rampup_length)) ** 2))


def linear_rampup(current, rampup_length):
    """Straight line from 0 to 1 over `rampup_length` steps, flat afterwards."""
    if current < 0 or rampup_length < 0:
        raise AssertionError("ramp arguments must be non-negative")
    return 1.0 if current >= rampup_length else current / rampup_length


def cosine_rampdown(current, rampdown_length):
    """Half cosine from 1 down to 0; only defined inside the ramp."""
    if not 0 <= current <= rampdown_length:
        raise AssertionError("cosine_rampdown is defined for 0 <= current <= rampdown_length")
    return float(0.5 * (1.0 + cos(pi * current / rampdown_length)))
