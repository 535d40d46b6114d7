"""dc / hd95 with medpy's definitions (Dice = 2|A&B| / (|A|+|B|); HD95 = 95th percentile of the symmetric surface distances,
connectivity 1, unit spacing), written against numpy / scipy for the test harness."""
import numpy as np
from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure


def dc(result, reference):
    a, b = np.atleast_1d(np.asarray(result).astype(bool)), np.atleast_1d(np.asarray(reference).astype(bool))
    s = np.count_nonzero(a) + np.count_nonzero(b)
    return 2.0 * np.count_nonzero(a & b) / float(s) if s else 0.0


def _surf(a, b):
    a, b = np.atleast_1d(np.asarray(a).astype(bool)), np.atleast_1d(np.asarray(b).astype(bool))
    if not a.any() or not b.any():
        raise RuntimeError("empty binary object")
    fp = generate_binary_structure(a.ndim, 1)
    ea, eb = a ^ binary_erosion(a, structure=fp, iterations=1), b ^ binary_erosion(b, structure=fp, iterations=1)
    return distance_transform_edt(~eb)[ea]


def hd95(result, reference, voxelspacing=None, connectivity=1):
    return float(np.percentile(np.hstack((_surf(result, reference), _surf(reference, result))), 95))
