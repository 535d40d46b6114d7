"""CPU: the augmentation oracle against the reference-generated fixture, and the product's draw order against the oracle's."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import augment_oracle as A  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "augment.npz"))


def test_augment_oracle_matches_reference_fixture():
    """dataset_semi.py:146-171 run unmodified (oracle/make_golden.py:augment_golden) vs the restatement, bit for bit."""
    ims, lbs = A.synth_slices(int(GOLD["n"]), int(GOLD["seed"]))
    seen = set()
    for i, (im, lb) in enumerate(zip(ims, lbs)):
        random.seed(1000 + i)
        np.random.seed(1000 + i)
        d = A.draw(lb)
        seen.add(d[0])
        oi, ol = A.apply(im, lb, d, tuple(GOLD["out_hw"]))
        assert np.array_equal(oi, GOLD["image"][i]) and np.array_equal(ol, GOLD["label"][i]), (i, d)
    assert seen == {"rot_flip", "rotate", "none"}


def test_product_draw_consumes_the_rng_like_the_reference():
    from wsl4mis_b200.dataloaders.dataset import RandomGenerator
    ims, lbs = A.synth_slices(64, 7)
    for i, lb in enumerate(lbs):
        random.seed(i)
        np.random.seed(i)
        want = A.to_params(A.draw(lb))
        tail_want = (random.random(), int(np.random.randint(0, 1 << 30)))
        random.seed(i)
        np.random.seed(i)
        got = RandomGenerator.draw(bool((lb == 4).any()))
        tail_got = (random.random(), int(np.random.randint(0, 1 << 30)))
        assert got == want and tail_got == tail_want, (i, got, want)


def test_rotation_matrix_is_scipys():
    """The host-side matrix/offset must reproduce scipy.ndimage.rotate's own float64 values (checked through affine_transform)."""
    from scipy import ndimage
    from wsl4mis_b200.dataloaders.dataset import _rotation
    rs = np.random.RandomState(3)
    for angle in (-20, -7, 0, 1, 13, 19):
        x = rs.rand(37, 52).astype(np.float32)
        rot, off = _rotation(angle, x.shape)
        a = ndimage.affine_transform(x, rot, off, x.shape, order=0)
        b = ndimage.rotate(x, angle, order=0, reshape=False)
        assert np.array_equal(a, b), angle


def test_loader_shards_like_distributed_sampler():
    """rank::world_size shares of one permutation: disjoint, equal length, together cover every slice (padding by wrapping)."""
    import torch
    from torch.utils.data.distributed import DistributedSampler
    from wsl4mis_b200.dataloaders.dataset import GpuLoader
    order = torch.randperm(21, generator=torch.Generator().manual_seed(3)).tolist()
    shares = [GpuLoader.shard(order, r, 4) for r in range(4)]
    assert all(len(s) == 6 for s in shares)
    assert set(sum(shares, [])) == set(range(21))
    ds = list(range(21))
    for r in range(4):     # same partition rule as torch's sampler applied to the identity permutation
        want = list(DistributedSampler(ds, num_replicas=4, rank=r, shuffle=False))
        assert GpuLoader.shard(ds, r, 4) == want


def test_sample_table_layout_matches_the_kernel_struct():
    """The numpy record the host fills per sample must be byte-compatible with AugSample in csrc/data_ops.cu."""
    from wsl4mis_b200._lib import LIB
    from wsl4mis_b200.dataloaders.dataset import _SAMPLE
    assert LIB.load().wsl_augment_sample_bytes() == _SAMPLE.itemsize == 80
    assert _SAMPLE.fields["off"][1] == 0 and _SAMPLE.fields["h"][1] == 8 and _SAMPLE.fields["m00"][1] == 32
