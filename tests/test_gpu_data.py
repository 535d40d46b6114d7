"""GPU input pipeline: wsl_augment_batch through the C ABI vs the reference fixture and the oracle."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import augment_oracle as A  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "augment.npz"))


def _store(ims, lbs):
    from wsl4mis_b200.dataloaders import SliceStore
    return SliceStore.from_arrays(ims, lbs)


def test_batch_equals_reference_fixture():
    """Same seeds as oracle/make_golden.py:augment_golden -> the whole batch is bit-identical to the reference's samples."""
    from wsl4mis_b200.dataloaders import RandomGenerator
    n = int(GOLD["n"])
    ims, lbs = A.synth_slices(n, int(GOLD["seed"]))
    st = _store(ims, lbs)
    params = []
    for i in range(n):
        random.seed(1000 + i)
        np.random.seed(1000 + i)
        params.append(RandomGenerator.draw(bool(st.has4[i])))
    img, lab = st.augment(list(range(n)), params, tuple(int(v) for v in GOLD["out_hw"]))
    assert torch.equal(img[:, 0].cpu(), torch.from_numpy(GOLD["image"]))
    assert torch.equal(lab.cpu(), torch.from_numpy(GOLD["label"]))


@pytest.mark.parametrize("out_hw", [(256, 256), (64, 80), (33, 47)])
def test_every_decision_against_the_oracle(out_hw):
    """all rot90/flip combinations and every angle the reference can draw, on ragged slices, image and label bit-exact"""
    ims, lbs = A.synth_slices(12, 99)
    st = _store(ims, lbs)
    decisions = [("none",)] + [("rot_flip", k, ax) for k in range(4) for ax in range(2)]
    decisions += [("rotate", ang, cv) for ang in range(-20, 20) for cv in ((4, 0) if ang % 5 == 0 else (4,))]
    idx, params, want_i, want_l = [], [], [], []
    for j, d in enumerate(decisions):
        i = j % len(ims)
        if d[0] == "rotate":
            d = ("rotate", d[1], 4 if st.has4[i] else 0) if d[2] == 4 else d
        idx.append(i)
        params.append(A.to_params(d))
        oi, ol = A.apply(ims[i], lbs[i], d, out_hw)
        want_i.append(oi)
        want_l.append(ol)
    img, lab = st.augment(idx, params, out_hw)
    got_i, got_l = img[:, 0].cpu().numpy(), lab.cpu().numpy()
    for j, d in enumerate(decisions):
        assert np.array_equal(got_i[j], want_i[j]), ("image", d, idx[j], ims[idx[j]].shape)
        assert np.array_equal(got_l[j], want_l[j]), ("label", d, idx[j])


def test_loader_order_and_shapes():
    """GpuLoader visits the slices in torch's RandomSampler order and yields the script's sample keys / dtypes."""
    from torch.utils.data import RandomSampler
    from wsl4mis_b200.dataloaders import GpuLoader
    ims, lbs = A.synth_slices(21, 5)
    st = _store(ims, lbs)
    torch.manual_seed(123)
    want = list(RandomSampler(range(21)))
    torch.manual_seed(123)
    random.seed(0)
    np.random.seed(0)
    batches = list(GpuLoader(st, 8, (64, 64), shuffle=True))
    assert [b["image"].shape[0] for b in batches] == [8, 8, 5]
    assert sum((b["idx"] for b in batches), []) == [st.names[i] for i in want]
    b = batches[0]
    assert b["image"].dtype == torch.float32 and b["image"].shape[1:] == (1, 64, 64) and b["image"].is_cuda
    assert b["label"].dtype == torch.uint8 and b["label"].shape[1:] == (64, 64)


def test_single_sample_call_has_the_reference_signature():
    from wsl4mis_b200.dataloaders import RandomGenerator
    ims, lbs = A.synth_slices(3, 1)
    random.seed(4)
    np.random.seed(4)
    d = A.draw(lbs[1])
    oi, ol = A.apply(ims[1], lbs[1], d, (64, 64))
    random.seed(4)
    np.random.seed(4)
    s = RandomGenerator((64, 64))({"image": ims[1], "label": lbs[1]})
    assert s["image"].shape == (1, 64, 64) and s["label"].dtype == torch.uint8
    assert np.array_equal(s["image"][0].numpy(), oi) and np.array_equal(s["label"].numpy(), ol)
