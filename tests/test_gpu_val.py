"""Validation path: GPU nearest-neighbour zoom == scipy.ndimage.zoom(order=0) bit for bit, overlap
counts == numpy, and test_single_volume == the reference's per-slice loop restated with scipy."""
import numpy as np
import pytest
import torch
from scipy.ndimage import zoom

pytestmark = pytest.mark.gpu
DEV = "cuda"

if torch.cuda.is_available():
    from wsl4mis_b200 import val_2D
    from wsl4mis_b200._lib import call
    from wsl4mis_b200.networks.unet import UNet, UNet_CCT


@pytest.mark.parametrize("hw", [(208, 174), (256, 216), (154, 231), (300, 280), (256, 256), (17, 40)])
def test_zoom_matches_scipy(hw):
    h, w = hw
    rs = np.random.RandomState(1)
    a = rs.rand(3, h, w).astype(np.float32)
    out = torch.empty((3, 256, 256), device=DEV)
    call("wsl_zoom_nearest", torch.from_numpy(a).to(DEV), 0, 3, h, w, 256, 256, out)
    ref = np.stack([zoom(a[i], (256 / h, 256 / w), order=0) for i in range(3)])
    assert np.array_equal(out.cpu().numpy(), ref)
    lab = rs.randint(0, 4, size=(3, 256, 256)).astype(np.uint8)
    back = torch.empty((3, h, w), dtype=torch.uint8, device=DEV)
    call("wsl_zoom_nearest", torch.from_numpy(lab).to(DEV), 1, 3, 256, 256, h, w, back)
    refb = np.stack([zoom(lab[i], (h / 256, w / 256), order=0) for i in range(3)])
    assert np.array_equal(back.cpu().numpy(), refb)


def test_overlap_counts_and_hd95():
    rs = np.random.RandomState(2)
    p = rs.randint(0, 4, size=(5, 64, 48)).astype(np.uint8)
    g = rs.randint(0, 4, size=(5, 64, 48)).astype(np.uint8)
    counts = torch.zeros(12, dtype=torch.int64, device=DEV)
    call("wsl_overlap_counts", torch.from_numpy(p).to(DEV), torch.from_numpy(g).to(DEV), p.size, 4, counts)
    c = counts.cpu().numpy().reshape(4, 3)
    for i in range(1, 4):
        assert c[i, 0] == np.count_nonzero((p == i) & (g == i)) and c[i, 1] == np.count_nonzero(p == i) and c[i, 2] == np.count_nonzero(g == i)
    # hd95 of two shifted cubes: every surface voxel is exactly 2 away in x -> percentile is 2
    a = np.zeros((8, 32, 32), bool); b = np.zeros((8, 32, 32), bool)
    a[2:6, 8:20, 8:20] = True; b[2:6, 8:20, 10:22] = True
    assert abs(val_2D.hd95(a, b) - 2.0) < 1e-6
    assert val_2D.calculate_metric_percase(np.zeros((4, 4)), np.ones((4, 4))) == (0, 0)


@pytest.mark.parametrize("cct", [False, True])
def test_single_volume_matches_per_slice_loop(cct):
    """batch-of-slices GPU path vs the reference's loop (val_2D.py:21-36) run with the same network in fp32 parity mode."""
    torch.manual_seed(4)
    net = (UNet_CCT if cct else UNet)(1, 4).to(DEV).set_precision("fp32")
    if cct:
        net.channel_keep = None
    rs = np.random.RandomState(3)
    S, h, w = 5, 40, 56
    image = torch.from_numpy(rs.rand(1, S, h, w).astype(np.float32))
    label = torch.from_numpy(rs.randint(0, 4, size=(1, S, h, w)).astype(np.uint8))
    fn = val_2D.test_single_volume_cct if cct else val_2D.test_single_volume
    got = fn(image, label, net, 4, patch_size=[64, 64])
    # the reference's loop, slice by slice
    pred = np.zeros((S, h, w), np.uint8)
    net.eval()
    for ind in range(S):
        sl = zoom(image[0, ind].numpy(), (64 / h, 64 / w), order=0)
        with torch.no_grad():
            out = net(torch.from_numpy(sl)[None, None].float().to(DEV))
        out = out[0] if cct else out
        o = torch.argmax(torch.softmax(out, 1), 1).squeeze(0).cpu().numpy()
        pred[ind] = zoom(o, (h / 64, w / 64), order=0)
    for i in range(1, 4):
        ref = val_2D.calculate_metric_percase(pred == i, label[0].numpy() == i)
        assert abs(got[i - 1][0] - ref[0]) < 1e-12 and abs(got[i - 1][1] - ref[1]) < 1e-9, (i, got[i - 1], ref)
