"""API-surface names of utils/losses.py that are off the hot path (FocalLoss, SizeLoss, SupConLoss, softmax_kl_loss, dice_loss1,
softmax_dice_loss: reference utils/losses.py:19-27,39-55,85-104,119-153,248-272,311-398) against values generated from the
unmodified reference by oracle/make_golden.py (tests/golden/losses_api.npz).  Plain tensor expressions -> runs on CPU."""
import os
import warnings

import numpy as np
import torch

from wsl4mis_b200.utils import losses as L


def test_api_surface_losses_match_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "losses_api.npz"))
    a, b, tgt = torch.from_numpy(g["a"]), torch.from_numpy(g["b"]), torch.from_numpy(g["target"])
    sa, sb = torch.softmax(a, 1), torch.softmax(b, 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = {
            "dice_loss1": L.dice_loss1(sa[:, 1], sb[:, 1]),
            "softmax_dice_loss": L.softmax_dice_loss(a, b),
            "softmax_kl_loss": L.softmax_kl_loss(a, b),
            "softmax_kl_loss_sigmoid": L.softmax_kl_loss(a, b, sigmoid=True),
            "focal": L.FocalLoss(gamma=2)(a, tgt),
            "focal_alpha_sum": L.FocalLoss(gamma=1.5, alpha=[0.1, 0.2, 0.3, 0.4], size_average=False)(a, tgt),
            "size_loss": L.SizeLoss(0.1)(torch.from_numpy(g["vol"]), torch.from_numpy(g["vol_target"])),
            "supcon_labels": L.SupConLoss()(torch.from_numpy(g["feats"]), torch.from_numpy(g["feat_labels"])),
            "supcon_simclr": L.SupConLoss()(torch.from_numpy(g["feats"])),
            "supcon_one": L.SupConLoss(contrast_mode="one")(torch.from_numpy(g["feats"]), torch.from_numpy(g["feat_labels"])),
        }
    for k, v in got.items():
        assert abs(float(v) - float(g[k])) <= 2e-6 * max(1.0, abs(float(g[k]))), (k, float(v), float(g[k]))


def test_every_public_name_of_the_reference_module_exists():
    """names a `from utils import losses` user can reach in the reference (losses.py, all top-level defs / classes)"""
    for name in ("dice_loss", "dice_loss1", "entropy_loss", "softmax_dice_loss", "entropy_loss_map", "softmax_mse_loss",
                 "softmax_kl_loss", "symmetric_mse_loss", "FocalLoss", "DiceLoss", "pDLoss", "entropy_minmization", "entropy_map",
                 "SizeLoss", "MumfordShah_Loss", "SupConLoss"):
        assert hasattr(L, name), name


def test_gatedcrf_general_arguments_match_the_reference(golden_dir):
    """Argument patterns no WSL4MIS script uses (several descriptors, radius 3, a modality above the prediction resolution,
    masks, kernel visualisation): the tensor-expression path of ModelLossSemsegGatedCRF against reference-generated values."""
    from wsl4mis_b200.utils.gate_crf_loss import ModelLossSemsegGatedCRF
    g = np.load(os.path.join(golden_dir, "crf_general.npz"))
    y = torch.from_numpy(g["y"]).requires_grad_(True)
    sample, ms, md = torch.from_numpy(g["sample"]), torch.from_numpy(g["mask_src"]), torch.from_numpy(g["mask_dst"])
    desc = [{"weight": 0.9, "xy": 6, "rgb": 0.1}, {"weight": 0.1, "xy": 4}]
    m = ModelLossSemsegGatedCRF()
    r = m(y, desc, 3, sample, 24, 40)
    (gy,) = torch.autograd.grad(r["loss"], y)
    assert abs(r["loss"].item() - float(g["plain:loss"])) < 2e-5 * abs(float(g["plain:loss"]))
    assert np.allclose(gy.numpy(), g["plain:grad"], rtol=1e-4, atol=1e-7)
    r = m(y, desc, 3, sample, 24, 40, mask_src=ms, mask_dst=md, out_kernels_vis=True)
    (gy,) = torch.autograd.grad(r["loss"], y)
    assert abs(r["loss"].item() - float(g["masked:loss"])) < 2e-5 * abs(float(g["masked:loss"]))
    assert np.allclose(gy.numpy(), g["masked:grad"], rtol=1e-4, atol=1e-7)
    assert np.allclose(r["kernels_vis"].detach().numpy(), g["masked:vis"], rtol=1e-5, atol=1e-7)
    r = m(y, [{"weight": 1, "xy": 6, "rgb": 0.1}], 2, sample[:, :1, ::2, ::2].contiguous(), 12, 20)
    assert abs(r["loss"].item() - float(g["r2:loss"])) < 2e-5 * abs(float(g["r2:loss"]))
    try:
        m(y, desc, 3, sample, 24, 40, compatibility=torch.ones(3, 3) - torch.eye(3))
        err = "none"
    except Exception as e:
        err = type(e).__name__
    assert err == str(g["compat:error"])
