"""Pin the CPU oracle (oracle/wsl_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import wsl_oracle as O

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _masks(seed, n, h, w):
    rs = np.random.RandomState(seed)
    names = ["encoder.in_conv.conv_conv.3"] + [f"encoder.down{i}.maxpool_conv.1.conv_conv.3" for i in range(1, 5)]
    return {nm: torch.from_numpy((rs.uniform(size=(n, O.FT[i], h >> i, w >> i)) >= O.ENC_DROP[i]).astype(np.uint8))
            for i, nm in enumerate(names)}


def _chan(seed, n):
    rs = np.random.RandomState(seed)
    return [torch.from_numpy((rs.uniform(size=(n, c)) >= 0.5).astype(np.uint8)) for c in O.FT]


def test_losses_match_reference(golden_dir):
    g = _load(golden_dir, "losses_kat.npz")
    logits = torch.from_numpy(g["logits"]).requires_grad_(True)
    logits2 = torch.from_numpy(g["logits2"])
    img = torch.from_numpy(g["image"])
    lab = torch.from_numpy(g["label"])
    s = torch.softmax(logits, 1)
    pseudo = torch.argmax(s.detach(), 1, keepdim=True)
    vals = {
        "pce": O.pce_loss(logits, lab),
        "gatedcrf": O.gated_crf_loss(s, img),
        "mumford_shah": O.mumford_shah_loss(img, s),
        "pdice_argmax": O.pdice_loss(s, pseudo),
        "pdice_ignore": O.pdice_loss(s, lab.long().unsqueeze(1)),
        "dice": O.dice_loss(s, pseudo),
        "entropy": O.entropy_minimization(s),
        "tv": O.tv_loss(s),
        "step_pce_gatedcrf": O.step_loss_pce_gatedcrf(logits, img, lab)[0],
    }
    for k, v in vals.items():
        ref = float(g["loss:" + k])
        assert abs(v.item() - ref) <= 2e-5 * max(1.0, abs(ref)), (k, v.item(), ref)
        (gr,) = torch.autograd.grad(v, logits, retain_graph=True)
        gref = torch.from_numpy(g["grad:" + k])
        scale = gref.abs().max().item() + 1e-12
        assert (gr - gref).abs().max().item() <= 2e-4 * scale, k
    mse = O.softmax_mse(logits, logits2)
    assert np.allclose(mse.detach().numpy(), g["softmax_mse"], atol=1e-6)
    s2 = torch.softmax(logits2, 1)
    mix = O.mix_pseudo_label(s, s2, float(g["beta"]))
    assert np.array_equal(mix.numpy().astype(np.uint8), g["pseudo_mix"])


def test_gatedcrf_border_case(golden_dir):
    g = _load(golden_dir, "losses_kat.npz")
    y = torch.from_numpy(g["crf2:y"]).requires_grad_(True)
    img = torch.from_numpy(g["crf2:image"])
    loss = O.gated_crf_loss(y, img)
    assert abs(loss.item() - float(g["crf2:loss"])) < 2e-5 * abs(float(g["crf2:loss"]))
    (gy,) = torch.autograd.grad(loss, y)
    assert np.allclose(gy.numpy(), g["crf2:grad_y"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("cct", [False, True])
def test_network_matches_reference(golden_dir, cct):
    g = _load(golden_dir, "unet_cct_dmpls.npz" if cct else "unet_pce_gatedcrf.npz")
    n, hw = int(g["n"]), int(g["hw"])
    decs = ("main_decoder", "aux_decoder1") if cct else ("decoder",)
    p = O.synth_params(1, 4, decs, int(g["pseed"]))
    image, label = torch.from_numpy(g["image"]), torch.from_numpy(g["label"])
    im2, lb2 = O.synth_batch(n, hw, hw, seed=2022, frac=0.06)
    assert torch.equal(im2, image) and torch.equal(lb2, label)
    masks = _masks(int(g["mseed"]), n, hw, hw)
    chan = _chan(int(g["cseed"]), n) if cct else None
    with torch.no_grad():
        if cct:
            em, ea = O.unet_cct_forward(p, image, False, None, chan)
            assert np.allclose(ea.numpy(), g["eval_aux"], atol=2e-5)
        else:
            em = O.unet_forward(p, image, False)
    assert np.allclose(em.numpy(), g["eval_main"], atol=2e-5)
    variant = "dmpls" if cct else "pce_gatedcrf"
    beta = float(g["beta"]) if cct else 0.5
    loss, grads, (main, aux) = O.full_step(p, image, label, variant, cct, masks, chan, beta)
    assert np.allclose(main.numpy(), g["train_main"], atol=5e-5)
    if cct:
        assert np.allclose(aux.numpy(), g["train_aux"], atol=5e-5)
        s1, s2 = torch.softmax(main, 1), torch.softmax(aux, 1)
        assert np.array_equal(O.mix_pseudo_label(s1, s2, beta).numpy().astype(np.uint8), g["pseudo"])
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    keys = [str(k) for k in g["grad_keys"]]
    stats = g["grad_stats"]
    for k, (s_sum, s_abs, s_l2) in zip(keys, stats):
        gr = grads[k].double()
        assert abs(gr.norm().item() - s_l2) <= 2e-3 * s_l2 + 1e-5, k  # conv biases before BN have ~1e-7 noise grads
        assert abs(gr.abs().sum().item() - s_abs) <= 2e-3 * s_abs + 1e-4, k
        if ("g:" + k) in g.files:
            ref = g["g:" + k]
            assert np.allclose(grads[k].numpy(), ref, rtol=2e-3, atol=max(2e-3 * np.abs(ref).max(), 2e-6)), k


def test_running_stats_update(golden_dir):
    g = _load(golden_dir, "unet_pce_gatedcrf.npz")
    n, hw = int(g["n"]), int(g["hw"])
    p = O.synth_params(1, 4, ("decoder",), int(g["pseed"]))
    image = torch.from_numpy(g["image"])
    new = {}
    with torch.no_grad():
        O.unet_forward(p, image, True, _masks(int(g["mseed"]), n, hw, hw), new)
    for k in [f for f in g.files if f.startswith("stat:")]:
        assert np.allclose(new[k[5:]].numpy(), g[k], rtol=1e-4, atol=1e-5), k


def test_sgd_matches_torch():
    torch.manual_seed(0)
    w = torch.randn(37)
    ref = torch.nn.Parameter(w.clone())
    opt = torch.optim.SGD([ref], lr=0.01, momentum=0.9, weight_decay=1e-4)
    params, moms = {"w": w.clone()}, {}
    for it in range(3):
        gr = torch.randn(37)
        ref.grad = gr.clone()
        opt.step()
        O.sgd_step(params, {"w": gr}, moms, 0.01)
        assert torch.allclose(params["w"], ref.data, atol=1e-7)


@pytest.mark.parametrize("which", ["ds", "3h"])
def test_other_heads_match_reference(golden_dir, which):
    """UNet_DS (Decoder_DS, four outputs) and UNet_CCT_3H (three outputs, aux_decoder1 run twice as written) against the
    fixture generated from the unmodified reference classes (restated ahead of the product)."""
    from torch.distributions.uniform import Uniform
    g = _load(golden_dir, "unet_heads.npz")
    n, hw = int(g["n"]), int(g["hw"])
    image, label = torch.from_numpy(g["image"]), torch.from_numpy(g["label"])
    masks = _masks(int(g["mseed"]), n, hw, hw)
    if which == "ds":
        p = O.synth_params(1, 4, ("decoder",), int(g["pseed"]), ds=True)
    else:
        p = O.synth_params(1, 4, ("main_decoder", "aux_decoder1", "aux_decoder2"), int(g["pseed"]))
    leaves = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
    if which == "ds":
        outs = O.unet_ds_forward(leaves, image, True, masks)
    else:
        torch.manual_seed(int(g["nseed"]))           # the reference's FeatureNoise draws, in feature order (unet.py:369)
        noises = [Uniform(-0.3, 0.3).sample((O.FT[i], hw >> i, hw >> i)) for i in range(5)]
        outs = O.unet_cct_3h_forward(leaves, image, True, masks, _chan(int(g["cseed"]), n), noises)
    assert len(outs) == (4 if which == "ds" else 3)
    for i, o in enumerate(outs):
        ref = torch.from_numpy(g[f"{which}:out{i}"])
        assert (o.detach() - ref).abs().max().item() < 5e-5 * max(1.0, ref.abs().max().item()), (which, i)
    loss = sum(O.pce_loss(o, label) for o in outs)
    assert abs(loss.item() - float(g[f"{which}:loss"])) < 2e-5 * abs(float(g[f"{which}:loss"]))
    keys = [str(k) for k in g[f"{which}:grad_keys"]]
    grads = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    for k, gr, (_, asum, l2) in zip(keys, grads, g[f"{which}:grad_stats"]):
        assert gr is not None, k
        assert abs(gr.double().norm().item() - l2) < 2e-3 * l2 + 1e-7, k
        assert abs(gr.double().abs().sum().item() - asum) < 2e-3 * asum + 1e-6, k
    unused = {str(k).split(".")[0] for k in g[f"{which}:no_grad_keys"]}
    assert unused == ({"decoder"} if which == "ds" else {"aux_decoder2"})      # out_conv_dp4 / the never-run third decoder


def test_pnet2d_matches_reference(golden_dir):
    """PNet2D restatement (dilated 3x3 blocks, concat, 1x1 heads, Dropout2d) against the fixture generated from the unmodified
    networks/pnet.py: eval and train logits, pCE loss, every parameter gradient."""
    g = _load(golden_dir, "pnet.npz")
    n = int(g["n"])
    p = O.pnet_synth_params(1, 4, int(g["pseed"]))
    image, label = torch.from_numpy(g["image"]), torch.from_numpy(g["label"])
    rs = np.random.RandomState(int(g["cseed"]))
    keeps = [torch.from_numpy((rs.uniform(size=(n, c)) >= 0.3).astype(np.uint8)) for c in (128, 64)]
    with torch.no_grad():
        ev = O.pnet2d_forward(p, image, False)
    assert (ev - torch.from_numpy(g["eval"])).abs().max().item() < 5e-5 * max(1.0, np.abs(g["eval"]).max())
    leaves = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
    o = O.pnet2d_forward(leaves, image, True, keeps)
    assert (o.detach() - torch.from_numpy(g["train"])).abs().max().item() < 5e-5 * max(1.0, np.abs(g["train"]).max())
    loss = O.pce_loss(o, label)
    assert abs(loss.item() - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    keys = [str(k) for k in g["grad_keys"]]
    grads = torch.autograd.grad(loss, [leaves[k] for k in keys])
    for k, gr, (_, asum, l2) in zip(keys, grads, g["grad_stats"]):
        assert abs(gr.double().norm().item() - l2) < 2e-3 * l2 + 1e-7, k
