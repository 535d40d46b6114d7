"""Generate golden fixtures from the UNMODIFIED reference (runs only where /root/reference exists).

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference modules are imported read-only from /root/reference/code.  Randomness that the
reference draws from torch's RNG (nn.Dropout, F.dropout2d) is made reproducible by temporarily
replacing ``torch.nn.functional.dropout`` / ``dropout2d`` with functions that apply keep-masks drawn
from a numpy RandomState stream (so the fixtures only store a seed); no reference source is edited
or copied.  Parameters come from ``wsl_oracle.synth_params`` (numpy stream) and are loaded with
``load_state_dict`` so the fixtures do not need to store 2.4 M weights.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/code"
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import wsl_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def elem_masks(seed, n, h, w):
    """Keep masks (uint8) for the five encoder nn.Dropout layers, keyed like the oracle expects."""
    rs = np.random.RandomState(seed)
    names = ["encoder.in_conv.conv_conv.3"] + [f"encoder.down{i}.maxpool_conv.1.conv_conv.3" for i in range(1, 5)]
    out = {}
    for i, nm in enumerate(names):
        keep = rs.uniform(size=(n, O.FT[i], h >> i, w >> i)) >= O.ENC_DROP[i]
        out[nm] = torch.from_numpy(keep.astype(np.uint8))
    return out


def chan_masks(seed, n):
    rs = np.random.RandomState(seed)
    return [torch.from_numpy((rs.uniform(size=(n, c)) >= 0.5).astype(np.uint8)) for c in O.FT]


class PatchedDropout:
    """Context manager: F.dropout / F.dropout2d consume queued keep-masks instead of torch RNG."""

    def __init__(self, elem, chan):
        self.elem = [elem[k] for k in sorted(elem, key=lambda s: ("down" in s, s))] if elem else []
        self.chan = list(chan) if chan else []

    def __enter__(self):
        self._d, self._d2 = F.dropout, F.dropout2d
        eq, cq = list(self.elem), list(self.chan)

        def dropout(x, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return x
            m = eq.pop(0)
            assert m.shape == x.shape, (m.shape, x.shape)
            return x * m.to(x.dtype) * (1.0 / (1.0 - p))

        def dropout2d(x, p=0.5, training=True, inplace=False):
            m = cq.pop(0)
            return x * (m.to(x.dtype) * (1.0 / (1.0 - p)))[:, :, None, None]

        F.dropout, F.dropout2d = dropout, dropout2d
        torch.nn.functional.dropout, torch.nn.functional.dropout2d = dropout, dropout2d
        return self

    def __exit__(self, *a):
        F.dropout, F.dropout2d = self._d, self._d2
        torch.nn.functional.dropout, torch.nn.functional.dropout2d = self._d, self._d2


def grad_summary(named_grads):
    """Per-parameter (sum, abs-sum, l2) plus full copies of the small tensors."""
    keys, stats, small = [], [], {}
    for k, g in named_grads.items():
        g = g.double()
        keys.append(k)
        stats.append([g.sum().item(), g.abs().sum().item(), g.norm().item()])
        if g.numel() <= 600:
            small["g:" + k] = g.float().numpy()
    return keys, np.asarray(stats, dtype=np.float64), small


def script_function(script, name):
    """A script-local function of the reference (the scripts parse argv and import tensorboardX at module level, so they cannot
    be imported): its `def` block is cut out of the UNMODIFIED source text and exec'd in a namespace holding torch / F."""
    import re
    src = open(os.path.join(REF, script)).read()
    m = re.search(r"^def %s\(.*?(?=^\S)" % name, src, flags=re.S | re.M)
    assert m, (script, name)
    ns = {"torch": torch, "F": F, "np": np, "nn": torch.nn}
    exec(compile(m.group(0), script, "exec"), ns)
    return ns[name]


def losses_api():
    """API-surface losses of utils/losses.py that no WSS script calls on the hot path (kept importable by wsl4mis_b200.utils.losses)."""
    from utils import losses as RL
    g = torch.Generator().manual_seed(4321)
    a, b = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    tgt = torch.randint(0, 4, (2, 8, 8), generator=g)
    out = {"a": a.numpy(), "b": b.numpy(), "target": tgt.numpy()}
    sa, sb = torch.softmax(a, 1), torch.softmax(b, 1)
    out["dice_loss1"] = np.float64(RL.dice_loss1(sa[:, 1], sb[:, 1]).item())
    out["softmax_dice_loss"] = np.float64(RL.softmax_dice_loss(a, b).item())
    out["softmax_kl_loss"] = np.float64(RL.softmax_kl_loss(a, b).item())
    out["softmax_kl_loss_sigmoid"] = np.float64(RL.softmax_kl_loss(a, b, sigmoid=True).item())
    out["focal"] = np.float64(RL.FocalLoss(gamma=2)(a, tgt).item())
    out["focal_alpha_sum"] = np.float64(RL.FocalLoss(gamma=1.5, alpha=[0.1, 0.2, 0.3, 0.4], size_average=False)(a, tgt).item())
    # SizeLoss sums over dims (2, 3) of a 5-D volume and assigns the per-sample label counts along the LAST axis
    # (losses.py:254-260): it only runs when #distinct labels == D; reproduced as written
    vol = torch.randn(2, 3, 4, 4, 3, generator=g)
    vt = torch.randint(0, 3, (2, 1, 4, 4, 3), generator=g)
    vt[:, 0, 0, 0, :3] = torch.arange(3)          # every class occurs in every sample
    out["vol"], out["vol_target"] = vol.numpy(), vt.numpy()
    out["size_loss"] = np.float64(RL.SizeLoss(0.1)(vol, vt).item())
    feats = torch.nn.functional.normalize(torch.randn(6, 2, 16, generator=g), dim=2)
    labels = torch.tensor([0, 1, 0, 2, 1, 2])
    out["feats"], out["feat_labels"] = feats.numpy(), labels.numpy()
    out["supcon_labels"] = np.float64(RL.SupConLoss()(feats, labels).item())
    out["supcon_simclr"] = np.float64(RL.SupConLoss()(feats).item())
    out["supcon_one"] = np.float64(RL.SupConLoss(contrast_mode="one")(feats, labels).item())
    np.savez_compressed(os.path.join(OUT, "losses_api.npz"), **out)
    print("losses_api:", {k: float(v) for k, v in out.items() if np.ndim(v) == 0})


def crf_general():
    """ModelLossSemsegGatedCRF beyond the scripts' argument pattern (gate_crf_loss.py:20-188): two kernel descriptors, radius 3,
    a modality at twice the prediction resolution (area downsampling), source / destination masks, kernel visualisation."""
    from utils.gate_crf_loss import ModelLossSemsegGatedCRF
    g = torch.Generator().manual_seed(99)
    y = torch.softmax(torch.randn(2, 3, 12, 20, generator=g), 1).requires_grad_(True)
    sample = torch.rand(2, 2, 24, 40, generator=g)
    desc = [{"weight": 0.9, "xy": 6, "rgb": 0.1}, {"weight": 0.1, "xy": 4}]
    ms = (torch.rand(2, 1, 12, 20, generator=g) > 0.2).float()
    md = (torch.rand(2, 1, 24, 40, generator=g) > 0.1).float()
    out = {"y": y.detach().numpy(), "sample": sample.numpy(), "mask_src": ms.numpy(), "mask_dst": md.numpy()}
    m = ModelLossSemsegGatedCRF()
    r = m(y, desc, 3, sample, 24, 40)
    (gy,) = torch.autograd.grad(r["loss"], y)
    out["plain:loss"], out["plain:grad"] = np.float64(r["loss"].item()), gy.numpy()
    r = m(y, desc, 3, sample, 24, 40, mask_src=ms.clone(), mask_dst=md.clone(), out_kernels_vis=True)
    (gy,) = torch.autograd.grad(r["loss"], y)
    out["masked:loss"], out["masked:grad"], out["masked:vis"] = np.float64(r["loss"].item()), gy.numpy(), r["kernels_vis"].detach().numpy()
    r = m(y, [{"weight": 1, "xy": 6, "rgb": 0.1}], 2, sample[:, :1, ::2, ::2].contiguous(), 12, 20)
    out["r2:loss"] = np.float64(r["loss"].item())
    try:
        m(y, desc, 3, sample, 24, 40, compatibility=torch.ones(3, 3) - torch.eye(3))
        out["compat:error"] = np.array("none")
    except Exception as e:                      # gate_crf_loss.py:105 calls `.diag.sum()` on a bound method
        out["compat:error"] = np.array(type(e).__name__)
    np.savez_compressed(os.path.join(OUT, "crf_general.npz"), **out)
    print("crf_general:", {k: (float(v) if np.ndim(v) == 0 and v.dtype.kind == "f" else str(v) if np.ndim(v) == 0 else v.shape) for k, v in out.items()})


def losses_kat():
    from utils import losses as RL
    from utils.gate_crf_loss import ModelLossSemsegGatedCRF

    g = torch.Generator().manual_seed(1234)
    logits = torch.randn(2, 4, 32, 32, generator=g)
    img = torch.rand(2, 1, 32, 32, generator=g)
    lab = torch.full((2, 32, 32), 4, dtype=torch.uint8)
    m = torch.rand(2, 32, 32, generator=g) < 0.1
    lab[m] = torch.randint(0, 4, (int(m.sum()),), generator=g, dtype=torch.uint8)
    logits2 = torch.randn(2, 4, 32, 32, generator=g)

    lg = logits.clone().requires_grad_(True)
    s = torch.softmax(lg, 1)
    out = {"logits": logits.numpy(), "logits2": logits2.numpy(), "image": img.numpy(), "label": lab.numpy()}

    tv_ref = script_function("train_weakly_supervised_pCE_TV_2D.py", "tv_loss")

    ce = torch.nn.CrossEntropyLoss(ignore_index=4)(lg, lab.long())
    crf = ModelLossSemsegGatedCRF()(s, [{"weight": 1, "xy": 6, "rgb": 0.1}], 5, img, 32, 32)["loss"]
    ms = RL.MumfordShah_Loss()(img, s)
    pseudo = torch.argmax(s.detach(), 1, keepdim=True)
    pd = RL.pDLoss(4, 4)(s, pseudo)
    # pDLoss with ignored pixels: mark label==4 pixels ignored, others by scribble label
    pd_ign = RL.pDLoss(4, 4)(s, lab.long().unsqueeze(1))
    dice = RL.DiceLoss(4)(s, pseudo)
    ent = RL.entropy_minmization(s)
    mse = RL.softmax_mse_loss(lg, logits2)
    names = ["pce", "gatedcrf", "mumford_shah", "pdice_argmax", "pdice_ignore", "dice", "entropy", "tv"]
    vals = [ce, crf, ms, pd, pd_ign, dice, ent, tv_ref(s)]
    for nm, v in zip(names, vals):
        out["loss:" + nm] = np.float64(v.item())
        (gr,) = torch.autograd.grad(v, lg, retain_graph=True)
        out["grad:" + nm] = gr.numpy()
    out["softmax_mse"] = mse.detach().numpy()
    (gr,) = torch.autograd.grad(mse.sum(), lg, retain_graph=True)
    out["grad:softmax_mse_sum"] = gr.numpy()
    beta = 0.37
    s2 = torch.softmax(logits2, 1)
    out["beta"] = np.float64(beta)
    out["pseudo_mix"] = torch.argmax(beta * s.detach() + (1 - beta) * s2, 1).numpy().astype(np.uint8)
    # composite step losses
    tot = ce + 0.1 * crf
    (gr,) = torch.autograd.grad(tot, lg, retain_graph=True)
    out["loss:step_pce_gatedcrf"] = np.float64(tot.item())
    out["grad:step_pce_gatedcrf"] = gr.numpy()
    # non-square / OOB-heavy case for the CRF border quirk (F10)
    g2 = torch.Generator().manual_seed(77)
    y2 = torch.softmax(torch.randn(1, 4, 16, 48, generator=g2), 1).requires_grad_(True)
    i2 = torch.rand(1, 1, 16, 48, generator=g2)
    c2 = ModelLossSemsegGatedCRF()(y2, [{"weight": 1, "xy": 6, "rgb": 0.1}], 5, i2, 16, 48)["loss"]
    (g2y,) = torch.autograd.grad(c2, y2)
    out["crf2:y"], out["crf2:image"] = y2.detach().numpy(), i2.numpy()
    out["crf2:loss"], out["crf2:grad_y"] = np.float64(c2.item()), g2y.numpy()
    np.savez_compressed(os.path.join(OUT, "losses_kat.npz"), **out)
    print("losses_kat:", {k: float(out[k]) for k in out if k.startswith("loss:")})


def net_golden(cct, n=2, hw=32, pseed=11, mseed=5, cseed=9):
    from networks.unet import UNet, UNet_CCT

    decs = ("main_decoder", "aux_decoder1") if cct else ("decoder",)
    params = O.synth_params(1, 4, decs, pseed)
    model = (UNet_CCT if cct else UNet)(1, 4)
    assert list(model.state_dict().keys()) == list(params.keys()), "oracle key order != reference state_dict"
    model.load_state_dict(params)
    image, label = O.synth_batch(n, hw, hw, seed=2022, frac=0.06)
    em = elem_masks(mseed, n, hw, hw)
    cm = chan_masks(cseed, n) if cct else None
    out = {"pseed": pseed, "mseed": mseed, "cseed": cseed, "n": n, "hw": hw,
           "image": image.numpy(), "label": label.numpy()}

    # eval forward (main head deterministic; aux head needs the channel masks even in eval, F5)
    model.eval()
    with torch.no_grad(), PatchedDropout(None, cm):
        o = model(image)
    if cct:
        out["eval_main"], out["eval_aux"] = o[0].numpy(), o[1].numpy()
    else:
        out["eval_main"] = o.numpy()

    # train forward + pCE+GatedCRF (+ DMPLS for cct) backward
    model.train()
    from utils import losses as RL
    from utils.gate_crf_loss import ModelLossSemsegGatedCRF
    with PatchedDropout(em, cm):
        o = model(image)
    main, aux = (o if cct else (o, None))
    out["train_main"] = main.detach().numpy()
    ce = torch.nn.CrossEntropyLoss(ignore_index=4)
    soft = torch.softmax(main, 1)
    if cct:
        out["train_aux"] = aux.detach().numpy()
        beta = 0.4321
        soft2 = torch.softmax(aux, 1)
        loss_ce = 0.5 * (ce(main, label.long()) + ce(aux, label.long()))
        pseudo = torch.argmax(beta * soft.detach() + (1 - beta) * soft2.detach(), 1)
        pdl = RL.pDLoss(4, 4)
        loss = loss_ce + 0.5 * 0.5 * (pdl(soft, pseudo.unsqueeze(1)) + pdl(soft2, pseudo.unsqueeze(1)))
        out["beta"] = np.float64(beta)
        out["pseudo"] = pseudo.numpy().astype(np.uint8)
    else:
        crf = ModelLossSemsegGatedCRF()(soft, [{"weight": 1, "xy": 6, "rgb": 0.1}], 5, image, hw, hw)["loss"]
        loss = ce(main, label.long()) + 0.1 * crf
    model.zero_grad()
    loss.backward()
    out["loss"] = np.float64(loss.item())
    grads = {k: p.grad for k, p in model.named_parameters()}
    keys, stats, small = grad_summary(grads)
    out["grad_keys"] = np.array(keys)
    out["grad_stats"] = stats
    out.update(small)
    sd = model.state_dict()
    for k in ("encoder.in_conv.conv_conv.1.running_mean", "encoder.in_conv.conv_conv.1.running_var",
              "encoder.down4.maxpool_conv.1.conv_conv.5.running_mean",
              "encoder.down4.maxpool_conv.1.conv_conv.5.running_var"):
        out["stat:" + k] = sd[k].numpy()
    name = "unet_cct_dmpls" if cct else "unet_pce_gatedcrf"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", out["loss"], "bytes", os.path.getsize(os.path.join(OUT, name + ".npz")))


def augment_golden(n=36, out_hw=(48, 56), seed=2022):
    """Training-sample transform of the reference (dataloaders/dataset_semi.py:146-171) on seeded ragged slices.  The
    module imports h5py at the top (absent here, unused by the transform): an empty stub module stands in for it."""
    import random
    import types

    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import warnings
    warnings.filterwarnings("ignore")
    from dataloaders import dataset_semi as D
    import augment_oracle as A

    ims, lbs = A.synth_slices(n, seed)
    tf = D.RandomGenerator(out_hw)
    out_i, out_l = [], []
    for i, (im, lb) in enumerate(zip(ims, lbs)):
        random.seed(1000 + i)
        np.random.seed(1000 + i)
        r = tf({"image": im, "label": lb})
        out_i.append(r["image"].numpy()[0])
        out_l.append(r["label"].numpy())
    np.savez_compressed(os.path.join(OUT, "augment.npz"), n=n, seed=seed, out_hw=np.array(out_hw),
                        image=np.stack(out_i).astype(np.float32), label=np.stack(out_l).astype(np.uint8))
    print("augment.npz:", np.stack(out_i).shape)


def heads_golden(n=2, hw=32, pseed=31, mseed=6, cseed=8, nseed=77):
    """UNet_DS (4 outputs) and UNet_CCT_3H (3 outputs) of the unmodified reference: train-mode forward, sum of the pCE of
    every head, gradient summaries.  FeatureNoise draws from torch's global RNG: the fixture stores the seed (the only
    torch-RNG consumers left in the patched forward are the five noise tensors, in feature order)."""
    from networks.unet import UNet_DS, UNet_CCT_3H

    image, label = O.synth_batch(n, hw, hw, seed=4, frac=0.1)
    out = {"n": n, "hw": hw, "pseed": pseed, "mseed": mseed, "cseed": cseed, "nseed": nseed,
           "image": image.numpy(), "label": label.numpy()}
    ce = torch.nn.CrossEntropyLoss(ignore_index=4)
    for name, cls, decs, ds in (("ds", UNet_DS, ("decoder",), True),
                                ("3h", UNet_CCT_3H, ("main_decoder", "aux_decoder1", "aux_decoder2"), False)):
        p = O.synth_params(1, 4, decs, pseed, ds=ds)
        m = cls(1, 4)
        assert list(m.state_dict().keys()) == list(p.keys()), name
        m.load_state_dict(p)
        m.train()
        em = elem_masks(mseed, n, hw, hw)
        with PatchedDropout(em, chan_masks(cseed, n) if name == "3h" else None):
            torch.manual_seed(nseed)
            outs = m(image)
        loss = sum(ce(o, label.long()) for o in outs)
        loss.backward()
        out[f"{name}:loss"] = loss.item()
        for i, o in enumerate(outs):
            out[f"{name}:out{i}"] = o.detach().numpy()
        named = {k: v.grad for k, v in m.named_parameters() if v.grad is not None}
        keys, stats, _ = grad_summary(named)
        out[f"{name}:grad_keys"] = np.array(keys)
        out[f"{name}:grad_stats"] = stats
        out[f"{name}:no_grad_keys"] = np.array([k for k, v in m.named_parameters() if v.grad is None])
    np.savez_compressed(os.path.join(OUT, "unet_heads.npz"), **out)
    print("unet_heads.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if ":out" in k})


def pnet_golden(n=2, hw=32, pseed=43, cseed=12):
    """PNet2D(1, 4, 64, [1, 2, 4, 8, 16]) as net_factory builds it (net_factory.py:18-19): train-mode forward with fixed Dropout2d
    masks, pCE loss, every parameter gradient; eval-mode forward."""
    from networks.pnet import PNet2D
    rs = np.random.RandomState(cseed)
    keeps = [torch.from_numpy((rs.uniform(size=(n, c)) >= 0.3).astype(np.uint8)) for c in (128, 64)]
    p = O.pnet_synth_params(1, 4, pseed)
    m = PNet2D(1, 4, 64, [1, 2, 4, 8, 16])
    assert list(m.state_dict().keys()) == list(p.keys())
    m.load_state_dict(p)
    image, label = O.synth_batch(n, hw, hw, seed=pseed + 1, frac=0.1)
    out = {"n": n, "hw": hw, "pseed": pseed, "cseed": cseed, "image": image.numpy(), "label": label.numpy()}
    m.eval()
    with torch.no_grad():
        out["eval"] = m(image).numpy()
    m.train()

    class _P(PatchedDropout):          # Dropout2d(0.3): keep / 0.7
        def __enter__(self):
            super().__enter__()
            cq = list(self.chan)

            def dropout2d(x, p=0.5, training=True, inplace=False):
                if not training:
                    return x
                k = cq.pop(0)
                return x * (k.to(x.dtype) * (1.0 / (1.0 - p)))[:, :, None, None]
            F.dropout2d = dropout2d
            torch.nn.functional.dropout2d = dropout2d
            return self

    with _P(None, keeps):
        o = m(image)
    loss = torch.nn.CrossEntropyLoss(ignore_index=4)(o, label.long())
    loss.backward()
    out["train"], out["loss"] = o.detach().numpy(), np.float64(loss.item())
    keys, stats, small = grad_summary({k: v.grad for k, v in m.named_parameters()})
    out["grad_keys"], out["grad_stats"] = np.array(keys), stats
    np.savez_compressed(os.path.join(OUT, "pnet.npz"), **out)
    print("pnet.npz: loss", float(out["loss"]), "train logits max", float(np.abs(out["train"]).max()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "heads":
        heads_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "augment":
        augment_golden()
        sys.exit(0)
    losses_kat()
    losses_api()
    crf_general()
    net_golden(False)
    net_golden(True)
    augment_golden()
    heads_golden()
    pnet_golden()
