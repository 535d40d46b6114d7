"""Diagnostic (not collected by pytest): per-stage error of the executor vs the CPU oracle on the golden case."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import wsl_oracle as O
from _gpu_util import chan_masks, elem_masks_nchw, ENC_MASK_KEYS, nchw
from wsl4mis_b200.networks.unet import UNet, UNet_CCT

DEV = "cuda"
for cct in (False, True):
    g = np.load(os.path.join(ROOT, "tests/golden", "unet_cct_dmpls.npz" if cct else "unet_pce_gatedcrf.npz"))
    decs = ("main_decoder", "aux_decoder1") if cct else ("decoder",)
    p = O.synth_params(1, 4, decs, int(g["pseed"]))
    n, hw = int(g["n"]), int(g["hw"])
    em = elem_masks_nchw(int(g["mseed"]), n, hw, hw)
    image = torch.from_numpy(g["image"])
    for use_tc in (False, True):
        m = (UNet_CCT if cct else UNet)(1, 4); m.load_state_dict(p); m = m.to(DEV); m.executor.use_tc = use_tc
        m.dropout_masks = {i: e.permute(0, 2, 3, 1).contiguous().to(DEV) for i, e in enumerate(em)}
        if cct:
            ck = chan_masks(int(g["cseed"]), n); m.channel_keep = [c.to(DEV) for c in ck]
        for mode in ("eval", "train"):
            m.train(mode == "train")
            with torch.no_grad():
                o = m(image.to(DEV))
            main = (o[0] if cct else o).cpu()
            ref = torch.from_numpy(g[mode + "_main"])
            print(f"cct={cct} tc={use_tc} {mode}: max|err|/max|ref| = {(main-ref).abs().max().item()/ref.abs().max().item():.5f}  "
                  f"rel_l2={((main-ref).norm()/ref.norm()).item():.5f} ref_scale={ref.abs().max().item():.3f}")
            if mode == "train" and use_tc:
                masks = {k: e for k, e in zip(ENC_MASK_KEYS, em)}
                with torch.no_grad():
                    feats = O.encoder_forward(p, image, True, masks)
                ex = m.executor
                for i, f in enumerate(feats):
                    key = [k for k in ex._bufs if k[1] == f"enc{i}.a2" and k[0] == "ng"][0]
                    mine = nchw(ex._bufs[key].cpu())
                    print(f"   enc{i}.a2 rel_l2 {((mine-f).norm()/f.norm()).item():.5f} max {(mine-f).abs().max().item():.4f} scale {f.abs().max().item():.3f}")
