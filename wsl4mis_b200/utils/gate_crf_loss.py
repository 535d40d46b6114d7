"""Drop-in for the reference's ``utils/gate_crf_loss.py`` (ModelLossSemsegGatedCRF).

The argument pattern every WSL4MIS script uses -- one kernel descriptor with 'xy' and one intensity
modality, radius 5, no masks, Potts compatibility, sample already at prediction resolution
(train_weakly_supervised_pCE_GatedCRFLoss_2D.py:103-104,115-122) -- runs the fused stencil kernel
(csrc/losses.cu: gatedcrf_kernel).  Every other argument pattern the reference accepts (several descriptors, other
radii, source / destination masks, modalities larger than the prediction, kernel visualisation) takes `_general_forward`:
a shift-and-accumulate formulation in tensor expressions on the caller's device -- same results as
gate_crf_loss.py:20-188, nothing materialised beyond one [N,1,H,W] kernel slice per window offset ("anything
else may fall back to a straightforward implementation but must not change results").
"""
import torch

from .. import functional as Fn


class ModelLossSemsegGatedCRF(torch.nn.Module):
    def forward(self, y_hat_softmax, kernels_desc, kernels_radius, sample, height_input, width_input,
                mask_src=None, mask_dst=None, compatibility=None, custom_modality_downsamplers=None,
                out_kernels_vis=False):
        # same shape contract as utils/gate_crf_loss.py:51-57
        assert y_hat_softmax.dim() == 4, 'Prediction must be a NCHW batch'
        N, C, height_pred, width_pred = y_hat_softmax.shape
        assert width_input % width_pred == 0 and height_input % height_pred == 0 and \
            width_input * height_pred == height_input * width_pred, \
            f'[{width_input}x{height_input}] !~= [{width_pred}x{height_pred}]'
        desc = kernels_desc[0] if len(kernels_desc) == 1 else {}
        mods = [k for k in desc if k != 'weight']
        fast = (mask_src is None and mask_dst is None and compatibility is None and custom_modality_downsamplers is None
                and not out_kernels_vis and len(kernels_desc) == 1 and 'xy' in mods and len(mods) == 2 and kernels_radius == 5
                and C == 4 and torch.is_tensor(sample) and tuple(sample.shape) == (N, 1, height_pred, width_pred))
        # the scripts' pattern always takes the fused CUDA kernel (Fn.gated_crf raises for non-CUDA tensors: no CPU fallback
        # on the hot path); only the argument patterns no script uses go through the tensor-expression formulation
        if not fast:
            return _general_forward(y_hat_softmax, kernels_desc, kernels_radius, sample, height_input, width_input, mask_src,
                                    mask_dst, compatibility, custom_modality_downsamplers, out_kernels_vis)
        other = [k for k in mods if k != 'xy'][0]
        loss = Fn.gated_crf(y_hat_softmax, sample, kernels_radius, float(desc['xy']), float(desc[other]),
                            float(desc['weight']))
        return {'loss': loss}


def _downsample(img, modality, h, w, custom):
    """gate_crf_loss.py:126-132: area resize unless a custom downsampler is registered for the modality."""
    f = custom[modality] if custom is not None and modality in custom else torch.nn.functional.adaptive_avg_pool2d
    return f(img, (h, w))


def _shifted(t, dy, dx):
    """t[n, c, y + dy, x + dx] with zeros outside the map (F.unfold's zero padding, gate_crf_loss.py:183-188)."""
    H, W = t.shape[-2:]
    out = torch.zeros_like(t)
    ys, ye = max(0, -dy), min(H, H - dy)
    xs, xe = max(0, -dx), min(W, W - dx)
    if ys < ye and xs < xe:
        out[..., ys:ye, xs:xe] = t[..., ys + dy:ye + dy, xs + dx:xe + dx]
    return out


def _general_forward(y, kernels_desc, r, sample, height_input, width_input, mask_src, mask_dst, compatibility, custom, vis):
    N, C, H, W = y.shape
    dev = y.device
    # per-descriptor feature stacks (:134-161): 'xy' = pixel mesh (x = column, y = row), any other modality = `sample`
    # resized to the prediction; each divided by its sigma.  Out-of-bounds neighbours see feature 0 (zero padding).
    stacks = []
    for desc in kernels_desc:
        feats = []
        for modality, sigma in desc.items():
            if modality == 'weight':
                continue
            if modality == 'xy':
                xs = torch.arange(W, dtype=torch.float32, device=dev).view(1, 1, 1, W).expand(N, 1, H, W)
                ys = torch.arange(H, dtype=torch.float32, device=dev).view(1, 1, H, 1).expand(N, 1, H, W)
                f = torch.cat((xs, ys), 1)
            else:
                f = _downsample(sample, modality, H, W, custom)
            feats.append(f / sigma)
        stacks.append((desc['weight'], torch.cat(feats, 1)))

    def fix_mask(mask, name):
        assert mask.dim() == 4 and mask.shape[:2] == (N, 1) and mask.dtype == torch.float32, \
            f'{name} mask must be a NCHW batch with C=1 and dtype float32'
        if mask.shape[2:] != (H, W):
            mask = _downsample(mask, 'mask', H, W, custom)
        mask = torch.where(mask != mask, torch.zeros_like(mask), mask)       # NaN -> 0 (:73)
        return torch.where(mask < 1.0, torch.zeros_like(mask), mask)          # edges of an interpolated mask -> 0 (:75)

    denom = N * H * W
    if mask_src is not None:
        mask_src = fix_mask(mask_src, 'Source')
        denom = mask_src.sum().clamp(min=1)
    if mask_dst is not None:
        mask_dst = fix_mask(mask_dst, 'Destination')
        denom = mask_dst.sum().clamp(min=1)
    if compatibility is not None:
        assert compatibility.shape == (C, C), f'Compatibility matrix expected shape [{C}x{C}]'
        assert (compatibility < 0).int().sum() == 0, 'Compatibility matrix must not have negative values'
        # the reference evaluates `compatibility.diag.sum()` (gate_crf_loss.py:105): `diag` is a bound method, so every call
        # with a compatibility matrix ends in this AttributeError there; reproduced rather than silently "fixed"
        raise AttributeError("'builtin_function_or_method' object has no attribute 'sum'")
    D = 2 * r + 1
    ksum = torch.zeros((), dtype=torch.float32, device=dev)
    pair = torch.zeros((), dtype=torch.float32, device=dev)
    vis_t = None
    if vis:
        nh, nw = len(range(r, H, D)), len(range(r, W, D))
        vis_t = torch.zeros((N, 1, D * nh, D * nw), dtype=torch.float32, device=dev)
    for iy in range(D):
        for ix in range(D):
            if iy == r and ix == r:
                continue                                                     # centre kernel value is 0 (:171)
            dy, dx = iy - r, ix - r
            k = None
            for weight, f in stacks:
                d = _shifted(f, dy, dx) - f
                kk = weight * torch.exp(-0.5 * (d * d).sum(1, keepdim=True))
                k = kk if k is None else k + kk
            if mask_src is not None:
                k = k * _shifted(mask_src, dy, dx)
            if mask_dst is not None:
                k = k * mask_dst
            ksum = ksum + k.sum()
            pair = pair + (k * (_shifted(y, dy, dx) * y).sum(1, keepdim=True)).sum()
            if vis_t is not None:
                vis_t[:, :, iy::D, ix::D] = k[:, :, r::D, r::D]
    out = {'loss': (ksum - pair) / denom}                                   # Potts shortcut (:95-99)
    if vis:
        v = vis_t[:, :, :H, :W]
        if v.shape[2:] != (H, W):
            v = torch.nn.functional.pad(v, [0, W - v.shape[3], 0, H - v.shape[2]])
        out['kernels_vis'] = torch.nn.functional.interpolate(v, (height_input, width_input), mode='nearest')
    return out
