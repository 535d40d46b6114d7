"""Drop-in for the reference's ``utils/gate_crf_loss.py`` (ModelLossSemsegGatedCRF).

The argument pattern every WSL4MIS script uses -- one kernel descriptor with 'xy' and one intensity
modality, radius 5, no masks, Potts compatibility, sample already at prediction resolution
(train_weakly_supervised_pCE_GatedCRFLoss_2D.py:103-104,115-122) -- runs the fused stencil kernel
(csrc/losses.cu: gatedcrf_kernel).  Anything else is rejected loudly rather than silently computed
another way.
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
        unsupported = []
        if mask_src is not None or mask_dst is not None:
            unsupported.append("masks")
        if compatibility is not None:
            unsupported.append("compatibility matrix")
        if custom_modality_downsamplers is not None:
            unsupported.append("custom downsamplers")
        if out_kernels_vis:
            unsupported.append("kernel visualisation")
        if len(kernels_desc) != 1:
            unsupported.append("multiple kernel descriptors")
        desc = kernels_desc[0]
        mods = [k for k in desc if k != 'weight']
        if 'xy' not in mods or len(mods) != 2:
            unsupported.append(f"modalities {mods}")
        if kernels_radius != 5:
            unsupported.append(f"radius {kernels_radius}")
        if tuple(sample.shape) != (N, 1, height_pred, width_pred):
            unsupported.append(f"sample shape {tuple(sample.shape)} (needs [N,1,H,W] at prediction size)")
        if unsupported:
            raise NotImplementedError("wsl4mis_b200 GatedCRF fused path does not cover: " + ", ".join(unsupported))
        other = [k for k in mods if k != 'xy'][0]
        loss = Fn.gated_crf(y_hat_softmax, sample, kernels_radius, float(desc['xy']), float(desc[other]),
                            float(desc['weight']))
        return {'loss': loss}
