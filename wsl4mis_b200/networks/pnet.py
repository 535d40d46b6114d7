"""PNet2D with the reference's public names, constructor signature and state_dict layout (networks/pnet.py), executed by the planned
executor in _pnet_engine.py.  As in networks/unet.py the nn.Modules are parameter containers registered in the reference's order (so
``torch.manual_seed(s); PNet2D(...)`` draws the reference's initial weights and checkpoints interchange); the arithmetic of the whole
network is one launch sequence."""
import torch
import torch.nn as nn

from .unet import _Holder, _NetFn, _only_whole_network


class PNetBlock(nn.Module):
    """two dilated conv3x3 + BatchNorm + LeakyReLU (reference pnet.py:16-41); keys conv1, conv2, in1, in2"""

    def __init__(self, in_channels, out_channels, dilation, padding):
        super().__init__()
        self.in_chns, self.out_chns, self.dilation, self.padding = in_channels, out_channels, dilation, padding
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=padding, dilation=dilation, groups=1, bias=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=padding, dilation=dilation, groups=1, bias=True)
        self.in1 = nn.BatchNorm2d(out_channels)
        self.in2 = nn.BatchNorm2d(out_channels)
        self.ac1 = nn.LeakyReLU()
        self.ac2 = nn.LeakyReLU()

    forward = _only_whole_network


class ConcatBlock(nn.Module):
    """conv1x1 -> LeakyReLU -> conv1x1 -> LeakyReLU (reference pnet.py:44-59)"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_chns, self.out_chns = in_channels, out_channels
        self.conv1 = nn.Conv2d(in_channels, in_channels, kernel_size=1, padding=0)
        self.conv2 = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)
        self.ac1 = nn.LeakyReLU()
        self.ac2 = nn.LeakyReLU()

    forward = _only_whole_network


class OutPutBlock(nn.Module):
    """Dropout2d(0.3) -> conv1x1 -> LeakyReLU -> Dropout2d(0.3) -> conv1x1 (reference pnet.py:62-81)"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_chns, self.out_chns = in_channels, out_channels
        self.conv1 = nn.Conv2d(in_channels, in_channels // 2, kernel_size=1, padding=0)
        self.conv2 = nn.Conv2d(in_channels // 2, out_channels, kernel_size=1, padding=0)
        self.drop1 = nn.Dropout2d(0.3)
        self.drop2 = nn.Dropout2d(0.3)
        self.ac1 = nn.LeakyReLU()

    forward = _only_whole_network


class PNet2D(nn.Module):
    """reference pnet.py:84-122: PNet2D(in_chns, out_chns, num_filters, ratios) -> logits [N, out_chns, H, W].
    ``channel_keep`` (tests): the two [N, C] keep masks of the Dropout2d layers."""

    precision = "bf16"

    def __init__(self, in_chns, out_chns, num_filters, ratios):
        super().__init__()
        assert len(ratios) == 5
        self.in_chns, self.out_chns, self.ratios, self.num_filters = in_chns, out_chns, ratios, num_filters
        nf = num_filters
        for b, r in enumerate(ratios, 1):
            setattr(self, f"block{b}", PNetBlock(in_chns if b == 1 else nf, nf, r, padding=r))
        self.catblock = ConcatBlock(nf * 5, nf * 2)
        self.out = OutPutBlock(nf * 2, out_chns)
        self._holder = None
        self.dropout_masks = None
        self.channel_keep = None

    def set_precision(self, precision):
        """'bf16' (default), 'fp16' or 'fp16x3' (fp32-accurate tensor-core mode); see networks/_engine.py:PRECISIONS"""
        assert precision in ("bf16", "fp16", "fp16x3"), "PNet2D: bf16, fp16 or fp16x3"
        object.__setattr__(self, "precision", precision)
        object.__setattr__(self, "_holder", None)
        return self

    @property
    def executor(self):
        if self._holder is None:
            from ._pnet_engine import PNetExecutor
            object.__setattr__(self, "_holder", _Holder(PNetExecutor(self, self.precision)))
        return self._holder.executor

    def forward(self, x):
        self.executor
        params = list(self.parameters())
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _NetFn.apply(x, self._holder, self.training, need, self.dropout_masks, self.channel_keep, *params)[0]
