"""Drop-in for the hot-path part of the reference's ``utils/losses.py``, on the fused CUDA kernels.

Names and call signatures follow the reference (file:line cited per symbol).  Functions that are not on
the north-star hot path keep the reference's name and semantics using plain tensor expressions so that
``from utils import losses`` keeps working for every script.
"""
import math

import torch
import torch.nn as nn

from .. import functional as Fn


class pDLoss(nn.Module):
    """utils/losses.py:195-232 (partial Dice with ignore_index; batch-summed mask quirk reproduced)."""

    def __init__(self, n_classes, ignore_index):
        super().__init__()
        assert n_classes == 4, "fused kernels are compiled for 4 classes"
        self.n_classes, self.ignore_index = n_classes, ignore_index

    def forward(self, inputs, target, weight=None):
        assert weight is None or all(w == 1 for w in weight), "class weights other than 1 are not on the fused path"
        assert inputs.shape[0] == target.shape[0] and inputs.shape[2:] == target.shape[2:], \
            'predict & target shape do not match'
        return Fn.pdice(inputs, target, self.ignore_index)


class DiceLoss(nn.Module):
    """utils/losses.py:156-192."""

    def __init__(self, n_classes):
        super().__init__()
        assert n_classes == 4, "fused kernels are compiled for 4 classes"
        self.n_classes = n_classes

    def forward(self, inputs, target, weight=None, softmax=False):
        assert weight is None or all(w == 1 for w in weight), "class weights other than 1 are not on the fused path"
        if softmax:
            inputs = Fn.softmax4(inputs)
        assert inputs.shape[0] == target.shape[0] and inputs.shape[2:] == target.shape[2:], \
            'predict & target shape do not match'
        return Fn.dice(inputs, target)


class MumfordShah_Loss(nn.Module):
    """utils/losses.py:275-309; forward(image, prediction) with the reference's argument roles."""

    def forward(self, image, prediction):
        assert image.shape[1] == 1, "fused Mumford-Shah expects a single-channel image"
        return Fn.mumford_shah(image, prediction)


class PartialCrossEntropy(nn.Module):
    """CrossEntropyLoss(ignore_index=4) as constructed by every script (train_weakly_supervised_pCE_2D.py:81).
    The scripts import torch's class directly; this module is the fused replacement used by the engine and
    offered to scripts that opt in."""

    def __init__(self, ignore_index=4):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, logits, target):
        return Fn.softmax_pce(logits, target, self.ignore_index)[0]


def tv_loss(predication):
    """script-local tv_loss (train_weakly_supervised_pCE_TV_2D.py:58-65)."""
    return Fn.tv_loss(predication)


# ---- helpers that are off the hot path: same names/semantics, tensor expressions -----------------------
def entropy_loss(p, C=2):
    """utils/losses.py:30-36."""
    return (-(p * torch.log(p + 1e-6)).sum(1) / math.log(C)).mean()


def entropy_loss_map(p, C=2):
    """utils/losses.py:58-62."""
    return -(p * torch.log(p + 1e-6)).sum(1, keepdim=True) / math.log(C)


def entropy_minmization(p):
    """utils/losses.py:235-239."""
    return (-(p * torch.log(p + 1e-6)).sum(1)).mean()


def entropy_map(p):
    """utils/losses.py:242-245."""
    return -(p * torch.log(p + 1e-6)).sum(1, keepdim=True)


def softmax_mse_loss(input_logits, target_logits, sigmoid=False):
    """utils/losses.py:65-82."""
    assert input_logits.size() == target_logits.size()
    if sigmoid:
        a, b = torch.sigmoid(input_logits), torch.sigmoid(target_logits)
    else:
        a, b = Fn.softmax4(input_logits), Fn.softmax4(target_logits)
    return (a - b) ** 2


def symmetric_mse_loss(input1, input2):
    """utils/losses.py:107-117."""
    assert input1.size() == input2.size()
    return torch.mean((input1 - input2) ** 2)


def dice_loss(score, target):
    """utils/losses.py:8-16."""
    target = target.float()
    inter = torch.sum(score * target)
    return 1 - (2 * inter + 1e-5) / (torch.sum(score * score) + torch.sum(target * target) + 1e-5)
