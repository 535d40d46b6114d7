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


def dice_loss1(score, target):
    """utils/losses.py:19-27: Dice with un-squared sums in the denominator."""
    target = target.float()
    return 1 - (2 * torch.sum(score * target) + 1e-5) / (torch.sum(score) + torch.sum(target) + 1e-5)


def softmax_dice_loss(input_logits, target_logits):
    """utils/losses.py:39-55: class-averaged dice_loss1 between the two softmax maps."""
    assert input_logits.size() == target_logits.size()
    a, b = torch.softmax(input_logits, dim=1), torch.softmax(target_logits, dim=1)
    n = input_logits.shape[1]
    return sum(dice_loss1(a[:, i], b[:, i]) for i in range(n)) / n


def softmax_kl_loss(input_logits, target_logits, sigmoid=False):
    """utils/losses.py:85-104: F.kl_div(log q, p, reduction='mean') (element mean, as the reference calls it)."""
    assert input_logits.size() == target_logits.size()
    if sigmoid:
        logq, p = torch.log(torch.sigmoid(input_logits)), torch.sigmoid(target_logits)
    else:
        logq, p = torch.log_softmax(input_logits, dim=1), torch.softmax(target_logits, dim=1)
    return torch.nn.functional.kl_div(logq, p, reduction='mean')


class FocalLoss(nn.Module):
    """utils/losses.py:119-153: -(1 - p_t)^gamma * alpha_t * log p_t with p_t detached inside the modulating factor."""

    def __init__(self, gamma=2, alpha=None, size_average=True):
        super().__init__()
        self.gamma = gamma
        if isinstance(alpha, (float, int)):
            alpha = torch.tensor([alpha, 1 - alpha], dtype=torch.float32)
        elif isinstance(alpha, list):
            alpha = torch.tensor(alpha, dtype=torch.float32)
        self.alpha = alpha
        self.size_average = size_average

    def forward(self, input, target):
        if input.dim() > 2:                                   # [N, C, *] -> [N * prod(*), C]
            input = input.flatten(2).transpose(1, 2).reshape(-1, input.size(1))
        target = target.reshape(-1, 1)
        logpt = torch.log_softmax(input, dim=1).gather(1, target).reshape(-1)
        pt = logpt.detach().exp()
        if self.alpha is not None:
            self.alpha = self.alpha.to(device=input.device, dtype=input.dtype)
            logpt = logpt * self.alpha.gather(0, target.reshape(-1))
        loss = -((1 - pt) ** self.gamma) * logpt
        return loss.mean() if self.size_average else loss.sum()


class SizeLoss(nn.Module):
    """utils/losses.py:248-272: squared distance of the soft class sizes to a +-margin band around the label counts
    (5-D volumes [B, C, H, W, D]; every class must occur in every sample, as the reference's assignment requires)."""

    def __init__(self, margin=0.1):
        super().__init__()
        self.margin = margin

    def forward(self, output, target):
        soft_counts = torch.softmax(output, dim=1).sum(dim=(2, 3))
        want = torch.zeros_like(soft_counts)
        for b in range(target.shape[0]):
            _, counts = torch.unique(target[b], sorted=True, return_counts=True)
            assert target[b].numel() == int(counts.sum())
            want[b, :] = counts
        lo, hi = want * (1 - self.margin), want * (1 + self.margin)
        pen = (soft_counts < lo).float() * (soft_counts - lo) ** 2 + (soft_counts > hi).float() * (soft_counts - hi) ** 2
        return (pen[:, 1:] / (output.shape[2] * output.shape[3] * output.shape[4])).mean()      # background excluded


class SupConLoss(nn.Module):
    """utils/losses.py:311-398 (supervised contrastive loss; SimCLR when neither labels nor mask is given)."""

    def __init__(self, temperature=0.07, contrast_mode='all', base_temperature=0.07):
        super().__init__()
        self.temperature, self.contrast_mode, self.base_temperature = temperature, contrast_mode, base_temperature

    def forward(self, features, labels=None, mask=None):
        if features.dim() < 3:
            raise ValueError('`features` needs to be [bsz, n_views, ...],at least 3 dimensions are required')
        features = features.flatten(2)
        bsz, views = features.shape[0], features.shape[1]
        dev = features.device
        if labels is not None and mask is not None:
            raise ValueError('Cannot define both `labels` and `mask`')
        if labels is None and mask is None:
            pos = torch.eye(bsz, dtype=torch.float32, device=dev)
        elif labels is not None:
            labels = labels.contiguous().view(-1, 1)
            if labels.shape[0] != bsz:
                raise ValueError('Num of labels {} does not match num of features {}'.format(labels.shape[0], bsz))
            pos = torch.eq(labels, labels.T).float().to(dev)
        else:
            pos = mask.float().to(dev)
        contrast = torch.cat(torch.unbind(features, dim=1), dim=0)           # [views * bsz, D]
        if self.contrast_mode == 'one':
            anchor, n_anchor = features[:, 0], 1
        elif self.contrast_mode == 'all':
            anchor, n_anchor = contrast, views
        else:
            raise ValueError('Unknown mode: {}'.format(self.contrast_mode))
        sim = anchor @ contrast.T / self.temperature
        sim = sim - sim.max(dim=1, keepdim=True).values.detach()
        pos = pos.repeat(n_anchor, views)
        not_self = torch.ones_like(pos)
        idx = torch.arange(bsz * n_anchor, device=dev)
        not_self[idx, idx] = 0
        pos = pos * not_self
        log_prob = sim - torch.log((torch.exp(sim) * not_self).sum(1, keepdim=True))
        mean_pos = (pos * log_prob).sum(1) / pos.sum(1)
        return (-(self.temperature / self.base_temperature) * mean_pos).view(n_anchor, bsz).mean()
