import numpy as np
import torch

import wsl_oracle as O


def nhwc(x_nchw, dtype=torch.bfloat16):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def elem_masks_nchw(seed, n, h, w):
    rs = np.random.RandomState(seed)
    return [torch.from_numpy((rs.uniform(size=(n, O.FT[i], h >> i, w >> i)) >= O.ENC_DROP[i]).astype(np.uint8))
            for i in range(5)]


def chan_masks(seed, n):
    rs = np.random.RandomState(seed)
    return [torch.from_numpy((rs.uniform(size=(n, c)) >= 0.5).astype(np.uint8)) for c in O.FT]


ENC_MASK_KEYS = ["encoder.in_conv.conv_conv.3"] + [f"encoder.down{i}.maxpool_conv.1.conv_conv.3" for i in range(1, 5)]


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()
