"""autograd glue between torch tensors and the C-ABI loss kernels (csrc/losses.cu).

Every function here launches hand-written kernels through ``_lib.call``; there is no torch-op
fallback for the arithmetic.  torch supplies memory, streams and the autograd tape only.
"""
import torch

from ._lib import call, workspace


def _chk(t, dtype=torch.float32):
    assert t.is_cuda, "wsl4mis_b200 kernels need CUDA tensors (no CPU fallback)"
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _as_u8(label):
    assert label.is_cuda
    return label if label.dtype == torch.uint8 else label.to(torch.uint8)


# ------------------------------------------------------------------------------------------------
# softmax + partial cross entropy (K7)
# ------------------------------------------------------------------------------------------------
class _SoftmaxPCE(torch.autograd.Function):
    """(logits, label) -> (pCE loss, softmax probabilities).  The backward fuses the pCE gradient with
    the softmax Jacobian applied to whatever gradient arrives at the probabilities."""

    @staticmethod
    def forward(ctx, logits, label, ignore_index):
        logits = _chk(logits)
        label = _as_u8(label).contiguous()
        N, C, H, W = logits.shape
        probs = torch.empty_like(logits)
        stats = torch.empty(2, dtype=torch.float32, device=logits.device)
        call("wsl_softmax_pce_fwd", logits, label, probs, N, C, H, W, int(ignore_index), stats, workspace("pce", logits.device))
        ctx.save_for_backward(probs, label, stats)
        ctx.ignore_index = int(ignore_index)
        ctx.mark_non_differentiable(stats)
        return stats[0], probs, stats

    @staticmethod
    def backward(ctx, g_loss, g_probs, _g_stats):
        probs, label, stats = ctx.saved_tensors
        N, C, H, W = probs.shape
        d = torch.empty_like(probs)
        w_ce = 1.0
        if g_loss is None:
            w_ce, go = 0.0, None
        else:
            go = _chk(g_loss).reshape(1)
        gp = None if g_probs is None else _chk(g_probs)
        call("wsl_head_bwd", probs, label if w_ce else None, stats, go, w_ce, gp, 1.0, N, C, H, W, ctx.ignore_index, d, None, 0)
        return d, None, None


def softmax_pce(logits, label, ignore_index=4):
    """-> (loss, probs).  Equivalent to (CrossEntropyLoss(ignore_index)(logits, label.long()),
    torch.softmax(logits, 1)) of train_weakly_supervised_pCE_2D.py:98-100, in one kernel."""
    loss, probs, _ = _SoftmaxPCE.apply(logits, label, ignore_index)
    return loss, probs


class _SoftmaxOnly(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        logits = _chk(logits)
        N, C, H, W = logits.shape
        probs = torch.empty_like(logits)
        stats = torch.empty(2, dtype=torch.float32, device=logits.device)
        call("wsl_softmax_pce_fwd", logits, None, probs, N, C, H, W, 255, stats, workspace("pce", logits.device))
        ctx.save_for_backward(probs)
        return probs

    @staticmethod
    def backward(ctx, g):
        (probs,) = ctx.saved_tensors
        N, C, H, W = probs.shape
        d = torch.empty_like(probs)
        call("wsl_head_bwd", probs, None, None, None, 0.0, _chk(g), 1.0, N, C, H, W, 255, d, None, 0)
        return d


def softmax4(logits):
    """torch.softmax(logits, dim=1) for 4-class NCHW logits."""
    return _SoftmaxOnly.apply(logits)


# ------------------------------------------------------------------------------------------------
# Gated CRF (K8)
# ------------------------------------------------------------------------------------------------
class _GatedCRF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, image, radius, sigma_xy, sigma_rgb, weight):
        probs, image = _chk(probs), _chk(image)
        N, C, H, W = probs.shape
        g = torch.empty_like(probs)
        out = torch.empty(2, dtype=torch.float32, device=probs.device)
        call("wsl_gatedcrf_fwd", probs, image, g, N, C, H, W, int(radius), float(sigma_xy), float(sigma_rgb),
             float(weight), out, workspace("crf", probs.device))
        ctx.save_for_backward(g)
        return out[0]

    @staticmethod
    def backward(ctx, go):
        (g,) = ctx.saved_tensors
        return g * go, None, None, None, None, None


def gated_crf(probs, image, radius=5, sigma_xy=6.0, sigma_rgb=0.1, weight=1.0):
    return _GatedCRF.apply(probs, image, radius, sigma_xy, sigma_rgb, weight)


# ------------------------------------------------------------------------------------------------
# Mumford-Shah (K9)
# ------------------------------------------------------------------------------------------------
class _MumfordShah(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, probs):
        image, probs = _chk(image), _chk(probs)
        N, C, H, W = probs.shape
        out = torch.empty(1, dtype=torch.float32, device=probs.device)
        cent = torch.empty(N * C, dtype=torch.float32, device=probs.device)
        call("wsl_mumford_shah_fwd", image, probs, N, C, H, W, out, cent, workspace("ms", probs.device))
        ctx.save_for_backward(image, probs, cent)
        return out[0]

    @staticmethod
    def backward(ctx, go):
        image, probs, cent = ctx.saved_tensors
        N, C, H, W = probs.shape
        g = torch.empty_like(probs)
        call("wsl_mumford_shah_bwd", image, probs, cent, N, C, H, W, 1.0, 0, g)
        return None, g * go


def mumford_shah(image, probs):
    return _MumfordShah.apply(image, probs)


# ------------------------------------------------------------------------------------------------
# pseudo labels + (partial) Dice (K10)
# ------------------------------------------------------------------------------------------------
def mix_argmax(p1, p2=None, beta=1.0):
    """argmax(beta*p1 + (1-beta)*p2, dim=1) as a uint8 [N,H,W] map (no gradient)."""
    p1 = _chk(p1.detach())
    p2 = None if p2 is None else _chk(p2.detach())
    N, C, H, W = p1.shape
    out = torch.empty((N, H, W), dtype=torch.uint8, device=p1.device)
    call("wsl_mix_argmax", p1, p2, float(beta), float(1.0 - beta), None, N, C, H, W, out)
    return out


class _PDice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, target, ignore_index, masked):
        probs = _chk(probs)
        target = _as_u8(target).contiguous()
        N, C, H, W = probs.shape
        msum = None
        mconst = 1.0
        if masked:
            # batch-summed ignore mask (reference broadcast quirk, utils/losses.py:209-211,219-220)
            msum = torch.empty((H, W), dtype=torch.float32, device=probs.device)
            call("wsl_mask_count", target, N, H, W, int(ignore_index), msum)
        out = torch.empty(13, dtype=torch.float32, device=probs.device)
        call("wsl_pdice_fwd", probs, target, msum, mconst, N, C, H, W, out, workspace("pdice", probs.device))
        ctx.save_for_backward(probs, target, out, msum if msum is not None else torch.empty(0, device=probs.device))
        ctx.masked = masked
        return out[0]

    @staticmethod
    def backward(ctx, go):
        probs, target, out, msum = ctx.saved_tensors
        N, C, H, W = probs.shape
        g = torch.empty_like(probs)
        call("wsl_pdice_bwd", probs, target, msum if ctx.masked else None, 1.0, out, N, C, H, W, 1.0, 0, g)
        return g * go, None, None, None


def pdice(probs, target, ignore_index=4):
    """pDLoss(4, ignore_index)(probs, target[N,1,H,W])."""
    return _PDice.apply(probs, target.reshape(target.shape[0], *target.shape[-2:]), ignore_index, True)


def dice(probs, target):
    """DiceLoss(4)(probs, target[N,1,H,W]) (softmax=False, weight=None)."""
    return _PDice.apply(probs, target.reshape(target.shape[0], *target.shape[-2:]), 255, False)


# ------------------------------------------------------------------------------------------------
# TV / contour loss (K11)
# ------------------------------------------------------------------------------------------------
class _TV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs):
        probs = _chk(probs)
        planes = probs.shape[0] * probs.shape[1]
        H, W = probs.shape[-2:]
        out = torch.empty(1, dtype=torch.float32, device=probs.device)
        g = torch.zeros_like(probs)
        call("wsl_tv_loss", probs, planes, H, W, 1.0, g, out, workspace("tv", probs.device))
        ctx.save_for_backward(g)
        return out[0]

    @staticmethod
    def backward(ctx, go):
        (g,) = ctx.saved_tensors
        return g * go


def tv_loss(probs):
    return _TV.apply(probs)
