"""Fused per-step training body (the hot loop of every code/train_weakly_supervised_*.py script).

``TrainStep`` replays, with the reference's arithmetic and hyper-parameters, the step body of
  pce            train_weakly_supervised_pCE_2D.py:97-108
  pce_gatedcrf   train_weakly_supervised_pCE_GatedCRFLoss_2D.py:111-130
  pce_ms         train_weakly_supervised_pCE_MumfordShah_Loss_2D.py:98-110
  pce_tv         train_weakly_supervised_pCE_TV_2D.py:109-121
  dmpls          train_weakly_supervised_segmentation_pCE_ours_proposed.py:108-132
as ONE launch sequence: executor forward -> fused loss head (softmax+pCE, regulariser, combined backward into
dlogits) -> executor backward into a flat fp32 gradient bucket -> (NCCL all-reduce of that bucket when
world_size > 1) -> fused SGD(momentum 0.9, wd 1e-4) over the flat parameter buffer with the poly LR read from
device memory.  The whole sequence is CUDA-graph captured after the first call (graph=True).

For models returning two heads (unet_cct) under a single-head script (pce_gatedcrf, ...) the loss is applied to
``main_seg`` only, as SURVEY F7 prescribes; the untouched aux decoder then has no gradient and, like
torch.optim.SGD with ``grad is None``, is skipped by the optimiser.
"""
from __future__ import annotations

import random

import torch

from . import ddp
from . import _lib
from ._lib import call, workspace

VARIANTS = ("pce", "pce_gatedcrf", "pce_ms", "pce_tv", "dmpls")


class TrainStep:
    def __init__(self, model, variant="pce_gatedcrf", base_lr=0.01, max_iterations=30000, momentum=0.9,
                 weight_decay=1e-4, graph=True, process_group=None, world_size=1):
        assert variant in VARIANTS
        self.model, self.variant = model, variant
        self.ex = model.executor
        self.base_lr, self.max_iterations = float(base_lr), int(max_iterations)
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        self.graph_enabled = bool(graph) and variant != "dmpls"   # dmpls draws a host-side beta every step
        self.world_size = int(world_size)
        self.pg = process_group
        self.iter_num = 0
        self._graphs = [None, None]
        self._warm = 0
        self._static = [None, None]
        dev = next(model.parameters()).device
        self.dev = dev
        # flat fp32 master parameters (views keep the nn.Parameter objects / state_dict intact)
        params = self.ex.params
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = {}
        for p in params:
            self.flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + p.numel()].view_as(p)
            self.offsets[id(p)] = (off, p.numel())
            off += p.numel()
        if self.world_size > 1:
            ddp.broadcast_flat(self.flat, 0, self.pg)     # identical replicas (DDP does the same at construction)
        self.mom = torch.zeros_like(self.flat)
        self.lr_dev = torch.full((1,), self.base_lr, dtype=torch.float32, device=dev)
        self.two_heads = len(self.ex.dec) == 2
        self.n_heads_trained = 2 if (variant == "dmpls") else 1
        if variant == "dmpls":
            assert self.two_heads, "dmpls needs a two-head model (unet_cct)"
        # parameter range that receives gradients (encoder + trained decoders are a prefix of parameters())
        if self.two_heads and self.n_heads_trained == 1:
            aux = self.model.aux_decoder1
            first_aux = next(aux.parameters())
            self.n_trained = self.offsets[id(first_aux)][0]
        else:
            self.n_trained = n
        self.loss_parts = {}
        self.launches_per_step = 0

    # ------------------------------------------------------------------
    def _head(self, logits_list, image, label, slot):
        """loss head: returns (loss tensor, [dlogits per decoder or None])."""
        ex = self.ex
        N, C, H, W = logits_list[0].shape
        dev = self.dev
        B = lambda name, shape, dt=torch.float32: ex.buf(slot, "head." + name, shape, dt)
        dl = [None] * len(logits_list)
        v = self.variant
        heads = range(self.n_heads_trained)
        probs, stats = [], []
        for h in heads:
            p = B(f"probs{h}", (N, C, H, W))
            st = B(f"stats{h}", (2,))
            call("wsl_softmax_pce_fwd", logits_list[h], label, p, N, C, H, W, 4, st, workspace("pce", dev))
            probs.append(p)
            stats.append(st)
        ce = stats[0][0]
        if v == "pce":
            loss = ce
            d = B("dl0", (N, C, H, W))
            call("wsl_head_bwd", probs[0], label, stats[0], None, 1.0, None, 0.0, N, C, H, W, 4, d)
            dl[0] = d
        elif v == "pce_gatedcrf":
            gp = B("gprobs0", (N, C, H, W))
            out = B("crf", (2,))
            call("wsl_gatedcrf_fwd", probs[0], image, gp, N, C, H, W, 5, 6.0, 0.1, 1.0, out, workspace("crf", dev))
            loss = ce + 0.1 * out[0]
            d = B("dl0", (N, C, H, W))
            call("wsl_head_bwd", probs[0], label, stats[0], None, 1.0, gp, 0.1, N, C, H, W, 4, d)
            dl[0] = d
            self.loss_parts = {"ce": ce, "crf": out[0]}
        elif v == "pce_ms":
            gp = B("gprobs0", (N, C, H, W))
            out = B("ms", (1,))
            cent = B("cent", (N * C,))
            call("wsl_mumford_shah_fwd", image, probs[0], N, C, H, W, out, cent, workspace("ms", dev))
            call("wsl_mumford_shah_bwd", image, probs[0], cent, N, C, H, W, 1e-6, 0, gp)
            loss = ce + 1e-6 * out[0]
            d = B("dl0", (N, C, H, W))
            call("wsl_head_bwd", probs[0], label, stats[0], None, 1.0, gp, 1.0, N, C, H, W, 4, d)
            dl[0] = d
        elif v == "pce_tv":
            # tv_loss(outputs_soft[1:]) -- batch slice, sample 0 gets no TV term
            gp = B("gprobs0", (N, C, H, W))
            gp.zero_()
            out = B("tv", (1,))
            if N > 1:
                call("wsl_tv_loss", probs[0][1:], (N - 1) * C, H, W, 1e-2, gp[1:], out, workspace("tv", dev))
                loss = ce + 1e-2 * out[0]
            else:
                loss = ce   # mean over an empty tensor is NaN in the reference; N=1 is never used with this script
            d = B("dl0", (N, C, H, W))
            call("wsl_head_bwd", probs[0], label, stats[0], None, 1.0, gp, 1.0, N, C, H, W, 4, d)
            dl[0] = d
        else:  # dmpls
            beta = random.random() + 1e-10            # host RNG, as the script (:117)
            pseudo = B("pseudo", (N, H, W), torch.uint8)
            call("wsl_mix_argmax", probs[0], probs[1], beta, 1.0 - beta, N, C, H, W, pseudo)
            loss = 0.5 * (stats[0][0] + stats[1][0])
            for h in heads:
                sums = B(f"pd{h}", (13,))
                call("wsl_pdice_fwd", probs[h], pseudo, None, float(N), N, C, H, W, sums, workspace("pdice", dev))
                gp = B(f"gprobs{h}", (N, C, H, W))
                call("wsl_pdice_bwd", probs[h], pseudo, None, float(N), sums, N, C, H, W, 0.25, 0, gp)
                loss = loss + 0.25 * sums[0]
                d = B(f"dl{h}", (N, C, H, W))
                call("wsl_head_bwd", probs[h], label, stats[h], None, 0.5, gp, 1.0, N, C, H, W, 4, d)
                dl[h] = d
            self.beta = beta
        return loss, dl

    def _fwd_bwd(self, image, label):
        ex = self.ex
        c0 = _lib.COUNTERS["launch_calls"]
        # single-head scripts on the two-head model never read the aux output: let its forward keep running on the
        # side stream underneath the loss head and the main backward (joined at the end of the backward)
        defer = self.two_heads and self.n_heads_trained == 1
        outs, slot = ex.forward(image, True, True, getattr(self.model, "dropout_masks", None),
                                getattr(self.model, "channel_keep", None), defer_join=defer)
        loss, dl = self._head(outs, image, label, slot)
        gflat = ex.backward(slot, dl)
        self._outs = outs
        self.launches_per_step = _lib.COUNTERS["launch_calls"] - c0 + 1     # + the SGD kernel (also inside the graph)
        return loss, gflat

    def _opt(self, gflat):
        n = self.n_trained
        call("wsl_sgd_step", self.flat, gflat, self.mom, n, self.lr_dev, self.base_lr, self.momentum, self.weight_decay,
             1.0 / self.world_size)

    def _allreduce(self, gflat):
        if self.world_size > 1:
            ddp.allreduce_flat(gflat[: self.n_trained], self.pg)

    # ------------------------------------------------------------------
    def __call__(self, image, label):
        """image: fp32 [N,1,H,W] CUDA, label: uint8 [N,H,W] CUDA.  Returns the loss (0-dim device tensor).

        Graph mode keeps TWO sets of static input buffers (``input_buffers(0/1)``) with one captured graph each, so a
        caller can copy batch k+1 from pinned host memory on a copy stream while the graph of batch k runs
        (bench.py's end-to-end leg does).  Tensors that are not one of those buffers are copied into set 0."""
        assert image.is_cuda and label.is_cuda and label.dtype == torch.uint8
        self.model.train()
        if not self.graph_enabled:
            loss, g = self._fwd_bwd(image, label)
            self._allreduce(g)
            self._opt(g)
        else:
            idx = 0
            for i, st in enumerate(self._static):
                if st is not None and st[0].data_ptr() == image.data_ptr():
                    idx = i
            simg, slab = self.input_buffers(idx, like=(image, label))
            assert simg.shape == image.shape, "graph mode needs a fixed batch shape"
            if simg.data_ptr() != image.data_ptr():
                simg.copy_(image, non_blocking=True)
            if slab.data_ptr() != label.data_ptr():
                slab.copy_(label, non_blocking=True)
            if self._warm < 2:                      # eager warm-up allocates every buffer / tensor map
                loss, g = self._fwd_bwd(simg, slab)
                self._allreduce(g)
                self._opt(g)
                self._warm += 1
            else:
                if self._graphs[idx] is None:
                    torch.cuda.synchronize()
                    g1 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1):
                        gloss, gg = self._fwd_bwd(simg, slab)
                        if self.world_size == 1:
                            self._opt(gg)
                    g2 = None
                    if self.world_size > 1:
                        g2 = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g2):
                            self._opt(gg)
                    self._graphs[idx] = (g1, g2, gloss, gg)
                g1, g2, gloss, gg = self._graphs[idx]
                g1.replay()
                if self.world_size > 1:
                    self._allreduce(gg)
                    g2.replay()
                loss = gloss
        # poly LR applied after the step with the pre-increment iteration (...pCE_2D.py:106-108)
        lr_ = self.base_lr * (1.0 - self.iter_num / self.max_iterations) ** 0.9
        self.lr_dev.fill_(lr_)
        self.iter_num += 1
        return loss

    def input_buffers(self, idx=0, like=None):
        """(image, label) device buffers read by captured graph `idx` (0 or 1)."""
        if self._static[idx] is None:
            src = like if like is not None else self._static[1 - idx]
            assert src is not None, "call the step once (or pass like=) before asking for input buffers"
            self._static[idx] = (torch.empty_like(src[0]), torch.empty_like(src[1]))
        return self._static[idx]

    def static_inputs(self):
        return self._static[0]

    @property
    def outputs(self):
        return self._outs
