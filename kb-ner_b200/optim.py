"""Fused optimizer step of the fine-tuning path.

Reference semantics (flair/trainers/finetune_trainer.py): two parameter groups keyed on names (:552-571) --
group 0 = parameters whose name neither contains 'embedding' nor is linear.weight / linear.bias (i.e. the CRF
`transitions`) at lr * lr_rate, group 1 = the rest (the whole encoder sits under `embeddings.*`, plus the tag
projection) at lr; transformers-3.0.0 AdamW (betas 0.9/0.999, eps 1e-6, bias correction, decoupled weight decay 0);
`clip_grad_norm_(model.parameters(), 5.0)` over ALL parameters before the step (:1010); loss / accumulation steps
(:939-946); linear decay to zero over t_total steps with 0 warm-up (:679-688).

Here every group is a flat fp32 arena: one `sumsq` launch per arena for the global norm, one tiny kernel that turns it
into the clip coefficient ON THE DEVICE (no host sync), one fused AdamW launch per arena.
"""
import math

import torch

from . import ops
from .encoder import ParamArena


class FusedAdamW:
    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, max_grad_norm=5.0):
        """groups: list of dicts {"arena": ParamArena, "lr": float}."""
        self.groups = []
        for g in groups:
            ar = g["arena"]
            self.groups.append({"arena": ar, "lr": float(g["lr"]), "base_lr": float(g["lr"]),
                                "m": torch.zeros_like(ar.flat), "v": torch.zeros_like(ar.flat)})
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.steps = 0
        dev = self.groups[0]["arena"].flat.device
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._partials = torch.zeros(2048, dtype=torch.float32, device=dev)    # fixed-order norm: same bits on every rank
        self._coef = torch.ones(1, dtype=torch.float32, device=dev)

    def _segments(self, arena, grads, i):
        """The gradient of group i as (lo, hi, buffer) segments of its arena: the arena's own fp32 .grad, one exchanged
        buffer, or the segment list GradExchange returns when the embedding table's rows went the sparse way.  With row
        skipping enabled on the arena (ParamArena.enable_row_skipping) the local fp32 gradient is split around the table."""
        n = arena.grad.numel()
        rt = arena.row_table
        if grads is None:
            if rt is None or self.weight_decay != 0.0:
                return [(0, n, arena.grad)]
            segs = [(0, rt["lo"], arena.grad[:rt["lo"]]), (rt["lo"], rt["hi"], arena.grad[rt["lo"]:rt["hi"]]),
                    (rt["hi"], n, arena.grad[rt["hi"]:])]
            return [sg for sg in segs if sg[1] > sg[0]]
        g = grads[i]
        return [(lo, hi, t) for lo, hi, t in g if hi > lo] if isinstance(g, (list, tuple)) else [(0, n, g)]

    def _is_table(self, arena, lo, hi, t):
        rt = arena.row_table
        return (rt is not None and self.weight_decay == 0.0 and lo == rt["lo"] and hi == rt["hi"] and t.dtype == torch.float32)

    def _norm_into_sumsq(self, grads):
        self._sumsq.zero_()
        for i, g in enumerate(self.groups):
            ar = g["arena"]
            for lo, hi, t in self._segments(ar, grads, i):
                if self._is_table(ar, lo, hi, t):
                    rt = ar.row_table
                    ops.sumsq_rows(t[:rt["V"] * rt["H"]].view(rt["V"], rt["H"]), rt["touched"], self._sumsq, self._partials)
                else:
                    ops.sumsq(t, self._sumsq, self._partials)

    def grad_norm(self, grad_scale=1.0, grads=None):
        """Global L2 norm of (grad * grad_scale) over all arenas -- a device tensor, no sync."""
        self._norm_into_sumsq(grads)
        return torch.sqrt(self._sumsq) * grad_scale

    def step(self, grad_scale=1.0, grads=None):
        """grad_scale multiplies every gradient first (1/accumulation, 1/world after a sum all-reduce).
        grads (optional): per group what to read INSTEAD of arena.grad -- the buffers a data-parallel exchange
        (distributed.GradExchange.reduce) left behind; the clip norm is taken over the same buffers (post-reduce,
        finetune_trainer.py:1010).  Arenas that carry a bf16 shadow get it rewritten by the same launch."""
        self.steps += 1
        self._norm_into_sumsq(grads)
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            ops.clip_coef(self._sumsq, grad_scale, self.max_grad_norm, self._coef)
            coef = self._coef
        else:
            coef = None
        for i, g in enumerate(self.groups):
            ar = g["arena"]
            ns = 0 if ar.shadow is None else ar.shadow.numel()
            for lo, hi, t in self._segments(ar, grads, i):
                if self._is_table(ar, lo, hi, t):         # marked rows only: the rest of the table has g = m = v = 0
                    rt = ar.row_table
                    tv = lambda b: b[lo:lo + rt["V"] * rt["H"]].view(rt["V"], rt["H"])
                    ops.adamw_rows(tv(ar.flat), t[:rt["V"] * rt["H"]].view(rt["V"], rt["H"]), tv(g["m"]), tv(g["v"]), rt["touched"],
                                   g["lr"], self.betas[0], self.betas[1], self.eps, self.weight_decay, self.steps,
                                   gscale_dev=coef, gscale_host=grad_scale)
                    continue
                ops.adamw_step(ar.flat[lo:hi], t, g["m"][lo:hi], g["v"][lo:hi], g["lr"], self.betas[0], self.betas[1], self.eps,
                               self.weight_decay, self.steps, gscale_dev=coef, gscale_host=grad_scale,
                               shadow=ar.shadow[lo:min(hi, ns)] if lo < ns else None)
            if ar.shadow is not None:
                ar.shadow_fresh = True

    def zero_grad(self):
        for g in self.groups:
            g["arena"].zero_grad()

    def set_linear_schedule(self, t_total):
        self.t_total = max(1, int(t_total))

    def scheduler_step(self):
        """get_linear_schedule_with_warmup(optimizer, 0, t_total): lr = base * max(0, (t_total - t) / t_total)."""
        f = max(0.0, (self.t_total - self.steps) / self.t_total)
        for g in self.groups:
            g["lr"] = g["base_lr"] * f


def build_reference_optimizer(tagger, lr=5e-6, lr_rate=10000.0, row_skipping=None, **kw):
    """The reference's two groups for a FastSequenceTagger on TransformerWordEmbeddings (finetune_trainer.py:552-571).
    row_skipping (default: on for a single process, KBNER_ROW_SKIPPING=0 turns it off): the optimizer passes over the
    word-embedding table visit only rows that were ever embedded (ParamArena.enable_row_skipping); data-parallel runs
    switch it on when their gradient exchange sends that table as rows (GradExchange.enable_sparse_rows)."""
    import os
    enc = tagger.embeddings.embeddings[0].model
    enc_arena = enc.ensure_arena()
    if row_skipping is None:
        distributed = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        row_skipping = os.environ.get("KBNER_ROW_SKIPPING", "1") != "0" and not distributed
    if row_skipping and kw.get("weight_decay", 0.0) == 0.0 and enc_arena.row_table is None:
        enc_arena.enable_row_skipping(enc.embeddings.word_embeddings.weight)
    head_arena = ParamArena([tagger.linear.weight, tagger.linear.bias])
    crf_arena = ParamArena([tagger.transitions])
    tagger._head_arena, tagger._crf_arena = head_arena, crf_arena
    return FusedAdamW([{"arena": crf_arena, "lr": lr * lr_rate}, {"arena": enc_arena, "lr": lr},
                       {"arena": head_arena, "lr": lr}], **kw)
