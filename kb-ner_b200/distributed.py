"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed for the rendezvous.

The path shards by sentence: inference / Viterbi need no collective at all; fine-tuning needs one
exchange per optimizer step -- an all-reduce of gradients only.  The reference is single-process
(`flair/trainers/finetune_trainer.py:699-700` has only a commented-out DataParallel); the semantics reproduced
here are those of its step (:939-957, :1007-1023): loss is a per-batch mean (sequence_tagger_model.py:2506), so
averaging rank gradients equals a global batch of world*B; clipping uses the post-reduce norm.
"""
from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world: int, pad: bool = True) -> List[int]:
    """Indices rank `rank` processes out of n items (round robin r::world).  With pad=True the tail is padded by
    wrapping so every rank runs the same number of steps (needed when a collective follows each step)."""
    idx = list(range(rank, n, world))
    if pad and n > 0:
        per = (n + world - 1) // world
        k = 0
        while len(idx) < per:
            idx.append((rank + k * world) % n)
            k += 1
    return idx


def allreduce_counts(counts: Iterable[int], device=None) -> List[int]:
    """Sum small integer vectors (tp/fp/fn of span-F1, sentence counts) over ranks; identity when not distributed."""
    counts = list(counts)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts
    t = torch.tensor(counts, dtype=torch.int64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


class GradBucketReducer:
    """Bucketed gradient all-reduce (sum, then / world) over flat fp32 buckets.

    Parameters are packed in registration order into buckets of ~bucket_mb; `reduce()` is called on the last
    gradient-accumulation micro-step only (the reference steps every `gradient_accumulation_steps` micro-batches,
    finetune_trainer.py:1007-1023).  With NCCL each bucket's all-reduce is launched asynchronously so it overlaps the
    copy-in of the next bucket; `reduce()` returns after all buckets are written back."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 64.0):
        self.params = [p for p in params if p.requires_grad]
        cap = int(bucket_mb * (1 << 20) / 4)
        self.buckets, cur, size = [], [], 0
        for p in self.params:
            if cur and size + p.numel() > cap:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            self.buckets.append(cur)

    @torch.no_grad()
    def reduce(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        work = []
        for bucket in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1).float() for g in grads])
            h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            work.append((h, flat, bucket))
        for h, flat, bucket in work:
            h.wait()
            flat.div_(world)
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(flat[off:off + n].view_as(p))
                off += n


def global_grad_norm(params: Iterable[torch.nn.Parameter]) -> torch.Tensor:
    """L2 norm over all gradients (identical on every rank after GradBucketReducer.reduce)."""
    sq = [p.grad.float().pow(2).sum() for p in params if p.grad is not None]
    return torch.sqrt(torch.stack(sq).sum()) if sq else torch.zeros(())


class OverlappedGradAllReduce:
    """All-reduce of the flat gradient arenas that overlaps the backward pass of the LAST accumulation micro-step.

    `with OverlappedGradAllReduce(encoder, arenas) as sync: loss.backward()` makes the encoder run its backward in layer
    chunks (encoder._backward_chunked) and starts the all-reduce of every finalised arena slice asynchronously while the
    next chunk computes; on exit the remaining arenas (tag projection, transitions) are reduced and every handle is
    waited for.  Sum only -- the caller divides by the world size (FusedAdamW.step(grad_scale=1/world)).  Without an
    initialised process group (or world size 1) it changes nothing.  Enabled with KBNER_OVERLAP_ALLREDUCE=1."""

    def __init__(self, encoder, arenas):
        self.encoder, self.arenas, self.handles = encoder, list(arenas), []
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    @staticmethod
    def enabled():
        import os
        return os.environ.get("KBNER_OVERLAP_ALLREDUCE", "0") == "1"

    def _slice_done(self, lo, hi):
        if hi > lo:
            self.handles.append(dist.all_reduce(self.encoder.arena.grad[lo:hi], async_op=True))

    def __enter__(self):
        if self.active:
            self.encoder._grad_sync = self._slice_done
        return self

    def __exit__(self, exc_type, exc, tb):
        if not self.active:
            return False
        self.encoder._grad_sync = None
        if exc_type is None:
            for ar in self.arenas:
                if ar is not self.encoder.arena:
                    self.handles.append(dist.all_reduce(ar.grad, async_op=True))
            for h in self.handles:
                h.wait()
        self.handles = []
        return False
