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


class GradExchange:
    """The one collective of the path: the per-optimizer-step gradient exchange of data-parallel fine-tuning.

    The reference steps every `gradient_accumulation_steps` micro-batches (finetune_trainer.py:1007-1023); here, on that
    boundary, every rank's flat fp32 gradient arenas are summed over ranks (the caller's FusedAdamW.step(grad_scale =
    1/world, grads=...) turns the sum into the mean and clips on the post-reduce norm).

    payload  "bf16" (default): each arena is packed fp32 -> bf16 by one kernel into a static buffer, NCCL all-reduces the
             bf16 buffer (half the 2.24 GB of XLM-R-large's fp32 gradient) and AdamW reads the gradient straight from
             it -- the fp32 arena is never written back.  "fp32" (KBNER_GRAD_COMM=fp32) all-reduces the arenas in place.
    sparse   (enable_sparse_rows; bf16 payload) the word-embedding gradient -- 1.02 GB of the 2.24, with at most
             tokens-per-optimizer-step touched rows -- is not all-reduced: every rank all-gathers the ids it touched and
             those rows (bf16), clears them locally and adds all ranks' rows back in RANK ORDER (the same sum on every
             rank: replicas stay bit-identical); AdamW reads that region from the fp32 arena.
    overlap  (default on; KBNER_OVERLAP_ALLREDUCE=0 turns it off) the last backward of the accumulation cycle runs in
             layer chunks (encoder._backward_chunked) and the pack + all-reduce of every finalised arena slice is started
             asynchronously under the next chunk.  KBNER_OVERLAP_SM_CARVEOUT=n sizes the persistent GEMM / attention
             grids for num_sms - n SMs meanwhile (kbner_set_sm_budget) so that NCCL's channel CTAs have SMs of their own;
             measured, that costs more than it gains (8 GPUs: 4825 sentences/s with n = 16, 4839 with 0, 4800 without
             overlap), so the default is 0.
    Without an initialised process group (or world size 1) nothing is exchanged and reduce() returns None.
    `kernels` (pack / rows_gather / rows_scatter_add / mark_rows) are injectable so the host logic can be exercised with gloo on CPU
    (tests/test_distributed_cpu.py); the product binds the CUDA ones."""

    def __init__(self, encoder, arenas, payload=None, overlap=None, pack=None, kernels=None):
        import os
        self.encoder, self.arenas = encoder, list(arenas)
        self.payload = payload or os.environ.get("KBNER_GRAD_COMM", "bf16")
        if self.payload not in ("bf16", "fp32"):
            raise ValueError("gradient payload must be 'bf16' or 'fp32'")
        self.overlap = (os.environ.get("KBNER_OVERLAP_ALLREDUCE", "1") == "1") if overlap is None else bool(overlap)
        self.carveout = int(os.environ.get("KBNER_OVERLAP_SM_CARVEOUT", "0"))
        self.sparse_wanted = os.environ.get("KBNER_SPARSE_EMB_GRAD", "1") != "0"
        k = dict(kernels or {})
        if pack is not None:
            k["pack"] = pack
        if not all(n in k for n in ("pack", "rows_gather", "rows_scatter_add", "mark_rows")):
            from . import ops
            k.setdefault("pack", ops.pack_bf16)
            k.setdefault("rows_gather", ops.rows_gather_bf16)
            k.setdefault("rows_scatter_add", ops.rows_scatter_add_bf16)
            k.setdefault("mark_rows", ops.mark_rows)
        self._k = k
        self._bufs = {}
        self._handles = []
        self._encoder_done = False
        self._sparse = None
        self.bytes_per_step = 0

    @property
    def active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    # ---- sparse rows -----------------------------------------------------------------------------------------------
    def enable_sparse_rows(self, arena, table, cap_tokens):
        """Exchange `table`'s gradient ([V,H] parameter living in `arena`) as touched rows.  cap_tokens: the largest number of
        sub-tokens ANY rank embeds within one optimizer step (the same value on every rank: buffers are all-gathered)."""
        if not (self.active and self.sparse_wanted and self.payload == "bf16"):
            return False
        world = dist.get_world_size()
        V, H = table.shape
        lo = arena.offsets[id(table)]
        dev = arena.grad.device
        cap = (int(cap_tokens) + 7) // 8 * 8
        if world * cap >= V:                      # nothing to gain: the union could be the whole table
            return False
        self._sparse = dict(arena=arena, lo=lo, hi=lo + (V * H + 7) // 8 * 8, V=V, H=H, cap=cap, n=0, done=False,
                            ids=torch.full((cap,), -1, dtype=torch.int32, device=dev),
                            rows=torch.empty((cap, H), dtype=torch.bfloat16, device=dev),
                            all_ids=torch.empty((world, cap), dtype=torch.int32, device=dev),
                            all_rows=torch.empty((world, cap, H), dtype=torch.bfloat16, device=dev))
        if self.encoder is not None:
            self.encoder._ids_hook = self.note_ids
        # every gradient that reaches the table now arrives with its ids: the optimizer may skip never-touched rows
        import os
        if os.environ.get("KBNER_ROW_SKIPPING", "1") != "0" and arena.row_table is None and hasattr(arena, "enable_row_skipping"):
            arena.enable_row_skipping(table)
        return True

    def note_ids(self, ids):
        """Called by the encoder's training forward with the [R,S] id tensor of every micro-batch."""
        sp = self._sparse
        if sp is None:
            return
        n = ids.numel()
        if sp["n"] + n > sp["cap"]:
            raise RuntimeError("GradExchange: %d sub-tokens in one optimizer step exceed the agreed capacity %d"
                               % (sp["n"] + n, sp["cap"]))
        sp["ids"][sp["n"]:sp["n"] + n].copy_(ids.reshape(-1), non_blocking=True)
        sp["n"] += n

    def _launch_sparse(self):
        sp = self._sparse
        if sp["done"]:
            return
        sp["done"] = True
        ids, _ = torch.sort(sp["ids"])                                  # -1 (unused slots) first, duplicates adjacent
        dup = torch.zeros_like(ids, dtype=torch.bool)
        dup[1:] = ids[1:] == ids[:-1]
        ids = torch.where(dup, torch.full_like(ids, -1), ids)           # keep the first of every run (no host sync)
        table_grad = sp["arena"].grad[sp["lo"]:sp["lo"] + sp["V"] * sp["H"]].view(sp["V"], sp["H"])
        self._k["rows_gather"](table_grad, ids, sp["rows"], True)       # touched rows -> bf16, cleared in the table
        sp["sorted"] = ids
        world = dist.get_world_size()
        if dist.get_backend() == "nccl":
            self._handles.append(dist.all_gather_into_tensor(sp["all_ids"].view(-1), ids, async_op=True))
            self._handles.append(dist.all_gather_into_tensor(sp["all_rows"].view(-1, sp["H"]), sp["rows"], async_op=True))
        else:
            self._handles.append(dist.all_gather(list(sp["all_ids"].unbind(0)), ids, async_op=True))
            self._handles.append(dist.all_gather(list(sp["all_rows"].unbind(0)), sp["rows"], async_op=True))
        self.bytes_per_step += world * (ids.numel() * 4 + sp["rows"].numel() * 2)

    def _finish_sparse(self):
        sp = self._sparse
        table_grad = sp["arena"].grad[sp["lo"]:sp["lo"] + sp["V"] * sp["H"]].view(sp["V"], sp["H"])
        for r in range(dist.get_world_size()):                          # rank order: the same sum on every rank
            self._k["rows_scatter_add"](sp["all_rows"][r], sp["all_ids"][r], table_grad)
        rt = getattr(sp["arena"], "row_table", None)
        if rt is not None:                                              # rows other ranks touched carry gradient here too
            self._k["mark_rows"](sp["all_ids"].view(-1), rt["touched"])
        sp["ids"].fill_(-1)
        sp["n"], sp["done"] = 0, False

    # ---- dense slices ----------------------------------------------------------------------------------------------
    def _buf(self, ar):
        b = self._bufs.get(id(ar))
        if b is None:
            b = self._bufs[id(ar)] = torch.empty(ar.grad.numel(), dtype=torch.bfloat16, device=ar.grad.device)
        return b

    def _launch(self, ar, lo, hi):
        if hi <= lo:
            return
        sp = self._sparse
        if sp is not None and ar is sp["arena"] and lo < sp["hi"] and hi > sp["lo"]:
            self._launch(ar, lo, min(hi, sp["lo"]))                    # dense part below the table
            self._launch_sparse()
            self._launch(ar, max(lo, sp["hi"]), hi)                    # dense part above it
            return
        if self.payload == "bf16":
            dst = self._buf(ar)[lo:hi]
            self._k["pack"](ar.grad[lo:hi], dst)
        else:
            dst = ar.grad[lo:hi]
        self.bytes_per_step += dst.numel() * dst.element_size()
        self._handles.append(dist.all_reduce(dst, op=dist.ReduceOp.SUM, async_op=True))

    def _set_budget(self, on):
        if self.carveout > 0 and self.arenas and self.arenas[0].grad.is_cuda:
            from . import _lib
            lib = _lib.load()
            sms = torch.cuda.get_device_properties(self.arenas[0].grad.device).multi_processor_count
            _lib.check(lib.kbner_set_sm_budget(max(2, sms - self.carveout) if on else 0), "set_sm_budget")

    def _slice_done(self, lo, hi):
        self._launch(self.encoder.arena, lo, hi)

    def backward(self, loss, boundary):
        """loss.backward(); on an accumulation boundary with overlap on, the encoder reports finalised arena slices."""
        if boundary and self.active and self.overlap and self.encoder is not None:
            self.bytes_per_step = 0
            self.encoder._grad_sync = self._slice_done
            self._set_budget(True)
            try:
                loss.backward()
            finally:
                self.encoder._grad_sync = None
                self._set_budget(False)
            self._encoder_done = True
        else:
            loss.backward()

    def reduce(self):
        """Exchange whatever has not been started yet, wait for everything, return per arena what the optimizer must read:
        a flat buffer (bf16 payload: the packed buffer; fp32: the arena's own .grad) or -- for the arena whose embedding
        table went the sparse way -- a list of (lo, hi, buffer) segments; None when not distributed."""
        if not self.active:
            return None
        if not self._encoder_done:
            self.bytes_per_step = 0
        for ar in self.arenas:
            if self._encoder_done and self.encoder is not None and ar is getattr(self.encoder, "arena", None):
                continue
            self._launch(ar, 0, ar.grad.numel())
        for h in self._handles:
            h.wait()
        self._handles, self._encoder_done = [], False
        sp = self._sparse
        if sp is not None:
            self._finish_sparse()
        out = []
        for ar in self.arenas:
            if self.payload != "bf16":
                out.append(ar.grad)
            elif sp is not None and ar is sp["arena"]:
                b, n = self._buf(ar), ar.grad.numel()
                out.append([(0, sp["lo"], b[:sp["lo"]]), (sp["lo"], sp["hi"], ar.grad[sp["lo"]:sp["hi"]]), (sp["hi"], n, b[sp["hi"]:n])])
            else:
                out.append(self._buf(ar))
        return out
