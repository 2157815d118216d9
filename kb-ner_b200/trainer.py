"""``ModelFinetuner`` -- the fine-tuning loop around the hot path.

Mirror of the step semantics of ``/root/reference/flair/trainers/finetune_trainer.py`` (``ModelFinetuner.train``
:379-1348, ``final_test`` :2136-2282) for the KB-NER configuration; everything that is not arithmetic on the hot path
(TensorBoard, distillation, language attention, pdb fallbacks, ...) is intentionally absent.

What is reproduced (file:line of the reference):
  * two parameter groups by NAME (:552-571): parameters whose name neither contains ``'embedding'`` nor is
    ``linear.weight`` / ``linear.bias`` (= the CRF transitions) at ``learning_rate * lr_rate``, the rest at
    ``learning_rate``; transformers-3.0.0 ``AdamW`` defaults; extra keyword arguments are accepted like the reference's
    ``**kwargs`` pass-through (:436);
  * linear decay to zero over ``ceil(n_batches / accumulation) * max_epochs`` steps, no warm-up (:679-688);
  * per micro-batch: ``loss = model.forward_loss(batch)``; ``loss /= gradient_accumulation_steps`` (tail handled,
    :939-946); ``loss.backward()`` (:956-957); every accumulation boundary: ``clip_grad_norm_(5.0)`` (:1010),
    ``optimizer.step()``, ``zero_grad()``, ``scheduler.step()`` (:1018-1023);
  * shuffling each epoch (:819, ``custom_data_loader.py:74-81``) -- seeded here, so data-parallel ranks agree
    (Appendix B.9), rank r takes batches r::world of the shuffled list;
  * ``samples/sec`` / ``decode_sents/sec`` style throughput logging (:1026-1037), evaluation through
    ``model.evaluate`` (:1111), ``best-model.pt`` / ``final-model.pt`` saving (:1280-1312), ``final_test`` tolerating
    the ``eval_train`` keyword ``train.py --test`` passes (Appendix B.11).

What is new: one process per GPU; the only collective is the all-reduce of the flat gradient arenas on accumulation
boundaries (NCCL), followed by the fused clip + AdamW with ``grad_scale = 1 / world``.
"""
import logging
import math
import os
import random
import time
from typing import List, Optional

import torch

from .data import BatchedData
from .distributed import allreduce_counts, shard_indices
from .optim import build_reference_optimizer

log = logging.getLogger("kbner_b200")


def make_batches(sentences, mini_batch_size: int, sort: bool = True) -> List[BatchedData]:
    """Sentence-level batching of ColumnDataLoader (custom_data_loader.py:84-149): sort by WORD count (use_bert is
    False for TransformerWordEmbeddings, Appendix B.3), then chunk."""
    from .datasets import ColumnDataLoader
    return list(ColumnDataLoader(list(sentences), mini_batch_size, sentence_level_batch=True, sort_data=sort).data)


class ModelFinetuner:
    def __init__(self, model, teachers=None, corpus=None, optimizer=None, epoch: int = 0, config=None,
                 distill_mode: bool = False, sentence_level_batch: bool = True, **kwargs):
        if distill_mode or teachers:
            raise NotImplementedError("knowledge distillation is outside the hot path (config: distill_mode false)")
        self.model = model
        self.corpus = corpus
        self.config = config
        self.epoch = epoch
        self.optimizer = optimizer
        self.sentence_level_batch = sentence_level_batch

    # ------------------------------------------------------------------------------------------------------------
    def train(self, base_path, learning_rate: float = 5e-5, mini_batch_size: int = 32, eval_mini_batch_size: int = None,
              max_epochs: int = 100, gradient_accumulation_steps: int = 1, lr_rate: float = 1.0,
              train_with_dev: bool = False, shuffle: bool = True, true_reshuffle: bool = False,
              save_final_model: bool = True, fine_tune_mode: bool = True, embeddings_storage_mode: str = "none",
              max_grad_norm: float = 5.0, seed: int = 1, log_every: int = 10, **kwargs):
        os.makedirs(str(base_path), exist_ok=True)
        world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        model = self.model
        emb = model.embeddings.embeddings[0]
        if fine_tune_mode:
            emb.fine_tune, emb.static_embeddings = True, False
        train_sents = list(self.corpus.train)
        if train_with_dev and getattr(self.corpus, "dev", None):
            train_sents += list(self.corpus.dev)
        batches = make_batches(train_sents, mini_batch_size)
        steps_per_epoch = math.ceil(math.ceil(len(batches) / world) / gradient_accumulation_steps)
        opt = self.optimizer or build_reference_optimizer(model, lr=learning_rate, lr_rate=lr_rate,
                                                          max_grad_norm=max_grad_norm)
        opt.set_linear_schedule(steps_per_epoch * max_epochs)
        arenas = [g["arena"] for g in opt.groups]
        rnd = random.Random(seed)
        best, history = -1.0, []
        for epoch in range(self.epoch, max_epochs):
            if shuffle:
                rnd.shuffle(batches)                     # same permutation on every rank (seeded)
            mine = [batches[i] for i in shard_indices(len(batches), rank, world)]
            model.train()
            emb.train()
            opt.zero_grad()
            seen, t0, run_loss = 0, time.time(), 0.0
            for bi, batch in enumerate(mine):
                batch.features = {}
                tail = len(mine) - (len(mine) // gradient_accumulation_steps) * gradient_accumulation_steps
                denom = tail if (tail and bi >= len(mine) - tail) else gradient_accumulation_steps
                loss = model.forward_loss(batch) / denom
                loss.backward()
                seen += len(batch)
                if (bi + 1) % gradient_accumulation_steps == 0 or bi == len(mine) - 1:
                    if world > 1:
                        for ar in arenas:
                            torch.distributed.all_reduce(ar.grad)
                    opt.step(grad_scale=1.0 / world)
                    opt.scheduler_step()
                    opt.zero_grad()
                    emb.model.sync_compute_weights_arena()
                if (bi + 1) % log_every == 0:
                    run_loss = float(loss.detach()) * denom
                    log.info("epoch %d - iter %d/%d - loss %.6f - samples/sec: %.2f", epoch + 1, bi + 1, len(mine),
                             run_loss, seen / max(time.time() - t0, 1e-9))
                batch.features = {}
            entry = {"epoch": epoch + 1, "train_samples_per_sec": seen * world / max(time.time() - t0, 1e-9)}
            if not train_with_dev and getattr(self.corpus, "dev", None):
                result, dev_loss = self.evaluate_split(self.corpus.dev, eval_mini_batch_size or mini_batch_size)
                entry.update(dev_f1=result["main_score"], dev_loss=dev_loss)
                if result["main_score"] > best and rank == 0:
                    best = result["main_score"]
                    model.save(os.path.join(str(base_path), "best-model.pt"))
            history.append(entry)
            log.info("EPOCH %d done: %s", epoch + 1, entry)
        if save_final_model and rank == 0:
            model.save(os.path.join(str(base_path), "final-model.pt"))
        return {"history": history, "best_dev_f1": best}

    # ------------------------------------------------------------------------------------------------------------
    def evaluate_split(self, sentences, mini_batch_size, out_path=None):
        """model.evaluate over this rank's shard; tp/fp/fn summed over ranks before P/R/F (the only collective of the
        inference path, and only for the metric)."""
        world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        batches = make_batches(list(sentences), mini_batch_size)
        mine = [batches[i] for i in shard_indices(len(batches), rank, world, pad=False)]
        result, loss = self.model.evaluate(mine, out_path=out_path, embeddings_storage_mode="none")
        if world > 1:
            tp, fp, fn = allreduce_counts([result["tp"], result["fp"], result["fn"]])
            p = round(tp / (tp + fp), 4) if tp + fp else 0.0
            r = round(tp / (tp + fn), 4) if tp + fn else 0.0
            result.update(tp=tp, fp=fp, fn=fn, precision=p, recall=r,
                          main_score=round(2 * p * r / (p + r), 4) if p + r else 0.0)
        return result, loss

    def final_test(self, base_path, eval_mini_batch_size: int = 32, overall_test: bool = True, quiet_mode: bool = False,
                   nocrf: bool = False, predict_posterior: bool = False, keep_embedding: int = -1, sort_data: bool = False,
                   eval_train: bool = False, **kwargs):
        """final_test (:2136-2282).  `eval_train` is accepted (and ignored) because train.py --test passes it."""
        path = os.path.join(str(base_path), "best-model.pt")
        if not os.path.exists(path):
            path = os.path.join(str(base_path), "final-model.pt")
        if os.path.exists(path):
            self.model = type(self.model).load(path)
        for e in self.model.embeddings.embeddings:
            e.fine_tune, e.static_embeddings = False, True           # (:2167-2169)
        self.model.eval()
        result, loss = self.evaluate_split(self.corpus.test, eval_mini_batch_size,
                                           out_path=os.path.join(str(base_path), "test.tsv"))
        log.info("final test: %s", result.get("log_line"))
        return result["main_score"]


class ListCorpus:
    """train / dev / test lists of sentences (flair/list_data.py:2-19, reduced to what the trainer reads)."""

    def __init__(self, train, dev=None, test=None):
        self.train, self.dev, self.test = list(train), list(dev or []), list(test or [])
