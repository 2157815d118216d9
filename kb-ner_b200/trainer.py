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

What is new: one process per GPU; the only collective is the exchange of the flat gradient arenas on accumulation
boundaries (``distributed.GradExchange``: packed to bf16, NCCL all-reduce), followed by the fused clip + AdamW with
``grad_scale = 1 / world`` reading the reduced buffers.
"""
import logging
import math
import os
import random
import time
from typing import List, Optional

import torch

from .data import BatchedData
from .distributed import GradExchange, allreduce_counts, shard_indices
from .optim import build_reference_optimizer
from .training_utils import Metric, frozen_gc

log = logging.getLogger("kbner_b200")


def make_batches(sentences, mini_batch_size: int, sort: bool = True) -> List[BatchedData]:
    """Sentence-level batching of ColumnDataLoader (custom_data_loader.py:84-149): sort by WORD count (use_bert is
    False for TransformerWordEmbeddings, Appendix B.3), then chunk."""
    from .datasets import ColumnDataLoader
    return list(ColumnDataLoader(list(sentences), mini_batch_size, sentence_level_batch=True, sort_data=sort).data)


class ModelFinetuner:
    def __init__(self, model, teachers=None, corpus=None, optimizer=None, epoch: int = 0, config=None,
                 distill_mode: bool = False, sentence_level_batch: bool = True, is_test: bool = False,
                 professors=None, **kwargs):
        """Positional order (model, teachers, corpus) and the `config=..., **config['ModelFinetuner'], is_test=...`
        keywords are what train.py passes (train.py:127-133; finetune_trainer.py:51-75)."""
        if distill_mode or teachers or professors:
            raise NotImplementedError("knowledge distillation is outside the hot path (config: distill_mode false)")
        self.model = model
        self.corpus = corpus
        self.config = config
        self.epoch = epoch
        self.optimizer = optimizer
        self.sentence_level_batch = sentence_level_batch
        self.is_test = is_test
        self.use_bert = False                 # 'bert' is not in "TransformerWordEmbeddings" (finetune_trainer.py:304-309)
        self.bert_tokenizer = None
        self.embeddings_storage_mode = "none"

    # ------------------------------------------------------------------------------------------------------------
    def train(self, base_path, learning_rate: float = 5e-5, mini_batch_size: int = 32, eval_mini_batch_size: int = None,
              max_epochs: int = 100, gradient_accumulation_steps: int = 1, lr_rate: float = 1.0,
              train_with_dev: bool = False, shuffle: bool = True, true_reshuffle: bool = False,
              save_final_model: bool = True, fine_tune_mode: bool = True, embeddings_storage_mode: str = "none",
              max_grad_norm: float = 5.0, seed: int = 1, log_every: int = 10, select_model_by_macro: bool = False,
              save_finetuned_embedding: bool = False, monitor_test: bool = False, use_warmup: bool = False, **kwargs):
        """Keywords = the YAML `train:` block (Appendix B.12: unknown keys land in **kwargs like in the reference)."""
        if use_warmup:
            raise NotImplementedError("use_warmup: every KB-NER config trains with 0 warm-up steps (finetune_trainer.py:679-688)")
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
        exchange = GradExchange(emb.model, [g["arena"] for g in opt.groups])
        if world > 1:
            # sparse exchange of the word-embedding gradient: capacity = the most sub-tokens any rank embeds within one
            # optimizer step (counted from the batches' own index tensors, which also warms the tokenisation cache)
            toks = [int(emb.build_batch(b)[0].numel()) for b in batches]
            cap = max(sum(sorted(toks, reverse=True)[:gradient_accumulation_steps]), 1) if toks else 1
            t = torch.tensor([cap], dtype=torch.int64, device=emb.device_)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            exchange.enable_sparse_rows(emb.model.ensure_arena(), emb.model.embeddings.word_embeddings.weight, int(t[0]))
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
            gc_guard = frozen_gc().__enter__()          # the corpus is long-lived: keep the per-batch collections off it
            for bi, batch in enumerate(mine):
                batch.features = {}
                tail = len(mine) - (len(mine) // gradient_accumulation_steps) * gradient_accumulation_steps
                denom = tail if (tail and bi >= len(mine) - tail) else gradient_accumulation_steps
                loss = model.forward_loss(batch) / denom
                boundary = (bi + 1) % gradient_accumulation_steps == 0 or bi == len(mine) - 1
                exchange.backward(loss, boundary)        # with overlap on, finished layer chunks are exchanged under the rest
                seen += len(batch)
                if boundary:
                    opt.step(grad_scale=1.0 / world, grads=exchange.reduce())     # None (local fp32 arenas) when world == 1
                    opt.scheduler_step()
                    opt.zero_grad()
                    emb.model.sync_compute_weights_arena()       # free: the optimizer launch rewrote the bf16 shadow
                if (bi + 1) % log_every == 0:
                    run_loss = float(loss.detach()) * denom
                    log.info("epoch %d - iter %d/%d - loss %.6f - samples/sec: %.2f", epoch + 1, bi + 1, len(mine),
                             run_loss, seen / max(time.time() - t0, 1e-9))
                batch.features = {}
            gc_guard.__exit__(None, None, None)
            entry = {"epoch": epoch + 1, "train_samples_per_sec": seen * world / max(time.time() - t0, 1e-9)}
            if not train_with_dev and getattr(self.corpus, "dev", None):
                result, dev_loss = self.evaluate_split(self.corpus.dev, eval_mini_batch_size or mini_batch_size)
                score = result.macro_score if select_model_by_macro else result.main_score     # (:1115-1118)
                entry.update(dev_f1=result.main_score, dev_macro_f1=result.macro_score, dev_loss=dev_loss)
                if score > best and rank == 0:
                    best = score
                    model.save(os.path.join(str(base_path), "best-model.pt"))
                    if save_finetuned_embedding:
                        self.save_finetuned_embeddings(base_path)
            if monitor_test and getattr(self.corpus, "test", None):
                result, test_loss = self.evaluate_split(self.corpus.test, eval_mini_batch_size or mini_batch_size)
                entry.update(test_f1=result.main_score, test_loss=test_loss)
            history.append(entry)
            log.info("EPOCH %d done: %s", epoch + 1, entry)
        if save_final_model and rank == 0:
            model.save(os.path.join(str(base_path), "final-model.pt"))
            if save_finetuned_embedding and (train_with_dev or not getattr(self.corpus, "dev", None)):
                self.save_finetuned_embeddings(base_path)
        return {"history": history, "best_dev_f1": best}

    def save_finetuned_embeddings(self, base_path):
        """tokenizer + encoder of every fine-tuned embedding under base_path/<last path component of its name>
        (finetune_trainer.py:1290-1298; also what `train.py --save_embedding` writes)."""
        for e in self.model.embeddings.embeddings:
            if getattr(e, "fine_tune", False):
                out = os.path.join(str(base_path), str(e.name).rstrip("/").split("/")[-1])
                os.makedirs(out, exist_ok=True)
                e.tokenizer.save_pretrained(out)
                e.model.save_pretrained(out)

    # ------------------------------------------------------------------------------------------------------------
    def evaluate_split(self, sentences, mini_batch_size, out_path=None):
        """model.evaluate over this rank's shard; tp/fp/fn summed over ranks before P/R/F (the only collective of the
        inference path, and only for the metric)."""
        world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        batches = make_batches(list(sentences), mini_batch_size)
        mine = [batches[i] for i in shard_indices(len(batches), rank, world, pad=False)]
        # every rank writes its own shard's predictions; rank 0 stitches them (in rank order) into out_path afterwards
        part = out_path if (world == 1 or out_path is None) else "%s.rank%d" % (out_path, rank)
        result, loss = self.model.evaluate(mine, out_path=part, embeddings_storage_mode="none")
        if world > 1 and out_path is not None:
            torch.distributed.barrier()
            if rank == 0:
                with open(out_path, "w", encoding="utf-8") as out:
                    for r in range(world):
                        with open("%s.rank%d" % (out_path, r), encoding="utf-8") as f:
                            out.write(f.read())
                        os.remove("%s.rank%d" % (out_path, r))
        if world > 1:
            # every rank needs the same class order: tag types come from the shared tag dictionary
            classes = sorted({it.split("-", 1)[1] for it in self.model.tag_dictionary.get_items() if "-" in it})
            vec = allreduce_counts(result.metric.to_vector(classes))
            merged = Metric.from_vector("Evaluation", classes, vec)
            result = merged.to_result()
            result.metric = merged
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
        log.info("final test: %s", result.log_line)
        if not quiet_mode:
            log.info(result.detailed_results)
        # per-corpus scores when several corpora were concatenated (:2216-2282)
        lists = getattr(self.corpus, "test_list", None)
        if overall_test and lists and len(lists) > 1:
            for name, sub in zip(self.corpus.targets, lists):
                if len(sub) == 0:
                    continue
                r, _ = self.evaluate_split(sub, eval_mini_batch_size,
                                           out_path=os.path.join(str(base_path), "%s-test.tsv" % name))
                log.info("%s\t%s\t%s", name, r.log_line, r.macro_score)
        elif quiet_mode:
            print("Average", end=" ")
            print(result.main_score, end=" ")
        return result.main_score


from .datasets import ListCorpus  # noqa: E402,F401  (kept importable from here: the first version lived in this module)
