"""``SequenceTagger`` / ``FastSequenceTagger`` on the B200 kernels.

Mirror of ``/root/reference/flair/models/sequence_tagger_model.py`` for the KB-NER configuration
(``use_crf=True, use_rnn=False``): same constructor keywords (:100-163), same methods and attribute names
(SURVEY.md 8(b)), same arithmetic (Appendix A) -- executed as

    forward            :844-1052   first-sub-token gather + WordDropout + Linear   -> 1 kernel
    _calculate_loss    :2426-2539  remove-X compaction + log Z + gold score         -> 2 kernels (+1 backward)
    _obtain_labels     :1157-1246  per-sentence Viterbi + S-X padding               -> 1 kernel for the batch
    _viterbi_decode    :1248-1327  kept for callers that decode one sentence

instead of Python loops over sentences and tokens with a ``.item()`` sync per token (:1296-1300).
"""
import logging
import time
from typing import List, Optional, Union

import torch

from . import ops
from .data import BatchedData, Dictionary, Label, LabelSeq, Sentence
from .training_utils import Metric, Result, frozen_gc, span_counts, store_embeddings

log = logging.getLogger("kbner_b200")

START_TAG = "<START>"
STOP_TAG = "<STOP>"


class _CrfNll(torch.autograd.Function):
    """sum_b w_b (logZ_b - gold_b) with the hand-written backward kernel."""

    @staticmethod
    def forward(ctx, emis, trans, tags, pos, klen, start, stop):
        emis = emis.contiguous()
        trans_c = trans.contiguous()
        logz, gold, alpha = ops.crf_nll_fwd(emis, tags, trans_c, klen, start, stop, pos=pos, want_alpha=True)
        ctx.save_for_backward(emis, trans_c, tags, pos, klen, alpha[0], alpha[1])
        ctx.start, ctx.stop = start, stop
        return logz - gold

    @staticmethod
    def backward(ctx, grad_out):
        emis, trans, tags, pos, klen, ahat, scale = ctx.saved_tensors
        d_emis, d_trans = ops.crf_nll_bwd(emis, tags, trans, klen, (ahat, scale), grad_out.contiguous().float(),
                                          ctx.start, ctx.stop, pos=pos)
        return d_emis, d_trans, None, None, None, None, None


class SequenceTagger(torch.nn.Module):
    def __init__(self, hidden_size: int, embeddings, tag_dictionary: Dictionary, tag_type: str,
                 use_crf: bool = True, use_mfvi: bool = False, use_rnn: bool = True, use_cnn: bool = False,
                 rnn_layers: int = 1, dropout: float = 0.0, word_dropout: float = 0.05, locked_dropout: float = 0.5,
                 train_initial_hidden_state: bool = False, pickle_module: str = "pickle", interpolation: float = 0.5,
                 sentence_loss: bool = False, distill_crf: bool = False, crf_attention: bool = False,
                 biaf_attention: bool = False, use_language_attention: bool = False,
                 token_level_attention: bool = False, target_languages: int = 1, config=None, word_map=None,
                 char_map=None, use_decoder_timer: bool = True, debug: bool = False, temperature: float = 1,
                 testing: bool = False, remove_x: bool = False, multi_view_training: bool = False, **kwargs):
        super().__init__()
        if not use_crf or use_rnn or use_cnn or use_mfvi:
            raise NotImplementedError("kbner_b200 implements the KB-NER tagger head: use_crf=True, use_rnn=False, "
                                      "use_cnn=False (config/*.yaml `model:` block)")
        off = dict(distill_crf=distill_crf, crf_attention=crf_attention, biaf_attention=biaf_attention,
                   use_language_attention=use_language_attention, token_level_attention=token_level_attention,
                   multi_view_training=multi_view_training)
        for k in ("enhanced_crf", "posterior_constraint", "predict_posterior", "distill_posterior", "use_language_vector",
                  "use_transition_attention", "unlabel_entropy_loss", "relearn_embeddings", "map_embeddings",
                  "embedding_selector", "use_rl", "use_gumbel", "use_embedding_masks", "embedding_attention"):
            off[k] = kwargs.get(k, False)
        bad = [k for k, v in off.items() if v]
        if bad:
            raise NotImplementedError("options outside the hot path (KD / multi-view / ACE variants): %s" % bad)
        # dropout / locked_dropout: the reference applies Dropout and LockedDropout to the embedded sentence tensor BEFORE
        # the use_rnn branch (sequence_tagger_model.py:958-964), i.e. also on this head, in train() mode.  Every KB-NER YAML
        # sets both to 0.0 and the fused projection kernel carries word dropout only: the values are stored (they are part
        # of the checkpoint format, :452-454) and a TRAINING forward with either > 0 raises -- see forward().
        # ---- attribute surface read by ModelFinetuner / train.py -----------------------
        self.debug, self.use_language_attention, self.biaf_attention = debug, False, False
        self.token_level_attention, self.use_language_vector, self.use_crf = False, False, True
        self.use_decoder_timer, self.sentence_level_loss = use_decoder_timer, sentence_loss
        self.temperature = temperature
        self.use_rnn, self.use_cnn, self.use_mfvi, self.use_bert = False, False, False, False
        self.hidden_size, self.rnn_layers = hidden_size, rnn_layers
        self.embeddings = embeddings
        self.config, self.word_map, self.char_map = config, word_map, char_map
        self.lemma_map = self.postag_map = None
        self.tag_dictionary: Dictionary = tag_dictionary
        self.tag_type: str = tag_type
        self.tagset_size: int = len(tag_dictionary)
        if self.tagset_size > 32:
            raise NotImplementedError("CRF kernels map one tag per warp lane: tagset must be <= 32 (got %d); every "
                                      "shipped KB-NER dictionary has <= 29 tags" % self.tagset_size)
        self.remove_x = remove_x
        self.target_languages = target_languages
        self.multi_view_training = False
        self.distill_crf = self.distill_posterior = self.distill_prob = self.distill_exact = False
        self.distill_emission = self.crf_attention = self.enhanced_crf = self.predict_posterior = False
        self.posterior_constraint = self.use_transition_attention = self.unlabel_entropy_loss = False
        self.relearn_embeddings = self.map_embeddings = self.embedding_selector = self.use_rl = False
        self.use_dropout, self.use_word_dropout, self.use_locked_dropout = dropout, word_dropout, locked_dropout
        self.pickle_module = pickle_module
        self.interpolation = interpolation
        self.time = 0.0
        self.mask = None
        self._keep = None
        # ---- parameters: names `linear.*` / `transitions` drive the trainer's LR groups (:552-553) ---
        self.linear = torch.nn.Linear(self.embeddings.embedding_length, self.tagset_size)
        self.start_idx = tag_dictionary.get_idx_for_item(START_TAG)
        self.stop_idx = tag_dictionary.get_idx_for_item(STOP_TAG)
        trans = torch.randn(self.tagset_size, self.tagset_size)
        trans[self.start_idx, :] = -1e12          # nothing transitions INTO <START>   (:402-410; [to, from])
        trans[:, self.stop_idx] = -1e12           # nothing transitions FROM <STOP>
        self.transitions = torch.nn.Parameter(trans)
        if not testing:
            dev = getattr(getattr(embeddings, "embeddings", [None])[0], "device_", None) or "cuda"
            self.to(dev)

    # ---- helpers -----------------------------------------------------------------------------------
    @property
    def device(self):
        return self.transitions.device

    @property
    def x_idx(self):
        """Index of 'S-X', looked up at call time like the reference (:1202, :2449): `train.py --parse --remove_x` adds the
        item to the dictionary AFTER the model was loaded.  Items beyond the L emission columns cannot be decoded."""
        idx = self.tag_dictionary.get_idx_for_item("S-X")
        if idx >= self.tagset_size:
            raise ValueError("'S-X' was added to the tag dictionary after the model was built (index %d, %d emission columns): "
                             "remove_x needs a model trained with S-X in its tag set" % (idx, self.tagset_size))
        return idx

    def _encoded(self, sentences):
        """The device-side batch left behind by embeddings.embed (one stacked embedding)."""
        return sentences.features[self.embeddings.embeddings[0].name]

    @staticmethod
    def sequence_mask(lengths, max_len=None):
        lengths = torch.as_tensor(lengths)
        max_len = int(max_len or lengths.max())
        return torch.arange(max_len, device=lengths.device)[None, :] < lengths[:, None]

    # ---- forward (:844-1052) -------------------------------------------------------------------------
    def forward(self, sentences, prediction_mode: bool = False):
        if not isinstance(sentences, BatchedData):
            sentences = BatchedData(sentences if isinstance(sentences, list) else [sentences])
        self._batch = sentences
        self.embeddings.embed(sentences)
        enc = self._encoded(sentences)
        lengths = enc.lengths
        T = max(lengths)
        if self.use_decoder_timer:
            self.time = time.time()
        drop_keep = None
        if self.training and (self.use_dropout > 0.0 or self.use_locked_dropout > 0.0):
            raise NotImplementedError("training with dropout=%s / locked_dropout=%s on the tagger head is not built (the KB-NER "
                                      "configs set both to 0.0; inference with such a checkpoint is unaffected)"
                                      % (self.use_dropout, self.use_locked_dropout))
        if self.training and self.use_word_dropout > 0.0:
            # WordDropout on [T,B,D]: one Bernoulli(1-p) draw per time step, shared by the batch, no rescale
            # (flair/nn.py:176-183)
            drop_keep = torch.empty(T, device=self.device).bernoulli_(1.0 - self.use_word_dropout).to(torch.uint8)
        features = ops.gather_tagproj_fwd(enc.hidden, enc.row_of, enc.first_idx, self.linear.weight.float().contiguous(),
                                          self.linear.bias.float().contiguous(), enc.S, drop_keep=drop_keep)
        if self.training and (self.linear.weight.requires_grad or self.linear.bias.requires_grad):
            features = _TagProjGrad.apply(features, self.linear.weight, self.linear.bias, enc, drop_keep)
        # word counts ride in the batch's single pinned H2D copy (a torch.tensor(..., device=cuda) here is a
        # blocking pageable copy that waits for the GPU: it serialised host and device, 3.7 ms per batch)
        self.lengths_t = enc.lengths_d if getattr(enc, "lengths_d", None) is not None else \
            torch.tensor(lengths, dtype=torch.int32, device=self.device)
        self.mask = self.sequence_mask(self.lengths_t, T).to(features.dtype)        # (:1028)
        self._keep = None
        return features

    # ---- loss (:1899-1921, :2426-2539) ------------------------------------------------------------------
    def forward_loss(self, data_points: Union[List[Sentence], Sentence], sort=True, return_features=False):
        features = self.forward(data_points)
        loss = self._calculate_loss(features, self._batch, self.mask)
        return (loss, features) if return_features else loss

    def _gold_tags(self, sentences, T):
        rows = []
        for s in sentences:
            tg = getattr(s, self.tag_type + "_tags", None)
            if tg is None:
                tg = torch.tensor([self.tag_dictionary.get_idx_for_item(tok.get_tag(self.tag_type).value)
                                   for tok in s.tokens], dtype=torch.int32)
            tg = torch.as_tensor(tg).to(torch.int32)
            if tg.numel() < T:
                tg = torch.cat([tg, torch.zeros(T - tg.numel(), dtype=torch.int32)])     # pad with 0 = <unk>
            rows.append(tg[:T])
        tags = torch.stack(rows, 0)
        # host-side range check (the tensor is still on the host): a gold tag outside the L emission columns -- e.g. an item
        # added to the dictionary after the model was built -- would index past emis / trans in the CRF kernels
        if tags.numel() and (int(tags.max()) >= self.tagset_size or int(tags.min()) < 0):
            raise ValueError("gold tag index outside [0, %d): the tag dictionary has items the model was not built with"
                             % self.tagset_size)
        return tags.to(self.device, non_blocking=True)

    def _calculate_loss(self, features: torch.Tensor, sentences, mask: torch.Tensor):
        B, T, L = features.shape
        tags = self._gold_tags(sentences, T)
        if not self.remove_x and mask is self.mask and getattr(self, "lengths_t", None) is not None \
                and self.lengths_t.numel() == B:
            # the mask forward() built is the prefix mask of the word counts: no compaction, no index list -- the CRF
            # kernels stream the emission rows with 16-byte copies
            pos, klen = None, self.lengths_t
        else:
            keep = mask.bool()
            if self.remove_x:
                keep = keep & (tags != self.x_idx)               # (:2448-2453)
                self.mask = keep.to(features.dtype)
            keep_u8 = keep.to(torch.uint8).contiguous()
            pos, klen = ops.crf_compact(keep_u8)                  # (:2474-2488)
        self._keep = (pos, klen)
        nll = _CrfNll.apply(features, self.transitions, tags.contiguous(), pos, klen, self.start_idx, self.stop_idx)
        # is_unlabel is a host-side attribute: decided here without touching the device (two device->host syncs per
        # micro-step sat between the forward and backward graphs in the first version)
        n_lab = sum(0 if getattr(s, "is_unlabel", False) else 1 for s in sentences)
        if n_lab == 0:
            return nll.sum() * 0.0                                 # (:2500-2501)
        if n_lab < len(sentences):
            return nll.sum() / n_lab                               # (:2502-2504) the UNMASKED sum over the labelled count
        return nll.mean()                                          # (:2506)

    # ---- decode (:1157-1246) ---------------------------------------------------------------------------
    def _decode_batch(self, feature: torch.Tensor):
        B, T, L = feature.shape
        slen = self.lengths_t
        if self.remove_x and self._keep is not None:
            pos, klen = self._keep                                 # the keep-mask side channel (Appendix B.8)
        else:
            pos, klen = None, slen
        return ops.crf_viterbi(feature.detach().float().contiguous(), self.transitions.detach().contiguous(), klen,
                               slen, self.start_idx, self.stop_idx, self.x_idx, pos=pos)

    def _obtain_labels(self, feature, sentences, get_all_tags: bool = False):
        if get_all_tags:
            raise NotImplementedError("get_all_tags (per-tag score dump, :1306-1325) is outside the hot path")
        return self._labels_from_handle(self._decode_async(feature), sentences), []

    def _decode_async(self, feature):
        """Launch the batched Viterbi and the device->host copies of its result (pinned, non-blocking); returns a
        handle `_labels_from_handle` turns into Label lists.  Lets `evaluate` overlap the host-side Label
        construction of batch i with the kernels of batch i+1."""
        tags, conf = self._decode_batch(feature)
        B, T = tags.shape
        pool = getattr(self, "_d2h_pool", None)
        if pool is None:
            pool = self._d2h_pool = []
        slot = None
        for s in pool:
            if not s["busy"] and s["tags"].numel() >= B * T:
                slot = s
                break
        if slot is None:
            n = max(B * T, 1 << 15)
            slot = {"tags": torch.empty(n, dtype=torch.int32).pin_memory(),
                    "conf": torch.empty(n, dtype=torch.float32).pin_memory(), "busy": False}
            pool.append(slot)
        slot["busy"] = True
        slot["tags"][:B * T].copy_(tags.view(-1), non_blocking=True)
        slot["conf"][:B * T].copy_(conf.view(-1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return (slot, ev, B, T)

    def _labels_from_handle(self, handle, sentences):
        slot, ev, B, T = handle
        ev.synchronize()
        tags_l = slot["tags"][:B * T].view(B, T).tolist()          # ONE device->host read per batch
        conf_l = slot["conf"][:B * T].view(B, T).tolist()
        slot["busy"] = False
        names = getattr(self, "_tag_names", None)
        if names is None:
            names = self._tag_names = self.tag_dictionary.get_items()
        out = []
        for b, s in enumerate(sentences):
            n = len(s.tokens)
            out.append(LabelSeq(names, tags_l[b][:n], conf_l[b][:n]))
        return out

    def _viterbi_decode(self, feats, all_scores: bool = False, current_idx=0):
        """Single-sentence API of the reference (:1248-1327): feats [T,L] -> (confidences, tag_seq, scores)."""
        if all_scores:
            raise NotImplementedError("all_scores")
        T = feats.shape[0]
        ln = torch.tensor([T], dtype=torch.int32, device=self.device)
        tags, conf = ops.crf_viterbi(feats.detach().float().contiguous()[None], self.transitions.detach().contiguous(),
                                     ln, ln, self.start_idx, self.stop_idx, self.x_idx)
        return conf[0].tolist(), tags[0].tolist(), []

    def _forward_alg(self, feats, lens_, distill_mode=False, T=1):
        if distill_mode or T != 1:
            raise NotImplementedError("distill_mode / temperature")
        B, Tm, L = feats.shape
        lens_ = torch.as_tensor(lens_).to(torch.int32).to(self.device)
        dummy = torch.zeros((B, Tm), dtype=torch.int32, device=self.device)
        logz, _, _ = ops.crf_nll_fwd(feats.detach().float().contiguous(), dummy, self.transitions.detach().contiguous(),
                                     lens_, self.start_idx, self.stop_idx)
        return logz

    def _score_sentence(self, feats, tags, lens_, mask=None):
        """Gold-path score (:2544-2591).  The reference multiplies by `mask` ([B,T], 1 on real tokens); its masks are
        left-packed prefix masks, for which that equals scoring the first mask.sum(1) tokens -- a mask that is not a prefix
        mask (remove-X before compaction) has to go through _calculate_loss, which compacts first like the reference."""
        if mask is not None:
            m = torch.as_tensor(mask).to(self.device) != 0
            n = m.sum(1)
            if bool((m != (torch.arange(m.shape[1], device=m.device)[None, :] < n[:, None])).any()):
                raise ValueError("_score_sentence: mask must be a prefix mask (compact with _calculate_loss / crf_compact first)")
            lens_ = n
        lens_ = torch.as_tensor(lens_).to(torch.int32).to(self.device)
        _, gold, _ = ops.crf_nll_fwd(feats.detach().float().contiguous(), tags.to(torch.int32).contiguous(),
                                     self.transitions.detach().contiguous(), lens_, self.start_idx, self.stop_idx)
        return gold

    # ---- evaluate (:2593-2729) / predict (:786-841) -----------------------------------------------------
    @torch.no_grad()
    def predict(self, sentences, mini_batch_size: int = 32, **_kw):
        if isinstance(sentences, Sentence):
            sentences = [sentences]
        self.eval()
        for i in range(0, len(sentences), mini_batch_size):
            batch = BatchedData(sentences[i:i + mini_batch_size])
            feature = self.forward(batch, prediction_mode=True)
            tags, _ = self._obtain_labels(feature, batch)
            for s, st in zip(batch, tags):
                for tok, lab in zip(s.tokens, st):
                    tok.add_tag_label(self.tag_type, lab)
        return sentences

    @torch.no_grad()
    def evaluate(self, data_loader, out_path=None, embeddings_storage_mode: str = "none", prediction_mode=False,
                 speed_test=False, materialize_labels=False):
        """-> (Result, eval_loss) like the reference (:2593-2729): per-class span counts in a Metric, the remove-X filter
        (:2653-2672) when self.remove_x, the prediction file "text gold pred score" streamed to out_path."""
        self.eval()
        with frozen_gc():
            return self._evaluate(data_loader, out_path, embeddings_storage_mode, prediction_mode, speed_test, materialize_labels)

    def _evaluate(self, data_loader, out_path, embeddings_storage_mode, prediction_mode, speed_test, materialize_labels):
        eval_loss, batches = 0.0, 0
        n_sent, t0 = 0, time.time()
        if speed_test:
            # forward + _obtain_labels only (:2611-2612, :2698-2700), software-pipelined: the Label lists of batch i
            # are built on the host while the kernels of batch i+1 run
            # materialize_labels: build every Label object like the reference's _obtain_labels (:1227-1232) instead of
            # leaving them to be created on access
            pending = None
            finish = (lambda h, b: [list(ls) for ls in self._labels_from_handle(h, b)]) if materialize_labels \
                else self._labels_from_handle
            for batch in data_loader:
                if not isinstance(batch, BatchedData):
                    batch = BatchedData(batch)
                n_sent += len(batch)
                features = self.forward(batch, prediction_mode=prediction_mode)
                handle = self._decode_async(features)
                batch.features = {}            # the EncodedBatch aliases the encoder's buffers: never leave it cached
                if pending is not None:
                    finish(*pending)
                pending = (handle, batch)
            if pending is not None:
                self.last_labels = finish(*pending)
            dt = time.time() - t0
            log.info("speed_test: %d sentences, %.2f sentences/s", n_sent, n_sent / max(dt, 1e-9))
            return {"sentences_per_sec": n_sent / max(dt, 1e-9), "sentences": n_sent}, 0.0
        metric = Metric("Evaluation")
        outfile = open(out_path, "w", encoding="utf-8") if out_path is not None else None
        try:
            for batch in data_loader:
                if not isinstance(batch, BatchedData):
                    batch = BatchedData(batch)
                batches += 1
                n_sent += len(batch)
                features = self.forward(batch, prediction_mode=prediction_mode)
                eval_loss += float(self._calculate_loss(features, batch, self.mask))     # (:2619-2621)
                tags, _ = self._obtain_labels(features, batch)
                for s, st in zip(batch, tags):
                    for tok, lab in zip(s.tokens, st):
                        tok.add_tag_label("predicted", lab)
                        if outfile is not None:                                         # "text gold pred score" (:2626-2643)
                            outfile.write("%s %s %s %s\n" % (tok.text, tok.get_tag(self.tag_type).value, lab.value, lab.score))
                    if outfile is not None:
                        outfile.write("\n")
                    gold_x = [tok.get_tag(self.tag_type).value == "S-X" for tok in s.tokens] if self.remove_x else None
                    span_counts(metric, s.get_spans(self.tag_type), s.get_spans("predicted"), gold_x, self.remove_x)
                store_embeddings(batch, embeddings_storage_mode)
                if hasattr(batch, "features"):
                    batch.features = {}        # the EncodedBatch aliases the encoder's buffers (valid until the next forward)
        finally:
            if outfile is not None:
                outfile.close()
        result = metric.to_result()
        result.metric = metric
        return result, eval_loss / max(batches, 1)

    # ---- checkpoint (:435-477, :1824-1897; flair/nn.py:60-108) ----------------------------------------------
    def _get_state_dict(self):
        """The reference's checkpoint dictionary, key for key (sequence_tagger_model.py:435-477; pinned by
        tests/golden/state_dict_golden.json).  Every option outside the hot path is written with its OFF value on purpose:
        the reference's loader falls back to defaults that are not all off when a key is missing
        (`relearn_embeddings = True if "relearn_embeddings" not in state`, :1881), which would rebuild a different head."""
        off = {k: False for k in (
            "train_initial_hidden_state", "use_mfvi", "use_language_attention", "use_language_vector", "enhanced_crf",
            "use_language_id", "use_transition_attention", "biaf_attention", "token_level_attention",
            "relearn_embeddings", "map_embeddings", "embedding_selector", "new_drop", "use_rl", "use_embedding_masks",
            "embedding_attention", "use_gumbel", "multi_view_training", "calculate_l2_loss", "l2_loss_only")}
        state = {"state_dict": self.state_dict(), "embeddings": self.embeddings, "hidden_size": self.hidden_size,
                 "tag_dictionary": self.tag_dictionary, "tag_type": self.tag_type, "use_crf": self.use_crf,
                 "use_rnn": self.use_rnn, "use_cnn": self.use_cnn, "rnn_layers": self.rnn_layers,
                 "use_word_dropout": self.use_word_dropout, "use_locked_dropout": self.use_locked_dropout,
                 "remove_x": self.remove_x, "sentence_level_loss": self.sentence_level_loss,
                 "target_languages": self.target_languages, "config": self.config,
                 "teacher_hidden": 256, "num_teachers": 4, "relearn_size": -1,
                 "word_map": self.word_map, "char_map": self.char_map}
        state.update(off)
        return state

    @classmethod
    def _init_model_with_state_dict(cls, state, testing=False):
        model = cls(hidden_size=state["hidden_size"], embeddings=state["embeddings"],
                    tag_dictionary=state["tag_dictionary"], tag_type=state["tag_type"], use_crf=state["use_crf"],
                    use_rnn=state["use_rnn"], use_cnn=state.get("use_cnn", False), rnn_layers=state["rnn_layers"],
                    dropout=state.get("use_dropout", 0.0),
                    word_dropout=state.get("use_word_dropout", 0.05), locked_dropout=state.get("use_locked_dropout", 0.5),
                    remove_x=state.get("remove_x", False), sentence_loss=state.get("sentence_level_loss", False),
                    target_languages=state.get("target_languages", 1), config=state.get("config"), testing=testing)
        model.load_state_dict(state["state_dict"])
        return model

    def save(self, model_file):
        torch.save(self._get_state_dict(), str(model_file), pickle_protocol=4)

    @classmethod
    def load(cls, model_file, device=None, tokenizer=None):
        """flair/nn.py:87-108.  Reads checkpoints written by this package AND by the reference (whose pickles name
        flair / transformers-3.0.0 classes: checkpoint_compat maps them, see there)."""
        from .checkpoint_compat import load_reference_checkpoint
        return load_reference_checkpoint(model_file, tokenizer=tokenizer, device=device, tagger_cls=cls)


class _TagProjGrad(torch.autograd.Function):
    """Backward of the fused gather + word-dropout + Linear: one kernel for d_hidden / dW / db, then -- when the
    embeddings are being fine-tuned -- the hand-written encoder backward (encoder.XLMRobertaEncoderB200.backward),
    which accumulates straight into the encoder's gradient arena."""

    @staticmethod
    def forward(ctx, features, weight, bias, enc, drop_keep):
        ctx.enc, ctx.drop_keep = enc, drop_keep
        ctx.save_for_backward(weight)
        return features.view_as(features)

    @staticmethod
    def backward(ctx, g):
        enc, dk = ctx.enc, ctx.drop_keep
        (weight,) = ctx.saved_tensors
        L, H = weight.shape
        d_hidden = torch.zeros((enc.hidden.shape[0], H), dtype=torch.float32, device=g.device)
        dW = torch.zeros((L, H), dtype=torch.float32, device=g.device)
        db = torch.zeros((L,), dtype=torch.float32, device=g.device)
        ops.gather_tagproj_bwd(enc.hidden, enc.row_of, enc.first_idx, weight.detach().float().contiguous(),
                               g.contiguous().float(), enc.S, d_hidden, dW, db, drop_keep=dk)
        if enc.saved is not None:
            enc.encoder.backward(enc.saved, d_hidden)
            enc.saved = None
        return None, dW, db, None, None


class FastSequenceTagger(SequenceTagger):
    """The class the KB-NER YAMLs instantiate (config/*.yaml `model: FastSequenceTagger`,
    sequence_tagger_model.py:1823); identical arithmetic to SequenceTagger on this path."""
