#!/usr/bin/env python
"""bench.py -- sentences/sec of the KB-NER hot path (XLM-R-large + CRF, seq 512, 13 tags) on B200.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W [--impl reference]`` prints ONE JSON line.

Workload at every N (BASELINE.json configs[1]): inference, 32 sentences x 512 sub-tokens per GPU per step,
encoder forward + first-sub-token gather / tag projection + batched Viterbi.  One step = one pass of the hot
path over one batch.  N>1 = N independent shards (sentences are independent: no data-path collective), weak
scaling, one process per GPU under torchrun; NCCL is used only for the timing barrier / max-over-ranks.

``value``      device-resident throughput: ids already in HBM, K steps between CUDA events.
``e2e``        same metric through the public API (FastSequenceTagger.forward + _obtain_labels, i.e. the
               reference's evaluate(speed_test=True) body) with HOST inputs: pinned-host ids -> H2D, kernels,
               tags/confidences D2H, Label objects built -- all inside the timed region.
``roofline``   the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event time of its launches in one
               instrumented step, against MEASURED_PEAKS.json's sustained bf16 figure.
``cpu_baseline`` the oracle port (fp32 torch encoder restatement + C Viterbi) on the box's host cores, rank 0.
``--impl reference`` times that same CPU port as the reference arm (the reference's own Python cannot travel to
               the box: /root/reference is absent there; transformers==3.0.0 is not installable offline).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

S_LEN, BATCH, N_TAGS = 512, 32, 13
TAGS = ["B-PER", "I-PER", "E-PER", "S-PER", "B-LOC", "I-LOC", "E-LOC", "S-LOC"]   # + <unk>, O, S-X, START, STOP = 13


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout=6.0):
        """nvidia-smi needs up to a second to start: block until it has written its first sample."""
        t_end = time.time() + timeout
        while self.proc is not None and time.time() < t_end:
            try:
                if os.path.getsize(self.path) > 0:
                    return True
            except OSError:
                pass
            time.sleep(0.05)
        return False

    def stop(self, t0=None, t1=None, t1_wide=None):
        """Samples whose timestamp falls inside the timed region [t0, t1] (epoch seconds).  If fewer than three do (a
        0.2 s region against a 20 ms sampling period plus nvidia-smi's own jitter), the window is widened to t1_wide --
        the end of the end-to-end run, which executes the same kernels -- and `window` says so."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        allrows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 8:
                    try:
                        ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except Exception:
                        ts = None
                    allrows.append((ts, f[1:]))
            os.unlink(self.path)
        except Exception:
            pass

        def inside(lo, hi):
            return [r for ts, r in allrows if ts is None or lo is None or (lo - 0.05 <= ts <= hi + 0.05)]
        rows = inside(t0, t1)
        out["window"] = "timed region"
        if len(rows) < 3 and t1_wide is not None:
            rows = inside(t0, t1_wide)
            out["window"] = "timed region + end-to-end run (same kernels)"
        if len(rows) < 3:
            rows = [r for _, r in allrows]
            out["window"] = "whole sampler lifetime (warm-up included)"
        if rows:
            sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
            try:
                out["sm_max_mhz"] = float(rows[0][1])
            except Exception:
                pass
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(r[3 + i].lower().startswith("active") for r in rows):
                    out["reasons"].append(n)
            out["samples"] = len(rows)
        return out


def synthetic_sentences(n, seed):
    """n sentences of 510 one-piece words => exactly 512 sub-tokens with <s> </s> (SURVEY 8(d) stress case)."""
    import random
    from kbner_b200.data import Sentence
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        out.append(Sentence(tokens=["w%03x" % rnd.randrange(4096) for _ in range(S_LEN - 2)]))
    return out


def build_model(device, large=True):
    import torch
    from kbner_b200.data import Dictionary
    from kbner_b200.embeddings import StackedEmbeddings, SyntheticTokenizer, TransformerWordEmbeddings
    from kbner_b200.encoder import EncoderConfig
    from kbner_b200.sequence_tagger import FastSequenceTagger
    torch.manual_seed(1234)
    cfg = EncoderConfig.xlmr_large() if large else EncoderConfig.xlmr_base()
    emb = TransformerWordEmbeddings(model=cfg.name, layers="-1", pooling_operation="first", fine_tune=False,
                                    tokenizer=SyntheticTokenizer(cfg.vocab_size), config=cfg, device=device)
    with torch.no_grad():     # non-zero biases / LN params so no term of the arithmetic is skipped
        for n, p in emb.model.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.02)
    d = Dictionary.make_tag_dictionary(TAGS, with_x=True)
    assert len(d) == N_TAGS
    tagger = FastSequenceTagger(hidden_size=256, embeddings=StackedEmbeddings([emb]), tag_dictionary=d, tag_type="ner",
                                use_crf=True, use_rnn=False, word_dropout=0.1, locked_dropout=0.0, remove_x=False,
                                sentence_loss=True)
    tagger.eval()
    emb.model.sync_compute_weights()
    return tagger, emb


def encoder_flops_per_sentence(cfg, S):
    H, F, NL = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    per_tok_layer = 2 * (3 * H * H + H * H + 2 * H * F) + 4 * S * H        # = 24H^2 + 4SH for F = 4H
    return NL * S * per_tok_layer


def run_b200(args):
    import torch
    import kbner_b200
    from kbner_b200 import ops
    from kbner_b200.data import BatchedData
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    kbner_b200._lib.check(kbner_b200._lib.load().kbner_device_check(local), "device_check")

    tagger, emb = build_model(dev, large=not args.base)
    cfg = emb.model.config
    K, W = args.steps, args.warmup
    # distinct sentences per step and per rank; tokenisation cache is warmed outside the timed region
    batches = [BatchedData(synthetic_sentences(BATCH, 1000 * rank + i)) for i in range(max(2, min(K + W, 8)))]
    host = []
    for b in batches:
        ids, key_len, row_of, first_idx, lengths, S = emb.build_batch(b)
        host.append((ids.pin_memory(), key_len.pin_memory(), row_of.pin_memory(), first_idx.pin_memory(), lengths, S))
    assert host[0][5] == S_LEN
    devb = [(h[0].to(dev), h[1].to(dev), h[2].to(dev), h[3].to(dev)) for h in host]
    slen = torch.tensor(host[0][4], dtype=torch.int32, device=dev)
    Wt, bt = tagger.linear.weight.float().contiguous(), tagger.linear.bias.float().contiguous()
    trans = tagger.transitions.detach().contiguous()

    def device_step(i):
        ids, key_len, row_of, first_idx = devb[i % len(devb)]
        hidden = emb.model.forward_hidden(ids, key_len)
        logits = ops.gather_tagproj_fwd(hidden, row_of, first_idx, Wt, bt, S_LEN)
        return ops.crf_viterbi(logits, trans, slen, slen, tagger.start_idx, tagger.stop_idx, tagger.x_idx)

    def api_run(k, offset=0):
        """The call a user of the reference makes for throughput: FastSequenceTagger.evaluate(loader, speed_test=True)
        (train.py:147-156 -> sequence_tagger_model.py:2611-2612,2698-2700): forward + _obtain_labels per batch."""
        loader = []
        for i in range(k):
            # a fresh BatchedData per step: static (non-fine-tuned) embeddings are cached per batch object exactly as in
            # the reference (embeddings.py:3030-3037), so re-using an object would skip the encoder
            loader.append(BatchedData(list(batches[(offset + i) % len(batches)])))
        tagger.evaluate(loader, embeddings_storage_mode="none", prediction_mode=True, speed_test=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms, wall * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    with torch.no_grad():
        for i in range(W):
            device_step(i)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            sampler.wait_ready()
            time.sleep(0.3)
            for i in range(W):
                device_step(i)
        l0 = kbner_b200._lib.launch_count()
        tw0 = time.time()
        ms_dev, _ = timed(device_step, K)
        tw1 = time.time()
        launches = kbner_b200._lib.launch_count() - l0
        api_run(W)
        barrier()
        t0 = time.perf_counter()
        api_run(K, offset=W)
        torch.cuda.synchronize()
        wall_e2e = time.perf_counter() - t0
        clocks = sampler.stop(tw0, tw1, time.time()) if rank == 0 else {}
        if dist is not None:
            t = torch.tensor([wall_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall_e2e = float(t[0])

        # ---- roofline of the dominant kernel: events around every GEMM launch of one instrumented step
        gemm_events = []
        real_gemm = ops.gemm_bf16_tn

        def probed(A, B, *a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = real_gemm(A, B, *a, **kw)
            e.record()
            gemm_events.append((s, e, 2.0 * A.shape[0] * A.shape[1] * B.shape[0]))
            return out
        real_gemm_ln = ops.gemm_ln

        def probed_ln(A, Wt, *a, **kw):       # the fused GEMM + bias + residual + LayerNorm launches count with their GEMM FLOPs only
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = real_gemm_ln(A, Wt, *a, **kw)
            e.record()
            gemm_events.append((s, e, 2.0 * A.shape[0] * A.shape[1] * Wt.shape[0]))
            return out
        ops.gemm_bf16_tn = probed
        ops.gemm_ln = probed_ln
        graphs_on = emb.model._use_graphs
        emb.model._use_graphs = False          # the instrumented step launches kernel by kernel
        try:
            device_step(0)
            torch.cuda.synchronize()
        finally:
            ops.gemm_bf16_tn = real_gemm
            ops.gemm_ln = real_gemm_ln
            emb.model._use_graphs = graphs_on
        gemm_ms = sum(s.elapsed_time(e) for s, e, _ in gemm_events)
        gemm_flops = sum(f for _, _, f in gemm_events)

    sust, burst, hbm, how = _peaks()
    n_sent = BATCH * K * world
    value = n_sent / (ms_dev / 1e3)
    e2e_val = n_sent / wall_e2e
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
    flops_sent = encoder_flops_per_sentence(cfg, S_LEN)
    h2d = sum(int(t.numel()) * t.element_size() for t in host[0][:4])
    d2h = BATCH * (S_LEN - 2) * 8
    line = {
        "metric": "sentences/sec XLM-R-large+CRF seq512 (inference: encoder fwd + Viterbi)",
        "value": round(value, 2), "unit": "sentences/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_dev / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic (seeded random-init weights, seeded 512-sub-token sentences)",
        "config": {"workload": "XLM-R-large+CRF inference seq_len=512 batch=32/GPU, 13 tags (BASELINE configs[1])",
                   "batch_per_gpu": BATCH, "seq_len": S_LEN, "tags": N_TAGS, "parallelism": "replicas x%d" % world,
                   "l2": "no flush: per-step working set (1.1 GB bf16 weights + >0.4 GB activations) exceeds the 126 MB L2",
                   "encoder": cfg.name},
        "e2e": {"value": round(e2e_val, 2), "unit": "sentences/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": round(wall_e2e * 1e3 / K, 3),
                "api": "FastSequenceTagger.evaluate(loader, speed_test=True): forward + _obtain_labels per batch, Label objects built"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tcgen05 GEMM kernels (gemm_bf16_kernel + gemm_ln_kernel, %d launches/step; the fused "
                               "LayerNorm epilogues are charged to the GEMM time, their FLOPs are not counted)" % len(gemm_events),
                     "achieved": round(achieved, 1), "peak": sust, "unit": "TFLOP/s", "frac": round(achieved / sust, 4),
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % how, "traffic": _gemm_traffic(),
                     "gemm_ms_per_step": round(gemm_ms, 3),
                     "model_tflops_whole_step": round(flops_sent * BATCH / (ms_dev / K / 1e3) / 1e12, 1)},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_port_baseline(emb, tagger, n_sentences=2, warm=1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def usable_cores():
    """Threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(p) + 0.5)))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = max(1, min(n, int(q / p + 0.5)))
        except Exception:
            pass
    return n


def run_train(args):
    """BASELINE configs[2] / [3]: fine-tuning, 8 sentences x 512 sub-tokens per micro-batch per GPU, gradient
    accumulation 4, fused AdamW + clip 5.0 on every 4th micro-batch.  A step = one micro-batch (forward_loss +
    backward through the hand-written kernels); the optimizer step (and, for N > 1, the NCCL all-reduce of the flat
    gradient arenas -- the only collective of the path) happens inside the timed region on accumulation boundaries.
    Dropout (hidden / attention 0.1, word dropout on the tag projection) is active as in the reference's train() mode."""
    import random
    import torch
    import kbner_b200
    from kbner_b200.data import BatchedData
    from kbner_b200.optim import build_reference_optimizer
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    tagger, emb = build_model(dev, large=not args.base)
    emb.fine_tune, emb.static_embeddings = True, False
    tagger.train()
    emb.train()
    cfg = emb.model.config
    MB, ACC = 8, 4
    K, W = args.steps, args.warmup
    K = (K + ACC - 1) // ACC * ACC            # whole accumulation cycles: every optimizer step of the region is counted
    rnd = random.Random(17 + rank)
    names = tagger.tag_dictionary.get_items()
    legal = [n for n in names if n not in ("<unk>", "S-X", "<START>", "<STOP>")]
    batches = []
    for i in range(4):
        sents = synthetic_sentences(MB, 500 + 10 * rank + i)
        for s in sents:
            for tok in s.tokens:
                tok.add_tag("ner", legal[rnd.randrange(len(legal))])
        batches.append(BatchedData(sents))
    opt = build_reference_optimizer(tagger, lr=5e-6, lr_rate=10000.0)
    opt.set_linear_schedule(1000)
    arenas = [g["arena"] for g in opt.groups]

    from kbner_b200.distributed import OverlappedGradAllReduce
    overlap = dist is not None and OverlappedGradAllReduce.enabled()

    def step(i):
        b = batches[i % len(batches)]
        b.features = {}
        loss = tagger.forward_loss(b) / ACC
        last = (i + 1) % ACC == 0
        if last and overlap:          # all-reduce of finished layer chunks runs under the rest of this backward
            with OverlappedGradAllReduce(emb.model, arenas):
                loss.backward()
        else:
            loss.backward()
        if last:
            if dist is not None and not overlap:
                for ar in arenas:
                    dist.all_reduce(ar.grad)
            opt.step(grad_scale=1.0 / world)
            opt.scheduler_step()
            opt.zero_grad()
            emb.model.sync_compute_weights_arena()
        return loss

    for i in range(max(W, ACC)):
        step(i)
    opt.zero_grad()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
        time.sleep(0.3)
    for i in range(ACC):          # every rank: step() contains the gradient all-reduce
        step(i)
    opt.zero_grad()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = kbner_b200._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    t0 = time.perf_counter()
    e0.record()
    last = None
    for i in range(K):
        last = step(i)
    lossv = float(last.detach())              # device -> host read of the step's result
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop(tw0, time.time()) if rank == 0 else {}
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms, wall * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall = float(t[0]), float(t[1]) / 1e3
    launches = kbner_b200._lib.launch_count() - l0
    flops_sent = 3 * encoder_flops_per_sentence(cfg, S_LEN)
    sust, burst, hbm, how = _peaks()
    n_sent = MB * K * world
    line = {"metric": "sentences/sec XLM-R-large+CRF seq512 (fine-tune: fwd + bwd + CRF loss, AdamW every 4th micro-batch)",
            "value": round(n_sent / (ms / 1e3), 2), "unit": "sentences/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (seeded random-init weights and sentences)",
            "config": {"workload": "XLM-R-large+CRF fine-tune seq_len=512 batch=8 grad-accum=4 (BASELINE configs[2]/[3])",
                       "micro_batch_per_gpu": MB, "grad_accum": ACC, "seq_len": S_LEN, "tags": N_TAGS,
                       "parallelism": "dp%d (NCCL all-reduce of gradients only%s)" % (world, ", overlapped with the last backward" if overlap else ""), "dropout": "hidden %.2f / attention %.2f as in transformers (stateless counter-hash masks, regenerated in the backward)" % (cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob),
                       "l2": "working set (2.2 GB fp32 masters + 0.6 GB bf16 + activations) exceeds the 126 MB L2"},
            "e2e": {"value": round(n_sent / wall, 2), "unit": "sentences/s", "h2d_bytes_per_step": MB * S_LEN * 4 * 2,
                    "d2h_bytes_per_step": 4, "api": "FastSequenceTagger.forward_loss + loss.backward + FusedAdamW.step"},
            "gpu_launches": int(launches), "final_loss": lossv, "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": round(flops_sent * MB / (ms / K / 1e3) / 1e12, 1), "peak": sust,
                         "unit": "TFLOP/s", "frac": round(flops_sent * MB / (ms / K / 1e3) / 1e12 / sust, 4),
                         "note": "whole-step model FLOPs (3 x forward) / step time, not a single kernel", "traffic": None}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def pick_threads():
    """Thread count for the CPU legs: the usable cores, or fewer if a 512x1024x4096 fp32 matmul proxy runs faster
    with fewer (shared hosts oversubscribe badly).  Returns the count it set."""
    import torch
    cores = usable_cores()
    a, b = torch.randn(512, 1024), torch.randn(1024, 4096)
    best, best_t = cores, None
    for n in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(n)
        (a @ b)
        t0 = time.perf_counter()
        for _ in range(10):
            (a @ b)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t * 0.9:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def cpu_port_baseline(emb, tagger, n_sentences, warm, seed=99):
    """The oracle port of the reference's CPU path on this box's host cores: fp32 encoder restatement
    (oracle/encoder_oracle.py, all torch threads) + C Viterbi (oracle/crf_oracle.c, 1 thread), 512-token sentences."""
    import numpy as np
    import torch
    import crf_oracle as O
    import encoder_oracle as E
    cores = pick_threads()
    c = emb.model.config
    cfg = dict(hidden=c.hidden_size, heads=c.num_attention_heads, ffn=c.intermediate_size, layers=c.num_hidden_layers,
               vocab=c.vocab_size, max_pos=c.max_position_embeddings, eps=c.layer_norm_eps, pad_id=c.pad_token_id)
    params = {k: v.detach().float().cpu() for k, v in emb.model.state_dict().items()}
    Wt, bt = tagger.linear.weight.detach().float().cpu(), tagger.linear.bias.detach().float().cpu()
    trans = tagger.transitions.detach().float().cpu().numpy()
    g = torch.Generator().manual_seed(seed)

    def one():
        ids = torch.randint(4, c.vocab_size, (1, S_LEN), generator=g)
        ids[0, 0], ids[0, -1] = 0, 2
        with torch.no_grad():
            h = E.encoder_forward(params, ids, torch.tensor([S_LEN]), cfg)
            logits = torch.nn.functional.linear(h[:, 1:-1], Wt, bt)
        lens = np.array([S_LEN - 2], np.int32)
        O.viterbi(logits.numpy(), trans, lens, start=tagger.start_idx, stop=tagger.stop_idx, x_idx=tagger.x_idx)

    for _ in range(warm):
        one()
    t0 = time.perf_counter()
    for _ in range(n_sentences):
        one()
    dt = time.perf_counter() - t0
    return {"value": round(n_sentences / dt, 4), "unit": "sentences/s", "cores": cores, "kind": "port",
            "sample": "%d x 512-sub-token sentences after %d warm-up, fp32 torch encoder restatement + C Viterbi" %
                      (n_sentences, warm)}


def run_reference(args):
    """Reference arm: the CPU port of the reference's path, all host threads, same config / metric."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import numpy as np
    import crf_oracle as O
    import encoder_oracle as E
    cores = pick_threads()
    cfg = dict(E.XLMR_BASE if args.base else E.XLMR_LARGE)
    params = E.init_params(cfg, seed=1234)
    L = N_TAGS
    g = torch.Generator().manual_seed(7)
    Wt, bt = torch.randn(L, cfg["hidden"], generator=g) * 0.02, torch.zeros(L)
    trans = torch.randn(L, L, generator=g).numpy()
    trans[L - 2, :] = -1e12
    trans[:, L - 1] = -1e12
    per_step = 1

    def step():
        ids = torch.randint(4, cfg["vocab"], (per_step, S_LEN), generator=g)
        ids[:, 0], ids[:, -1] = 0, 2
        with torch.no_grad():
            h = E.encoder_forward(params, ids, torch.full((per_step,), S_LEN), cfg)
            logits = torch.nn.functional.linear(h[:, 1:-1], Wt, bt)
        O.viterbi(logits.numpy(), trans, np.full(per_step, S_LEN - 2, np.int32))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = round(per_step * args.steps / dt, 4)
    sample = "%d x 512-sub-token sentence(s) per step (bounded sample of the 32-sentence batch)" % per_step
    line = {"impl": "reference", "metric": "sentences/sec XLM-R-large+CRF seq512 (inference: encoder fwd + Viterbi)",
            "value": v, "unit": "sentences/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt * 1e3 / args.steps, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "XLM-R-large+CRF inference seq_len=512 batch=32/GPU, 13 tags (BASELINE configs[1])",
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": "sentences/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _gemm_traffic():
    """DRAM bytes per GEMM launch (mean over the four shapes of a layer) from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01", "gemm_dram_traffic.json")) as f:
            return int(json.load(f)["mean_per_launch"])
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50,
                    help="timed steps (default 50: ~0.6 s of device time; the end-to-end arm pays one batch of pipeline fill per run)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--base", action="store_true", help="xlm-roberta-base shapes (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE configs[1] (the contract's default); train = configs[2]/[3] fine-tuning step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
