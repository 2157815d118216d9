#!/usr/bin/env python
"""Where the end-to-end (public API, host inputs) step time goes: device-side busy / idle time per batch from CUDA events
recorded around FastSequenceTagger.forward + _decode_async, and host timestamps of the same points."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from kbner_b200.data import BatchedData

tagger, emb = bench.build_model(torch.device("cuda", 0), large=True)
batches = [BatchedData(bench.synthetic_sentences(32, i)) for i in range(8)]
for b in batches:
    emb.build_batch(b)
rec = []
fwd, dec, lab = tagger.forward, tagger._decode_async, tagger._labels_from_handle

def forward(*a, **k):
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    r = {"e0": e0, "h0": time.perf_counter()}
    rec.append(r)
    out = fwd(*a, **k)
    e1 = torch.cuda.Event(enable_timing=True); e1.record()
    r["e1"] = e1; r["h1"] = time.perf_counter()
    return out

def decode(f):
    h = dec(f)
    e2 = torch.cuda.Event(enable_timing=True); e2.record()
    rec[-1]["e2"] = e2; rec[-1]["h2"] = time.perf_counter()
    return h

def labels(*a):
    t = time.perf_counter()
    out = lab(*a)
    rec[-1].setdefault("lab", []).append((t, time.perf_counter()))
    return out

tagger.forward, tagger._decode_async, tagger._labels_from_handle = forward, decode, labels

def run(k):
    loader = [BatchedData(list(batches[i % 8])) for i in range(k)]
    tagger.evaluate(loader, speed_test=True, prediction_mode=True)
    torch.cuda.synchronize()

run(4)
rec.clear()
t0 = time.perf_counter(); run(24); wall = time.perf_counter() - t0
rows = []
for i, r in enumerate(rec):
    row = {"i": i, "host_fwd_ms": round((r["h1"] - r["h0"]) * 1e3, 2), "host_dec_ms": round((r["h2"] - r["h1"]) * 1e3, 2),
           "dev_fwd_ms": round(r["e0"].elapsed_time(r["e1"]), 2), "dev_dec_ms": round(r["e1"].elapsed_time(r["e2"]), 2)}
    if i + 1 < len(rec):
        row["dev_gap_to_next_ms"] = round(r["e2"].elapsed_time(rec[i + 1]["e0"]), 3)
        row["host_iter_ms"] = round((rec[i + 1]["h0"] - r["h0"]) * 1e3, 2)
    if "lab" in r:
        row["host_labels_ms"] = [round((b - a) * 1e3, 2) for a, b in r["lab"]]
    rows.append(row)
print(json.dumps({"wall_ms_per_batch": round(wall * 1e3 / 24, 3), "rows": rows[:3] + rows[10:14] + rows[-2:]}, indent=0))
