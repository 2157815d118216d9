#!/usr/bin/env python
"""BASELINE configs[4]: CRF Viterbi / NLL batch sweep 64-4096 sentences x 512 tokens x 13 tags on one B200.
Per-kernel CUDA-event timing (L2 flushed between iterations by writing a 256 MB buffer), algorithmic bytes per
sentence from SURVEY.md 8(d): Viterbi 30,720 B, NLL fwd 28,680 B.  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
from kbner_b200 import ops


def main():
    T, L = 512, int(os.environ.get("L", "13"))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = float(peaks["hbm_gbs"])
    rng = np.random.RandomState(0)
    trans = rng.randn(L, L).astype(np.float32)
    trans[L - 2, :] = -1e12
    trans[:, L - 1] = -1e12
    trans = torch.from_numpy(trans).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"T": T, "L": L, "hbm_peak_gbs": hbm, "rows": []}
    sweep = [int(x) for x in os.environ["SWEEP_B"].split(",")] if os.environ.get("SWEEP_B") else (64, 128, 256, 512, 1024, 2048, 4096, 16384)
    for B in sweep:
        emis = torch.randn(B, T, L, device="cuda") * 3
        lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
        tags = torch.randint(1, L - 2, (B, T), device="cuda", dtype=torch.int32)
        res = {"B": B}
        _, _, alpha = ops.crf_nll_fwd(emis, tags, trans, lens, L - 2, L - 1, want_alpha=True)
        w = torch.full((B,), 1.0 / B, device="cuda")
        for name, fn, bytes_per in (
                ("viterbi", lambda: ops.crf_viterbi(emis, trans, lens, lens, L - 2, L - 1), T * L * 4 + T * 8),
                ("nll_fwd", lambda: ops.crf_nll_fwd(emis, tags, trans, lens, L - 2, L - 1), T * L * 4 + T * 4 + 8),
                ("nll_fwd_alpha", lambda: ops.crf_nll_fwd(emis, tags, trans, lens, L - 2, L - 1, want_alpha=True),
                 2 * T * L * 4 + T * 12 + 8),
                # emissions + stored alpha (fp32 vector + fp64 scale) + tags read, d_emis written
                ("nll_bwd", lambda: ops.crf_nll_bwd(emis, tags, trans, lens, alpha, w, L - 2, L - 1),
                 3 * T * L * 4 + T * 12)):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ms = sorted(ts)[len(ts) // 2]
            gbs = B * bytes_per / (ms / 1e3) / 1e9
            res[name] = {"ms": round(ms, 4), "sent_per_s": round(B / (ms / 1e3), 1), "GBps": round(gbs, 1),
                         "frac_hbm": round(gbs / hbm, 4)}
        out["rows"].append(res)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
