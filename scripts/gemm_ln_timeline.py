#!/usr/bin/env python
"""clock64 timeline of the fused GEMM + LayerNorm kernel (grid version): where an item's epilogue spends its time.
Per CTA: MMA warp (1) stamps 0 wait-empty / 1 acquired / 2 committed; epilogue warps (2..9) stamps
0 loop top / 1 accumulator ready / 2 residual landed / 3 pass 1 done / 4 published + z stored / 5 statistics complete /
6 merged / 7 pass 2 issued.  Prints medians over CTAs in SM cycles relative to the CTA's first stamp, the spread of the
phases over warps, and the start skew of the CTAs on the global timer."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kbner_b200 import ops, _lib

lib = _lib.load()
shapes = ((16384, 1024, 1024), (16384, 1024, 4096))
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias, gamma, beta = torch.randn(N, device="cuda"), torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ws = ops.gemm_ln_workspace(M, N, "cuda")
    for _ in range(3):
        ops.gemm_ln(a, w, bias, resid, gamma, beta, 1e-5, out=out, ws=ws)
    tl = torch.zeros(148 * 10 * 8 * 8, dtype=torch.int64, device="cuda")
    lib.kbner_debug_gemm_ln_timeline(tl.data_ptr())
    ops.gemm_ln(a, w, bias, resid, gamma, beta, 1e-5, out=out, ws=ws)
    torch.cuda.synchronize()
    lib.kbner_debug_gemm_ln_timeline(None)
    t = tl.view(148, 10, 8, 8).cpu().double()
    t[t == 0] = float("nan")
    base = t[:, 1:, 0, 0].nan_to_num(nan=float("inf")).amin(dim=1)           # first stamp of the CTA
    rel = t - base[:, None, None, None]
    rows = {"shape": [M, N, K]}
    mma = rel[0::2, 1]                                                        # leader CTAs
    epi = rel[:, 2:10]
    q = lambda x, p: round(float(torch.nanquantile(x.flatten(), p)), 0)
    for it in range(4):
        rows["item%d" % it] = {
            "mma": [round(float(torch.nanmedian(mma[:, it, e])), 0) for e in range(3)],
            "epi_median": [round(float(torch.nanmedian(epi[:, :, it, e])), 0) for e in range(8)],
            "epi_max": [round(float(epi[:, :, it, e].nan_to_num(nan=-1).amax()), 0) for e in range(8)],
        }
    # phase durations of item 1 (steady state): quantiles over all epilogue warps, and medians per warp index / CTA parity
    ph = {"pass1": (2, 3), "wait_stats": (4, 5), "pass2": (6, 7), "whole": (1, 7)}
    for name, (e0, e1) in ph.items():
        d = epi[:, :, 1, e1] - epi[:, :, 1, e0]
        rows[name + "_q10_50_90_100"] = [q(d, 0.1), q(d, 0.5), q(d, 0.9), q(d, 1.0)]
        rows[name + "_by_warp"] = [round(float(torch.nanmedian(d[:, i])), 0) for i in range(8)]
        rows[name + "_by_cta_parity"] = [round(float(torch.nanmedian(d[0::2])), 0), round(float(torch.nanmedian(d[1::2])), 0)]
    # global-time skew: when did each CTA's epilogue warps start (ns, relative to the earliest), and absolute time of
    # "pass 1 done" of item 0 / 1 per CTA pair (converted with the measured clock rate)
    g0 = t[:, 2, 7, 0]
    c0 = t[:, 2, 7, 1]
    start_ns = g0 - torch.nan_to_num(g0, nan=float("inf")).min()
    rows["cta_start_ns_q0_50_90_100"] = [q(start_ns, 0.0), q(start_ns, 0.5), q(start_ns, 0.9), q(start_ns, 1.0)]
    cyc_per_ns = 1.9
    for it in (0, 1):
        p1 = start_ns + (t[:, 2, it, 3] - c0) / cyc_per_ns                    # warp 2 of every CTA, absolute ns
        grp = p1[: (p1.numel() // 8) * 8].view(-1, 8)                        # 4 pairs x 2 CTAs of one panel
        spread = grp.amax(dim=1) - grp.amin(dim=1)
        rows["item%d_pass1_done_spread_within_panel_ns_q50_90_100" % it] = [q(spread, 0.5), q(spread, 0.9), q(spread, 1.0)]
        rows["item%d_pass1_done_abs_ns_q0_50_100" % it] = [q(p1, 0.0), q(p1, 0.5), q(p1, 1.0)]
    print(json.dumps(rows), flush=True)
