#!/usr/bin/env python
"""Host-side profile of the public-API inference loop (evaluate(speed_test=True)) -- where the CPU time goes."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from kbner_b200.data import BatchedData
tagger, emb = bench.build_model(torch.device("cuda", 0), large=True)
batches = [BatchedData(bench.synthetic_sentences(32, i)) for i in range(4)]
def run(k):
    loader = []
    for i in range(k):
        b = batches[i % 4]; b.features = {}; loader.append(b)
    tagger.evaluate(loader, speed_test=True)
    torch.cuda.synchronize()
run(4)
pr = cProfile.Profile(); pr.enable(); run(12); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
