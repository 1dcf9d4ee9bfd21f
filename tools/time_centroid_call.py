#!/usr/bin/env python
"""Call-level timing of the a6 -> a7 chain (Class_Features.update_from_features): eager and CUDA-graph replay, per label pattern."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diga_b200 as D
from diga_b200 import synthetic as S
from tools.sweep_accum import class_logits, PEAK
sys.path.insert(0, ROOT)
from bench import time_loop

dev = torch.device("cuda", 0)
g = S.gen(5, dev)
c, d, h, w = 19, 2048, 65, 129
rows = []
for n in (8, 1):
    feats = [S.features((n, d, h, w), g) for _ in range(3 if n > 1 else 12)]
    cf = D.Class_Features(c, d)
    for pattern in ("iid", "blocks4", "regions16"):
        out = class_logits(pattern, n, c, h, w, g)
        k = [0]

        def call():
            k[0] += 1
            cf.update_from_features(feats[k[0] % len(feats)], out, None, "mean")

        def call_fixed():
            cf.update_from_features(feats[0], out, None, "mean")

        eager = time_loop(call, 50, 5)
        graph = time_loop(call_fixed, 50, 5, graph=True)
        bytes_ = n * h * w * (d * 4 + c * 4 + 1)
        row = {"n": n, "pattern": pattern, "eager_ms": round(eager, 5), "graph_ms": round(graph, 5),
               "frac_eager": round(bytes_ / (eager * 1e-3) / 1e9 / PEAK, 4), "frac_graph": round(bytes_ / (graph * 1e-3) / 1e9 / PEAK, 4),
               "note": "graph replay re-reads one 550 MB (n=8) / 69 MB (n=1: L2-resident!) tensor"}
        rows.append(row)
        print(json.dumps(row), flush=True)
with open(os.path.join(ROOT, "gpurun_out", "time_centroid_call.jsonl"), "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
