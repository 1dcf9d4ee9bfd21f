#!/usr/bin/env python
"""Quick numerical + timing check of the tensor-core distance kernel against fp64 and the FP32 CUDA-core kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diga_b200 as D
from diga_b200 import _lib as L, synthetic as S

dev = torch.device("cuda", 0)
g = S.gen(17, dev)
for (n, d, h, w, c) in ((1, 256, 8, 16, 19), (1, 256, 65, 129, 19), (2, 2048, 65, 129, 19), (3, 512, 33, 65, 16), (8, 2048, 65, 129, 19),
                        (1, 2048, 129, 257, 19)):
    feat = S.features((n, d, h, w), g)
    cf = D.Class_Features(c, d)
    cf.objective_vectors = S.centroids(c, d, g, feat.mean().item())
    d64 = torch.cdist(feat.double().permute(0, 2, 3, 1).reshape(n, h * w, d), cf.objective_vectors.double().unsqueeze(0).expand(n, -1, -1))
    d64 = d64.reshape(n, h, w, c).permute(0, 3, 1, 2)
    w64 = torch.softmax(-d64, 1)
    res = {}
    for path in (1, 0):
        L.set_tunable("proto_path", path)
        dist = cf.feat_centroid_distance(feat); wt = cf.get_centroid_weight(feat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): cf.get_centroid_weight(feat)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        rel = ((dist.double() - d64).abs().max() / d64.abs().max()).item()
        werr = (wt.double() - w64).abs().max().item()
        flips = int((dist.argmin(1) != d64.argmin(1)).sum())
        gbs = feat.numel() * 4 / ms / 1e6
        print(f"shape {(n,d,h,w,c)} path {'fp32' if path else 'umma'}: dist rel err {rel:.2e} weight abs err {werr:.2e} argmin flips {flips}/{n*h*w} "
              f"{ms*1e3:.1f} us {gbs:.0f} GB/s", flush=True)
L.set_tunable("proto_path", 0)
