#!/usr/bin/env python
"""Timings of the two small loader / evaluation kernels (f5 confusion matrix, f3 label resize + remap) replayed from a CUDA
graph, over the `confusion_ctas_per_sm` tunable and two kinds of prediction maps (noise-like: arg-max of i.i.d. stride-8
logits; network-like: the ground truth with perturbed 8x8 blocks).  Prints one JSON line per case.

    python tools/time_eval_loader.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diga_b200 as D  # noqa: E402
from diga_b200 import _lib as L, synthetic as S  # noqa: E402
from diga_b200.util.labels import resize_remap_labels, trainid_lut  # noqa: E402
from diga_b200.util.metrics import runningScore  # noqa: E402

dev = torch.device("cuda", 0)
C, n, hh, ww = 19, 4, 1024, 2048


def replay_ms(fn, iters=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    g = S.gen(1234, dev)
    gt = S.block_labels(n, hh, ww, g)
    l1, l2 = S.logits((n, C, 129, 257), g), S.logits((n, C, 65, 129), g)
    preds = {"noise": D.pseudo_label_two_scale(l1, l2, (hh, ww), want_conf=False)[0],
             "network-like": S.perturb_labels(gt, g).to(torch.uint8)}
    for kind, pred in preds.items():
        for per_sm in (2, 4, 8, 16):
            L.set_tunable("confusion_ctas_per_sm", per_sm)
            rs = runningScore(C)
            ms = replay_ms(lambda: rs.update(gt, pred))
            print(json.dumps({"kernel": "confusion", "pred": kind, "ctas_per_sm": per_sm, "us": round(ms * 1e3, 2),
                              "gbs": round(n * hh * ww * 9 / ms / 1e6, 1)}))
    L.set_tunable("confusion_ctas_per_sm", 8)
    raw = torch.randint(0, 34, (8, 1024, 2048), device=dev, dtype=torch.uint8, generator=g)
    lut = trainid_lut()
    ms = replay_ms(lambda: resize_remap_labels(raw, (512, 1024), lut))
    print(json.dumps({"kernel": "label_resize_remap", "us": round(ms * 1e3, 2), "gbs": round(8 * 512 * 1024 * 9 / ms / 1e6, 1)}))


if __name__ == "__main__":
    main()
