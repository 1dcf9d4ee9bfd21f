#!/usr/bin/env python
"""Arg-max label kernels (a4 consensus_select, a3+f1 pseudo_label_two_scale) on network-like (peaked) and i.i.d. inputs.
The `prune_max` tunables only exist in the experiment build kept as tools/experiments/r02_select_pruned.cu.txt (exact candidate
pruning: measured, not faster, not shipped — DESIGN.md §8); on the shipped library both settings time the same kernel.
JSON lines to gpurun_out/time_labels.jsonl."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diga_b200 as D
from diga_b200 import _lib as L, synthetic as S
from bench import time_loop

dev = torch.device("cuda", 0)
g = S.gen(9, dev)
rows = []


def report(**kw):
    rows.append(kw)
    print(json.dumps(kw), flush=True)


for name, (b, lo, hi) in {"config3 B=8 65x129->512x1024": (8, (65, 129), (512, 1024)), "config5 1x129x257->1024x2048": (1, (129, 257), (1024, 2048))}.items():
    pl = S.block_labels(b, hi[0], hi[1], g, 16, 19)
    for kind in ("peaked", "iid"):
        z = S.logits_peaked((b, 19, *lo), g) if kind == "peaked" else S.logits((b, 19, *lo), g)
        wl = torch.softmax(z, 1)
        for dt in (torch.int64, torch.uint8):
            lab = pl.to(dt)
            for prune in (0, 12):
                L.set_tunable("select_prune_max", prune)
                ms = time_loop(lambda: D.consensus_select(lab, wl, want_feat_pseudo=False), 50, 5, graph=True)
                report(kernel="consensus_select", shape=name, input=kind, labels=str(dt).split(".")[-1], prune_max=prune, ms=round(ms, 5),
                       gpx_per_s=round(b * hi[0] * hi[1] / ms / 1e6, 1))
        if hasattr(D, "pseudo_label_two_scale"):
            z2 = torch.nn.functional.interpolate(z, size=((lo[0] + 1) // 2, (lo[1] + 1) // 2), mode="bilinear", align_corners=True)
            if kind == "iid":
                z2 = S.logits(tuple(z2.shape), g)
            for prune in (0, 12):
                L.set_tunable("plu_prune_max", prune)
                ms = time_loop(lambda: D.pseudo_label_two_scale(z, z2, hi, want_conf=False), 50, 5, graph=True)
                report(kernel="pseudo_label_two_scale", shape=name, input=kind, prune_max=prune, ms=round(ms, 5),
                       gpx_per_s=round(b * hi[0] * hi[1] / ms / 1e6, 1))
L.set_tunable("select_prune_max", 12)
L.set_tunable("plu_prune_max", 12)
with open(os.path.join(ROOT, "gpurun_out", "time_labels.jsonl"), "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
