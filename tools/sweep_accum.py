#!/usr/bin/env python
"""GPU sweep of the centroid accumulation kernel shapes (diga_centroid_accum, tunable accum_variant) against the label
pattern: i.i.d. classes (worst case for the shared-memory accumulators), 4x4 blocks, 16-px regions, one class.
Every variant is checked against the round-1 kernel (variant 9) before it is timed.  Writes JSON lines.

    python tools/sweep_accum.py [out.jsonl] [variant ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diga_b200 import _lib as L, synthetic as S  # noqa: E402

PEAK = 6530.3
if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def class_logits(pattern, n, c, h, w, g):
    if pattern == "iid":
        return S.logits((n, c, h, w), g)
    if pattern == "one":
        z = S.logits((n, c, h, w), g)
        z[:, 3] += 100.0
        return z
    block = {"blocks4": 4, "regions16": 16}[pattern]
    lab = S.block_labels(n, h, w, g, block, c, 0.0)
    return S.logits((n, c, h, w), g) + 12.0 * torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float()


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "sweep_accum.jsonl")
    variants = [int(v) for v in sys.argv[2:]] or [9, 14, 13, 15, 20, 21, 16, 17, 18, 19, 22]
    dev = torch.device("cuda", 0)
    g = S.gen(11, dev)
    c = 19
    rows = []
    shapes = ((8, 2048, 65, 129), (1, 2048, 65, 129), (8, 256, 65, 129))
    if os.environ.get("SWEEP_QUICK"):
        shapes = shapes[:1]
    for (n, d, h, w) in shapes:
        hw = h * w
        feats = [S.features((n, d, h, w), g) for _ in range(3 if n > 1 else 12)]
        for pattern in ("iid", "blocks4", "regions16", "one"):
            out = class_logits(pattern, n, c, h, w, g)
            cls = torch.empty((n, hw), dtype=torch.uint8, device=dev)
            clsw = torch.empty((int(L.lib.diga_centroid_clsw_bytes(n, hw)) // 4,), dtype=torch.int32, device=dev)
            counts = torch.empty((n, c), dtype=torch.int32, device=dev)
            L.check(L.lib.diga_centroid_assign(out.data_ptr(), None, n, c, hw, cls.data_ptr(), counts.data_ptr(), clsw.data_ptr(), L.stream()))
            ref = None
            for variant in variants:
                L.set_tunable("accum_variant", variant)
                sums = torch.zeros((n, c, d), dtype=torch.float32, device=dev)

                def run(k):
                    L.check(L.lib.diga_centroid_accum(feats[k % len(feats)].data_ptr(), cls.data_ptr(), counts.data_ptr(), clsw.data_ptr(),
                                                      n, d, c, hw, sums.data_ptr(), L.stream()))

                run(0)
                torch.cuda.synchronize()
                got = torch.where((counts > 0).unsqueeze(2), sums, torch.zeros_like(sums))
                if ref is None:
                    ref = got.clone()
                err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
                for k in range(5):
                    run(k)
                torch.cuda.synchronize()
                iters = 30
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(iters):
                    run(k)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                gbs = n * d * hw * 4 / (ms * 1e-3) / 1e9
                row = {"shape": [n, d, h, w], "pattern": pattern, "variant": variant, "ms": round(ms, 5), "gbs": round(gbs, 1),
                       "frac": round(gbs / PEAK, 4), "rel_err_vs_first": err}
                rows.append(row)
                print(json.dumps(row), flush=True)
                assert err < 1e-5, row
        del feats
    L.set_tunable("accum_variant", 0)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
