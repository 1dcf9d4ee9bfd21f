"""[round-1 ABI: diga_centroid_accum has since gained the counts / class-word arguments; kept as the record of the round-1
experiment, not runnable against the current library]
Two experiments, one GPU call (results -> gpurun_out/exp_accum_overlap.jsonl):
 (1) centroid_accum_quad_kernel against the label pattern (constant / large regions / 4x4 blocks / iid) for the default
     launch shape and the cross-item prefetch variants (13-15);
 (2) config 5 per image: the ALU-bound two-scale label kernel on a side stream beside the HBM-bound distance kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diga_b200 as D  # noqa: E402
from diga_b200 import _lib as L, synthetic as S  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = 6530.3
fout = open(os.path.join(ROOT, "gpurun_out", "exp_accum_overlap.jsonl"), "a")


def emit(rec):
    print(json.dumps(rec), flush=True)
    fout.write(json.dumps(rec) + "\n")
    fout.flush()


def timeit(fn, iters=30, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def accum():
    g = S.gen(3, dev)
    for (n, d, h, w) in ((8, 2048, 65, 129), (8, 2048, 64, 128)):
        feat = S.features((n, d, h, w), g)
        sums = torch.empty((n, 19, d), device=dev)
        ref = None
        pats = {
            "constant": torch.full((n, h * w), 3, device=dev, dtype=torch.uint8),
            "regions16": S.block_labels(n, h, w, S.gen(9, dev), 16, 19, 0.1).reshape(n, h * w).to(torch.uint8).contiguous(),
            "blocks4": S.block_labels(n, h, w, S.gen(9, dev), 4, 19, 0.1).reshape(n, h * w).to(torch.uint8).contiguous(),
            "iid": torch.randint(0, 19, (n, h * w), device=dev, dtype=torch.uint8),
        }
        for name, cls in pats.items():
            ref = None
            for variant in (0, 16, 17):
                L.set_tunable("accum_variant", variant)
                fn = lambda: L.check(L.lib.diga_centroid_accum(feat.data_ptr(), cls.data_ptr(), n, d, 19, h * w, sums.data_ptr(), L.stream()))
                ms = timeit(fn)
                if ref is None:
                    ref = sums.clone()
                same = bool(torch.equal(ref, sums))
                err = float(((ref - sums).abs().amax(-1) / ref.abs().amax(-1).clamp_min(1e-30)).max())
                emit({"exp": "accum", "shape": [n, d, h, w], "labels": name, "variant": variant, "ms": round(ms, 4),
                      "frac": round(feat.numel() * 4 / ms / 1e6 / PEAK, 3), "bit_equal_to_default": same, "max_rel_err_vs_default": err})
        L.set_tunable("accum_variant", 0)
        del feat


def config5():
    g = S.gen(4321, dev)
    C, d = 19, 2048
    cf = D.Class_Features(C, d)
    cf.objective_vectors = S.centroids(C, d, g)
    pool = [(S.features((1, d, 129, 257), g), S.logits((1, C, 129, 257), g), S.logits((1, C, 65, 129), g)) for _ in range(4)]
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    n_img = 400

    def serial():
        for k in range(n_img):
            f, la, lb = pool[k % 4]
            lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
            D.consensus_select(lab, cf.get_centroid_weight(f), want_feat_pseudo=False)

    def overlap(label_first):
        evs = [torch.cuda.Event() for _ in range(n_img)]
        side.wait_stream(main)
        for k in range(n_img):
            f, la, lb = pool[k % 4]
            if label_first:
                with torch.cuda.stream(side):
                    lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
                    evs[k].record(side)
                wts = cf.get_centroid_weight(f)
            else:
                wts = cf.get_centroid_weight(f)
                with torch.cuda.stream(side):
                    lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
                    evs[k].record(side)
            main.wait_event(evs[k])
            lab.record_stream(main)
            D.consensus_select(lab, wts, want_feat_pseudo=False)
        main.wait_stream(side)

    def parts():
        f, la, lb = pool[0]
        lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
        wts = cf.get_centroid_weight(f)
        return {"label_ms": timeit(lambda: D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)),
                "dist_ms": timeit(lambda: cf.get_centroid_weight(pool[1][0])),
                "select_ms": timeit(lambda: D.consensus_select(lab, wts, want_feat_pseudo=False))}

    emit({"exp": "config5_parts", **{k: round(v, 4) for k, v in parts().items()}})
    for name, fn in (("serial", serial), ("overlap_label_first", lambda: overlap(True)), ("overlap_dist_first", lambda: overlap(False))):
        ms = timeit(fn, iters=3, warm=1)
        emit({"exp": "config5", "mode": name, "us_per_image": round(ms * 1e3 / n_img, 2)})


if __name__ == "__main__":
    which = sys.argv[1:] or ["accum", "config5"]
    if "accum" in which:
        accum()
    if "config5" in which:
        config5()
