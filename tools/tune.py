#!/usr/bin/env python
"""Launch-shape sweep on the GPU: times each hot kernel for every tunable combination with CUDA events and prints
one JSON line per measurement (gpurun_out/tune.jsonl).  Defaults in csrc/*.cu are set from these sweeps."""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diga_b200 as D  # noqa: E402
from diga_b200 import _lib as L, synthetic as S  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = 6530.3
out_path = os.path.join(ROOT, "gpurun_out", "tune.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, "a")


def timeit(fn, iters=20, warm=3):
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


def report(kernel, cfg, ms, nbytes):
    rec = {"kernel": kernel, **cfg, "ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / PEAK, 3)}
    print(json.dumps(rec), flush=True)
    fout.write(json.dumps(rec) + "\n")
    fout.flush()


def sweep_kd():
    g = S.gen(1, dev)
    shape = (8, 19, 512, 1024)
    sets = [(S.logits(shape, g), S.logits(shape, g)) for _ in range(2)]
    px = 8 * 512 * 1024
    ws = L.kd_workspace(dev)
    loss = torch.empty((), device=dev)
    ds = torch.empty(shape, device=dev)
    up = torch.tensor(0.25, device=dev)
    k = [0]

    def fwd():
        t, s = sets[k[0] % 2]; k[0] += 1
        L.check(L.lib.diga_kd_fwd(t.data_ptr(), s.data_ptr(), 8, 19, 512 * 1024, 0.5, loss.data_ptr(), ws.data_ptr(), L.stream()))

    def bwd():
        t, s = sets[k[0] % 2]; k[0] += 1
        L.check(L.lib.diga_kd_bwd(t.data_ptr(), s.data_ptr(), 8, 19, 512 * 1024, 0.5, up.data_ptr(), ds.data_ptr(), L.stream()))

    def both():
        t, s = sets[k[0] % 2]; k[0] += 1
        L.check(L.lib.diga_kd_fwd_bwd(t.data_ptr(), s.data_ptr(), 8, 19, 512 * 1024, 0.5, 0.25, loss.data_ptr(), ds.data_ptr(), ws.data_ptr(), L.stream()))

    for vec, block, waves in itertools.product((2, 1), (256, 128), (1, 2, 4, 8)):
        L.set_tunable("kd_vec", vec); L.set_tunable("kd_block", block); L.set_tunable("kd_waves_fwd", waves); L.set_tunable("kd_waves_bwd", waves)
        cfg = {"vec": vec, "block": block, "waves": waves}
        report("kd_fwd", cfg, timeit(fwd), px * 152)
        report("kd_bwd", cfg, timeit(bwd), px * 228)
        report("kd_fwd_bwd", cfg, timeit(both), px * 228)
    L.set_tunable("kd_vec", 2); L.set_tunable("kd_block", 256); L.set_tunable("kd_waves_fwd", 1); L.set_tunable("kd_waves_bwd", 8)


def sweep_pl():
    g = S.gen(2, dev)
    n, hh, ww = 4, 1024, 2048
    z, z2 = S.logits((n, 19, hh, ww), g), S.logits((n, 19, hh, ww), g)
    px = n * hh * ww
    for vec, waves in itertools.product((4, 2, 1), (1, 2, 4, 8)):
        L.set_tunable("pl_vec", vec); L.set_tunable("pl_vec2", vec); L.set_tunable("pl_waves", waves); L.set_tunable("pl_waves2", waves)
        cfg = {"vec": vec, "waves": waves}
        report("pseudo_label_1", cfg, timeit(lambda: D.pseudo_label(z)), px * 81)
        report("pseudo_label_2", cfg, timeit(lambda: D.pseudo_label(z, z2)), px * 157)
    L.set_tunable("pl_vec", 4); L.set_tunable("pl_vec2", 2); L.set_tunable("pl_waves", 2); L.set_tunable("pl_waves2", 8)


def sweep_accum():
    """Superseded by tools/sweep_accum.py (round 2: every shape x label pattern x kernel variant, checked against the round-1
    kernel); kept as the short form.  Class maps are plain byte maps here, so the class words come from diga_centroid_clsw_build."""
    g = S.gen(3, dev)
    for (n, d, h, w) in ((8, 2048, 65, 129), (8, 2048, 64, 128), (1, 2048, 65, 129)):
        hw = h * w
        feat = S.features((n, d, h, w), g)
        sums = torch.empty((n, 19, d), device=dev)
        clsw = torch.empty((int(L.lib.diga_centroid_clsw_bytes(n, hw)) // 4,), dtype=torch.int32, device=dev)
        for label, cls in (("random", torch.randint(0, 19, (n, hw), device=dev, dtype=torch.uint8)),
                           ("4x4 blocks", S.block_labels(n, h, w, S.gen(9, dev), 4, 19, 0.1).reshape(n, hw).to(torch.uint8).contiguous())):
            L.check(L.lib.diga_centroid_clsw_build(cls.data_ptr(), n, 19, hw, clsw.data_ptr(), L.stream()))
            for variant in (0, 14, 17, 9, 2):
                L.set_tunable("accum_variant", variant)
                ms = timeit(lambda: L.check(L.lib.diga_centroid_accum(feat.data_ptr(), cls.data_ptr(), None, clsw.data_ptr(), n, d, 19, hw,
                                                                      sums.data_ptr(), L.stream())))
                report("centroid_accum", {"variant": variant, "shape": [n, d, h, w], "labels": label}, ms, feat.numel() * 4)
    L.set_tunable("accum_variant", 0)


def sweep_select():
    g = S.gen(6, dev)
    b, hh, ww, h, w = 8, 512, 1024, 65, 129
    wl = torch.softmax(S.logits((b, 19, h, w), g), 1)
    tl = S.block_labels(b, hh, ww, g)
    for px, ry in itertools.product((1, 2), (8, 16, 32, 64, 128)):
        L.set_tunable("select_px", px); L.set_tunable("select_ry", ry)
        report("consensus_select", {"px": px, "ry": ry}, timeit(lambda: D.consensus_select(tl, wl)), b * hh * ww * 24)
        report("consensus_select_no_feat_pseudo", {"px": px, "ry": ry},
               timeit(lambda: D.consensus_select(tl, wl, want_feat_pseudo=False)), b * hh * ww * 16)
    L.set_tunable("select_px", 2); L.set_tunable("select_ry", 16)


def sweep_cm():
    import random
    g = S.gen(4, dev)
    b, hh, ww = 8, 512, 1024
    sl = S.block_labels(b, hh, ww, g); tl = S.perturb_labels(sl, g)
    xa, xb = S.images((b, 3, hh, ww), g), S.images((b, 3, hh, ww), g)
    from diga_b200.classmix import present_classes, select_classes
    classes = select_classes(present_classes(sl), random.Random(1))
    px = b * hh * ww
    import numpy as np
    lut = np.zeros((b, 256), dtype=np.uint8)
    for i, sel in enumerate(classes):
        lut[i, sel] = 1
    mix = torch.empty_like(xa); ml = torch.empty_like(tl); mask = torch.empty((b, hh, ww), device=dev)
    for vec, waves in itertools.product((4, 2, 1), (1, 2, 4, 8)):
        L.set_tunable("cm_vec", vec); L.set_tunable("cm_waves", waves)
        fn = lambda: L.check(L.lib.diga_classmix_blend(sl.data_ptr(), lut.ctypes.data, xa.data_ptr(), xb.data_ptr(), tl.data_ptr(),
                                                       b, 3, hh * ww, None, mix.data_ptr(), ml.data_ptr(), L.stream()))
        report("classmix_blend_dacs_kernel", {"vec": vec, "waves": waves}, timeit(fn), px * 60)
        fn2 = lambda: L.check(L.lib.diga_classmix_blend(sl.data_ptr(), lut.ctypes.data, xa.data_ptr(), xb.data_ptr(), None,
                                                        b, 3, hh * ww, mask.data_ptr(), mix.data_ptr(), None, L.stream()))
        report("classmix_blend_img_mask_kernel", {"vec": vec, "waves": waves}, timeit(fn2), px * 48)
    L.set_tunable("cm_vec", 4); L.set_tunable("cm_waves", 2)
    report("classmix_python_total", {}, timeit(lambda: D.classmix(sl, xa, xb, tl, rng=random.Random(1), return_mask=False)), px * 68)
    bm = torch.empty((b, 8), dtype=torch.int32, device=dev); fl = torch.empty(1, dtype=torch.int32, device=dev)
    report("class_presence", {}, timeit(lambda: L.check(L.lib.diga_class_presence(sl.data_ptr(), b, hh * ww, bm.data_ptr(), fl.data_ptr(), L.stream()))), px * 8)


def sweep_ce():
    g = S.gen(6, dev)
    n, hh, ww = 8, 512, 1024
    x = S.logits((n, 19, hh, ww), g); t = S.block_labels(n, hh, ww, g)
    ws = L.ce_workspace(dev); loss = torch.empty((), device=dev); den = torch.empty((), device=dev)
    dx = torch.empty_like(x); up = torch.tensor(1.0, device=dev)
    px = n * hh * ww
    for vec, waves in itertools.product((2, 1), (1, 2, 4, 8)):
        L.set_tunable("ce_vec", vec); L.set_tunable("ce_waves_fwd", waves); L.set_tunable("ce_waves_bwd", waves)
        report("ce_fwd", {"vec": vec, "waves": waves}, timeit(lambda: L.check(L.lib.diga_cross_entropy2d_fwd(
            x.data_ptr(), t.data_ptr(), None, n, 19, hh * ww, 1, loss.data_ptr(), den.data_ptr(), ws.data_ptr(), L.stream()))), px * 84)
        report("ce_bwd", {"vec": vec, "waves": waves}, timeit(lambda: L.check(L.lib.diga_cross_entropy2d_bwd(
            x.data_ptr(), t.data_ptr(), None, n, 19, hh * ww, 1, up.data_ptr(), den.data_ptr(), dx.data_ptr(), L.stream()))), px * 160)
    L.set_tunable("ce_vec", 2); L.set_tunable("ce_waves_fwd", 1); L.set_tunable("ce_waves_bwd", 8)


def sweep_misc():
    g = S.gen(5, dev)
    b, hh, ww, h, w, d = 8, 512, 1024, 65, 129, 2048
    feat = S.features((b, d, h, w), g)
    cf = D.Class_Features(19, d); cf.objective_vectors = S.centroids(19, d, g)
    report("proto_distance_fp32", {"shape": [b, d, h, w]}, timeit(lambda: cf.get_centroid_weight(feat)), feat.numel() * 4)
    wl = cf.get_centroid_weight(feat)
    tl = S.block_labels(b, hh, ww, g)
    report("consensus_select", {}, timeit(lambda: D.consensus_select(tl, wl)), b * hh * ww * 24)
    l1, l2 = S.logits((4, 19, 129, 257), g), S.logits((4, 19, 65, 129), g)
    report("pseudo_label_fused_upsample", {}, timeit(lambda: D.pseudo_label_two_scale(l1, l2, (1024, 2048))), 4 * 1024 * 2048 * 5)
    # reference-point: torch copy bandwidth on this box
    a = torch.empty(1 << 28, device=dev); bb = torch.empty_like(a)
    report("torch_copy_1GiB", {}, timeit(lambda: bb.copy_(a)), a.numel() * 8)


if __name__ == "__main__":
    which = sys.argv[1:] or ["kd", "pl", "accum", "cm", "ce", "misc"]
    for name in which:
        globals()["sweep_" + name]()
