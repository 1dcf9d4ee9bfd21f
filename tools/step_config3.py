#!/usr/bin/env python
"""One GPU's share of config 3 — the per-step hot path of train_DiGA_gta2city_self_training.py:259-356 without the backbone
(B=8 @512x1024, D=2048) — run N times through the patched call sites of INTEGRATION.md; prints ms per step.  Meant to be
run under `ncu --metrics gpu__time_duration.sum` for the launch list of the step (profiles/).

    python tools/step_config3.py [steps] [dropin|fused]
"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import diga_b200 as D  # noqa: E402
from diga_b200 import synthetic as S  # noqa: E402
from diga_b200.calc_centroids import _labels_on_feature_grid  # noqa: E402

dev = torch.device("cuda", 0)
C, b, hh, ww, h, w, d = 19, 8, 512, 1024, 65, 129, 2048


def build():
    g = S.gen(4321, dev)
    x = {}
    x["sl"] = S.block_labels(b, hh, ww, g)
    x["tl"] = S.perturb_labels(x["sl"], g)
    x["feat"], x["s_feat"] = S.features((b, d, h, w), g), S.features((b, d, h, w), g)
    x["t_pred"], x["s_pred"] = S.logits((b, C, h, w), g), S.logits((b, C, h, w), g)
    x["rec"], x["saug"] = S.images((b, 3, hh, ww), g), S.images((b, 3, hh, ww), g)
    x["tdata_aug"], x["sdata"] = S.images((b, 3, hh, ww), g), S.images((b, 3, hh, ww), g)
    x["tea_cat"], x["stu_cat"] = S.logits((2 * b, C, h, w), g), S.logits((2 * b, C, h, w), g)
    x["cross_low"] = S.logits((b, C, h, w), g)
    cf = D.Class_Features(C, d)
    cf.objective_vectors = S.centroids(C, d, g)
    cf.objective_vectors_num = torch.full((C,), 150.0)
    x["cf"], x["rng"], x["lam"] = cf, random.Random(99), torch.tensor(0.25, device=dev)
    return x


def st_step(x, fused):
    cf, rng, sl, tl, feat = x["cf"], x["rng"], x["sl"], x["tl"], x["feat"]
    if fused:      # one presence pass over slabelv serves both ClassMix blocks; its host round trip hides behind the statements
        # that do not depend on the class choice (a5, a4, the two centroid updates: they read neither mix) — in the script
        # it hides behind the backbone passes.  Label down-sampling (:328-330, :336-337) folded into the assign kernel.
        pres = D.present_classes_async(sl)
        wts = cf.get_centroid_weight(feat)                                                         # :301
        kept, _ = D.consensus_select(tl, wts, (hh, ww))                                            # :302-304
        cf.update_from_features(feat, x["t_pred"], start_mean=False, labels_full=kept)             # :327-334
        cf.update_from_features(x["s_feat"], x["s_pred"], start_mean=False, labels_full=sl)        # :336-341
        _, mix1 = D.classmix(sl, x["rec"], x["saug"], rng=rng, present=pres, return_mask=False)    # :259-275
        _, mix2, mixlabel = D.classmix(sl, x["tdata_aug"], x["sdata"], kept, rng=rng, present=pres, return_mask=False)
    else:
        _, mix1 = D.classmix(sl, x["rec"], x["saug"], rng=rng)
        wts = cf.get_centroid_weight(feat)
        kept, _ = D.consensus_select(tl, wts, (hh, ww))
        _, mix2, mixlabel = D.classmix(sl, x["tdata_aug"], x["sdata"], kept, rng=rng)
    if not fused:
        cf.update_from_features(feat, x["t_pred"], _labels_on_feature_grid(kept, (h, w)), start_mean=False)
        cf.update_from_features(x["s_feat"], x["s_pred"], _labels_on_feature_grid(sl, (h, w)), start_mean=False)
    stu = x["stu_cat"].detach().requires_grad_(True)
    cpm = x["cross_low"].detach().requires_grad_(True)
    if fused:          # loss weights (lambda_seg = 1, lambda_distil = 0.25, :102-103) known up front: one pass each
        part, l_src, l_kd = D.seg_distillation_total_upsampled(x["tea_cat"], stu, sl, 1.0, 0.25, 0.5,   # :289,:348-352,:382
                                                               targets_nonnegative=True)   # loader labels: trainIds or 255
        total = part + D.cross_entropy2d_upsampled(cpm, mixlabel)                                  # :344,:355-356
    else:
        up = lambda t: F.interpolate(t, size=(hh, ww), mode="bilinear", align_corners=True)
        l_src = D.cross_entropy2d(up(stu[:b]), sl)
        l_kd = D.distillation_loss(up(x["tea_cat"]), up(stu), 0.5)
        l_mix = D.cross_entropy2d(up(cpm), mixlabel)
        total = (l_src + l_mix) + x["lam"] * l_kd                                                  # :356,:382
    g_stu, g_mix = torch.autograd.grad(total, [stu, cpm])
    return total, g_stu, g_mix, mix1, mix2


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    fused = (sys.argv[2] if len(sys.argv) > 2 else "fused") == "fused"
    x = build()
    for _ in range(3):
        st_step(x, fused)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        st_step(x, fused)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"{'fused' if fused else 'dropin'}: {ms:.3f} ms per step, {b * hh * ww / ms / 1e6:.2f} Gpx/s")


if __name__ == "__main__":
    main()
