"""Generate the committed golden fixtures by running the REFERENCE ITSELF (build container only).

Run:  ``python tests/golden/make_golden.py``  (needs ``/root/reference``; writes ``tests/golden/*.npz``).

* Function-level hot path (``distillation_loss``, ``process_label``, ``Class_Features``): the real
  functions are imported through ``oracle/ref_loader.py``.
* Inline script blocks (pseudo-label math, ClassMix, consensus selection, online centroid update):
  the scripts cannot be imported (argparse / dataset / checkpoint side effects at module level),
  so the exact source *line ranges* are read from the reference file at generation time and
  ``exec``-ed over seeded inputs.  No reference source text is stored in this repo; only the
  numeric inputs and outputs are.

Every fixture stores its inputs, so tests never need the reference at run time.
"""
from __future__ import annotations

import os
import random
import sys
import textwrap

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

G = os.path.join(ref_loader.REF_ROOT, ref_loader.TREES["G"])


def _lines(relpath: str, first: int, last: int) -> str:
    """Source text of 1-based inclusive line range of a reference file, dedented."""
    with open(os.path.join(G, relpath)) as f:
        src = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(src))


def _blocky_labels(gen, b, h, w, block, n_cls=19, p_ignore=0.1):
    """Piecewise-constant int64 label maps with ~p_ignore of the blocks set to 255."""
    gh, gw = (h + block - 1) // block, (w + block - 1) // block
    coarse = torch.randint(0, n_cls, (b, gh, gw), generator=gen)
    coarse[torch.rand((b, gh, gw), generator=gen) < p_ignore] = 255
    lab = coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :h, :w]
    return lab.contiguous().long()


ONLY = set(sys.argv[1:])      # `python make_golden.py losses_up` rewrites only the named fixtures


def save(name, **arrays):
    if ONLY and name not in ONLY:
        return
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def main():
    ref = ref_loader.load("G")
    gen = torch.Generator().manual_seed(20231017)

    # ---- a1 KD loss fwd + bwd (G/util/loss.py:125-143) ------------------------------------
    for tag, shape, scale, sigma in (("kd_c19", (4, 19, 8, 16), 0.5, 3.0),
                                     ("kd_c16_s025", (2, 16, 6, 10), 0.25, 3.0),
                                     ("kd_saturated", (4, 19, 4, 12), 0.5, 12.0)):
        t = sigma * torch.randn(shape, generator=gen)
        s = (sigma * torch.randn(shape, generator=gen)).requires_grad_(True)
        loss = ref.distillation_loss(t, s, scale)
        (loss * 0.25).backward()                       # upstream grad 0.25 (lambda_distil-like)
        save(tag, teacher=t, student=s, scale=scale, upstream=0.25, loss=loss, grad=s.grad)

    # ---- f2 cross_entropy2d (G/util/loss.py:48-62) ---------------------------------------------
    for tag, shape, use_w, avg in (("ce_c19", (3, 19, 7, 10), False, True), ("ce_weighted_sum", (2, 16, 5, 6), True, False)):
        x = (3.0 * torch.randn(shape, generator=gen)).requires_grad_(True)
        tgt = torch.randint(0, shape[1], (shape[0], shape[2], shape[3]), generator=gen)
        tgt[torch.rand(tgt.shape, generator=gen) < 0.2] = 255
        tgt[0, 0, :3] = -1                                   # dropped pixels (target >= 0 mask, :52,:56)
        wt = (0.5 + torch.rand(shape[1], generator=gen)) if use_w else None
        loss = ref.cross_entropy2d(x, tgt, weight=wt, size_average=avg)
        (loss * 0.7).backward()
        save(tag, input=x, target=tgt, weight=(wt if use_w else torch.zeros(0)), size_average=avg, upstream=0.7,
             loss=loss, grad=x.grad)

    # ---- f4 EMA teacher update (G/util/utils.py:103-116) ---------------------------------------
    def small_net(seed):
        torch.manual_seed(seed)
        return nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Conv2d(8, 5, 1), nn.Linear(7, 3))
    teacher, student = small_net(1), small_net(2)
    before = [p.detach().clone().reshape(-1) for p in teacher.parameters()]
    stud = [p.detach().clone().reshape(-1) for p in student.parameters()]
    after = {}
    for it, kw in ((0, {}), (7, {}), (5000, {}), (3, {"stage0": False, "mean": True}), (3, {"stage0": False})):
        t2 = small_net(1)
        ref.update_teacher_params(t2, student, it, **kw)
        after[f"after_it{it}_{'_'.join(k for k in kw) or 'stage0'}"] = torch.cat([p.detach().reshape(-1) for p in t2.parameters()])
    save("ema", teacher=torch.cat(before), student=torch.cat(stud),
         sizes=np.array([p.numel() for p in teacher.parameters()]), **after)

    # ---- process_label (G/util/utils.py:158-163) -------------------------------------------
    lab = torch.randint(0, 19, (2, 1, 5, 7), generator=gen).float()
    lab[0, 0, 0, :3] = 255.0
    lab[1, 0, 2, 2] = 19.0
    save("process_label", label=lab, onehot=ref.process_label(lab))

    # ---- a6 calculate_mean_vector, with and without labels (G/calc_centroids.py:120-145) ---
    cf = ref.Class_Features(numbers=19)
    n, d, h, w = 2, 32, 9, 11
    feat = torch.randn((n, d, h, w), generator=gen)
    out = 3.0 * torch.randn((n, 19, h, w), generator=gen)
    out[:, 5:] -= 4.0                                   # few dominant classes -> counts >= 5, some < 5
    labels = F.softmax(out, 1).argmax(1, keepdim=True).float()
    flip = torch.rand((n, 1, h, w), generator=gen) < 0.3
    labels[flip] = torch.randint(0, 19, (int(flip.sum()),), generator=gen).float()
    labels[0, 0, :2, :] = 255.0
    v0, i0 = cf.calculate_mean_vector(feat, out)
    v1, i1 = cf.calculate_mean_vector(feat, out, labels)
    v2, i2 = cf.calculate_mean_vector_by_output(feat, out)
    save("mean_vector", feat=feat, out=out, labels=labels,
         vec_nolabel=torch.stack(v0).reshape(len(i0), d), ids_nolabel=np.array(i0),
         vec_label=torch.stack(v1).reshape(len(i1), d), ids_label=np.array(i1),
         vec_by_output=torch.stack(v2).reshape(len(i2), d), ids_by_output=np.array(i2))

    # ---- a7 update_objective_SingleVector sequences (G/calc_centroids.py:147-164) ----------
    cf = ref.Class_Features(numbers=19)
    cf.objective_vectors = torch.zeros(19, 8)
    seq_ids, seq_vecs, seq_modes, seq_sm = [], [], [], []
    modes = ["mean", "moving_average"]
    for k in range(260):
        cid = int(torch.randint(0, 4, (1,), generator=gen))
        vec = torch.randn(8, 1, 1, generator=gen)
        if k % 37 == 5:
            vec = torch.zeros(8, 1, 1)                   # all-zero vector -> skipped (:148)
        mode = modes[k % 2]
        sm = bool(k % 3)
        seq_ids.append(cid), seq_vecs.append(vec.reshape(8)), seq_modes.append(k % 2), seq_sm.append(sm)
        cf.update_objective_SingleVector(cid, vec if k % 5 else vec.numpy(), mode, start_mean=sm)
    # clamp at 3000 (:156,161)
    cf.objective_vectors_num[3] = 2999.0
    for k in range(3):
        vec = torch.randn(8, 1, 1, generator=gen)
        seq_ids.append(3), seq_vecs.append(vec.reshape(8)), seq_modes.append(0), seq_sm.append(False)
        cf.update_objective_SingleVector(3, vec, "mean", start_mean=False)
    save("centroid_update", ids=np.array(seq_ids), vecs=torch.stack(seq_vecs), modes=np.array(seq_modes),
         start_mean=np.array(seq_sm), clamp_inject_at=260, clamp_inject_class=3, clamp_inject_value=2999.0,
         objective_vectors=cf.objective_vectors, objective_vectors_num=cf.objective_vectors_num)

    # ---- a5 distance / weight (G/calc_centroids.py:166-180) ---------------------------------
    cf = ref.Class_Features(numbers=19)
    for tag, d, hh, ww in (("proto_d256", 256, 5, 7), ("proto_d64", 64, 9, 13)):
        feat = torch.randn((2, d, hh, ww), generator=gen)
        cf.objective_vectors = 0.5 * torch.randn((19, d), generator=gen) + feat.mean()
        save(tag, feat=feat, centroids=cf.objective_vectors, dist=cf.feat_centroid_distance(feat),
             weight=cf.get_centroid_weight(feat), negdist=cf.get_centroid_distance(feat))

    # ---- a3 pseudo-label math: exec G/pseudolabel_generator.py:80-85 ------------------------
    block = _lines("pseudolabel_generator.py", 80, 85)
    for tag, sigma in (("pseudo_label", 3.0), ("pseudo_label_sat", 10.0)):
        z = sigma * torch.randn((1, 19, 24, 40), generator=gen)
        z_ds = sigma * torch.randn((1, 19, 24, 40), generator=gen)
        ns = {"torch": torch, "nn": nn, "np": np, "output": z.clone(), "output_ds": z_ds.clone()}
        exec(block, ns)
        save(tag, output=z, output_ds=z_ds, label=ns["label"], prob_hwc=ns["output"])
    # with the two bilinear up-samplings of :77-78 in front (CPU interpolation kernel)
    up = nn.Upsample(size=[64, 96], mode="bilinear", align_corners=True)
    lo, lo_ds = 3.0 * torch.randn((1, 19, 9, 13), generator=gen), 3.0 * torch.randn((1, 19, 5, 7), generator=gen)
    ns = {"torch": torch, "nn": nn, "np": np, "upsample_1024": up, "output": lo.clone(), "output_ds": lo_ds.clone()}
    exec(_lines("pseudolabel_generator.py", 77, 85), ns)
    save("pseudo_label_two_scale", logits=lo, logits_ds=lo_ds, size=np.array([64, 96]), label=ns["label"])

    # ---- f3 palette PNG: exec G/pseudolabel_generator.py:38-49 (palette, colorize_mask) + :91-92 -----------------
    import io
    from PIL import Image
    ns = {"np": np, "Image": Image}
    exec(_lines("pseudolabel_generator.py", 38, 49), ns)
    gen_png = torch.Generator().manual_seed(4242)                                             # own stream: other fixtures unchanged
    lab_f64 = torch.randint(0, 19, (24, 40), generator=gen_png).numpy().astype(np.float64)   # float64 staging array (:66)
    out_img = ns["colorize_mask"](np.asarray(lab_f64, dtype=np.uint8))                       # :92, :100
    bio = io.BytesIO()
    out_img.save(bio, format="PNG")
    back = Image.open(io.BytesIO(bio.getvalue()))
    save("palette_png", label=lab_f64, palette=np.array(ns["palette"]), mode=np.array([ord(ch) for ch in back.mode]),
         indices=np.array(back), png_palette=np.array(back.getpalette()))

    # ---- a2 ClassMix: exec self_training.py:259-275 (image only) and :306-325 (DACS) -------
    b, hh, ww = 3, 32, 48
    slabel = _blocky_labels(gen, b, hh, ww, 8)
    slabel[2] = 255                                                   # an all-ignore image
    img_a = torch.randn((b, 3, hh, ww), generator=gen).clamp(-1, 1)
    img_b = torch.randn((b, 3, hh, ww), generator=gen).clamp(-1, 1)
    img_a[0, 0, 0, 0] = -0.0
    tl = _blocky_labels(gen, b, hh, ww, 8)
    rnd = random.Random(77)
    ns = {"torch": torch, "random": rnd, "i_iter": 0, "slabelv": slabel.clone(),
          "rec_s2t": img_a.clone(), "sdatav_aug": img_b.clone()}
    exec(_lines("train_DiGA_gta2city_self_training.py", 259, 275), ns)
    rnd2 = random.Random(78)
    ns2 = {"torch": torch, "random": rnd2, "i_iter": 0, "slabelv": slabel.clone(), "tlabelv_pseudo": tl.clone(),
           "tdatav_aug": img_a.clone(), "sdatav": img_b.clone()}
    exec(_lines("train_DiGA_gta2city_self_training.py", 306, 325), ns2)
    save("classmix", slabel=slabel, a=img_a, b=img_b, tlabel=tl, seed_img=77, seed_dacs=78,
         mask_img=ns["mask"], mix_img=ns["sdatav_aug_crdomix"],
         mask_dacs=ns2["mask"], mix_dacs=ns2["cross_mix"], mixlabel_dacs=ns2["crossmix_label"])

    # ---- a4 consensus selection: exec self_training.py:298-304 ------------------------------
    cf = ref.Class_Features(numbers=19)
    d, fh, fw, oh, ow = 32, 9, 13, 64, 96
    t_feat = torch.randn((2, d, fh, fw), generator=gen)
    cf.objective_vectors = 0.5 * torch.randn((19, d), generator=gen)
    pp = _blocky_labels(gen, 2, oh, ow, 8)
    t_pred = 3.0 * torch.randn((2, 19, fh, fw), generator=gen)
    ns = {"torch": torch, "class_features": cf, "tlabelv_pseudo_prob": pp.clone(), "tdatav": None,
          "teacher": lambda x: (None, None, t_pred, t_feat),
          "upsample_tgt": nn.Upsample(size=[oh, ow], mode="bilinear", align_corners=True)}
    exec(_lines("train_DiGA_gta2city_self_training.py", 298, 304), ns)
    up_w = ns["feat_weights"]
    top2 = up_w.topk(2, dim=1).values
    save("consensus", t_feat=t_feat, centroids=cf.objective_vectors, pseudo_prob=pp, out_size=np.array([oh, ow]),
         weights_lowres=cf.get_centroid_weight(t_feat), tlabelv_pseudo=ns["tlabelv_pseudo"],
         feat_pseudo=ns["feat_pseudo"], top2_margin=(top2[:, 0] - top2[:, 1]))

    # ---- online centroid update block: exec self_training.py:327-341 -----------------------
    cf = ref.Class_Features(numbers=19)
    cf.objective_vectors = 0.5 * torch.randn((19, d), generator=gen)
    cf.objective_vectors_num = torch.full((19,), 150.0)
    before = cf.objective_vectors.clone()
    sl = _blocky_labels(gen, 2, oh, ow, 8)
    s_feat = torch.randn((2, d, fh, fw), generator=gen)
    s_pred = 3.0 * torch.randn((2, 19, fh, fw), generator=gen)
    s_pred[:, :4] += 5.0
    t_pred2 = t_pred.clone()
    t_pred2[:, 2:6] += 5.0
    ns = {"torch": torch, "F": F, "class_features": cf, "tlabelv_pseudo": ns["tlabelv_pseudo"].clone(),
          "t_feat_tea": t_feat, "t_pred_tea": t_pred2, "slabelv": sl, "s_feat_tea_aug": s_feat,
          "s_pred_tea_aug_raw": s_pred}
    exec(_lines("train_DiGA_gta2city_self_training.py", 327, 341), ns)
    save("online_update", tlabelv_pseudo=ns["tlabelv_pseudo"], t_feat=t_feat, t_pred=t_pred2, slabel=sl,
         s_feat=s_feat, s_pred=s_pred, centroids_before=before, num_before=np.full(19, 150.0, np.float32),
         centroids_after=cf.objective_vectors, num_after=cf.objective_vectors_num,
         ids_t=np.array(ns["ids_t"]), ids_s=np.array(ns["ids_s"]),
         newlabels_t=ns["newlabels_t"], newlabels_s=ns["newlabels_s"])

    # ---- f1 for the loss consumers: exec self_training.py:289 (teacher up-sampling), :344 (cross_pred_mix), :348-352
    #      (seg loss + KD on the shared student logits), :355-356, :382 (total loss), backward to the STRIDE-8 logits ----
    gen_up = torch.Generator().manual_seed(9090)                       # own stream: other fixtures unchanged
    bsz, c, fh, fw, oh, ow = 2, 19, 9, 13, 64, 96
    stu_low = (3.0 * torch.randn((2 * bsz, c, fh, fw), generator=gen_up)).requires_grad_(True)
    tea_low = 3.0 * torch.randn((2 * bsz, c, fh, fw), generator=gen_up)
    mix_low = (3.0 * torch.randn((bsz, c, fh, fw), generator=gen_up)).requires_grad_(True)
    sl = _blocky_labels(gen_up, bsz, oh, ow, 8)
    ml = _blocky_labels(gen_up, bsz, oh, ow, 8, p_ignore=0.3)
    up = nn.Upsample(size=[oh, ow], mode="bilinear", align_corners=True)
    ns = {"torch": torch, "upsample_src": up, "upsample_tgt": up, "seg_loss": ref.cross_entropy2d,
          "distillation_loss": ref.distillation_loss, "lambda_seg": 1.0, "lambda_distil": 0.25,
          "s_pred_cat_tea": tea_low.clone(), "s_pred_cat_stu": stu_low, "s_pred_stu": stu_low[:bsz],
          "cross_pred_mix": mix_low, "slabelv": sl, "crossmix_label": ml}
    for first, last in ((289, 289), (344, 344), (348, 349), (351, 352)):
        exec(_lines("train_DiGA_gta2city_self_training.py", first, last), ns)
    loss_src = ns["loss_semseg"].detach().clone()
    for first, last in ((355, 356), (382, 382), (385, 385)):
        exec(_lines("train_DiGA_gta2city_self_training.py", first, last), ns)
    save("losses_up", student_low=stu_low, teacher_low=tea_low, mix_low=mix_low, slabel=sl, mixlabel=ml,
         size=np.array([oh, ow]), lambda_seg=1.0, lambda_distil=0.25, kd_scale=0.5,
         loss_semseg_src=loss_src, loss_semseg=ns["loss_semseg"], loss_distil=ns["loss_s_distil"],
         total_loss=ns["total_loss"], grad_student_low=stu_low.grad, grad_mix_low=mix_low.grad)

    # ---- f5 evaluation confusion matrix: runningScore.update / get_scores (G/util/metrics.py:26-76) ----------------
    import contextlib
    import io as _io
    gen_m = torch.Generator().manual_seed(777)
    gt = _blocky_labels(gen_m, 3, 40, 56, 8).numpy()                                   # int64 with 255 = ignore
    pred = _blocky_labels(gen_m, 3, 40, 56, 4, p_ignore=0.0).numpy()
    pred[0] = np.where(gt[0] < 19, gt[0], pred[0])                                     # one well-predicted image
    rs = ref.runningScore(19)
    rs.update(gt[:2], pred[:2])
    rs.update(gt[2:], pred[2:])
    with contextlib.redirect_stdout(_io.StringIO()):                                   # :62-63 prints per-class IoU
        score, cls_iu = rs.get_scores()
    save("running_score", gt=gt, pred=pred, confusion_matrix=rs.confusion_matrix,
         overall_acc=score['Overall Acc: \t'], mean_acc=score['Mean Acc : \t'], fwavacc=score['FreqW Acc : \t'],
         mean_iu=score['Mean IoU : \t'], cls_iu=np.array([cls_iu[k] for k in range(19)]))

    # ---- f3 reader half: exec G/util/loader/CityLoader.py:92-96 (NEAREST resize) and :122-132 (id re-assignment) on two
    #      synthetic PNG-like images (a labelIds map with ids 0..33 and a 'P'-mode pseudo-label map with indices 0..18, 255)
    from PIL import Image

    class _NP:                                     # numpy with the `np.compat.long` the reference still uses (removed in numpy 2)
        compat = type("compat", (), {"long": np.int64})

        def __getattr__(self, name):
            return getattr(np, name)

    gen_l = torch.Generator().manual_seed(515)
    h0, w0, crop = 60, 104, (32, 56)
    ids = torch.randint(0, 34, (h0 // 4, w0 // 4), generator=gen_l).repeat_interleave(4, 0).repeat_interleave(4, 1).numpy().astype(np.uint8)
    pl = torch.randint(0, 19, (h0 // 4, w0 // 4), generator=gen_l).repeat_interleave(4, 0).repeat_interleave(4, 1).numpy().astype(np.uint8)
    pl[:8, :12] = 255
    pl[20:24, 40:44] = 19                                                              # an index outside the class range
    me = type("Self", (), {"crop_size": crop, "use_pseudo": True, "n_classes": 19,
                           "id_to_trainid": {7: 0, 8: 1, 11: 2, 12: 3, 13: 4, 17: 5, 19: 6, 20: 7, 21: 8, 22: 9, 23: 10, 24: 11,
                                             25: 12, 26: 13, 27: 14, 28: 15, 31: 16, 32: 17, 33: 18}})()
    ns = {"np": _NP(), "Image": Image, "self": me, "label": Image.fromarray(ids), "pseudo_label": Image.fromarray(pl),
          "image": Image.fromarray(np.zeros((h0, w0, 3), np.uint8))}
    exec(_lines("util/loader/CityLoader.py", 92, 96), ns)
    exec(_lines("util/loader/CityLoader.py", 122, 132), ns)
    save("city_loader_labels", ids=ids, pseudo=pl, crop_size=np.array(crop), label_copy=ns["label_copy"],
         pseudo_label_copy=ns["pseudo_label_copy"])

    # ---- OhemCrossEntropy (G/util/loss.py:65-122): low-resolution score (the loss up-samples it itself) and full-size score,
    #      threshold decided by thresh (many easy pixels) and by the min_kept-th order statistic ----------------------------
    gen_o = torch.Generator().manual_seed(8181)
    for tag, shape, hw, thres, kept, use_w in (("ohem_low_thresh", (2, 16, 9, 13), (64, 96), 0.7, 1000, False),
                                               ("ohem_low_minkept", (2, 19, 9, 13), (64, 96), 0.05, 3000, True),
                                               ("ohem_full", (1, 16, 24, 40), (24, 40), 0.7, 100, False)):
        sc = (3.0 * torch.randn(shape, generator=gen_o)).requires_grad_(True)
        tg = _blocky_labels(gen_o, shape[0], hw[0], hw[1], 8, n_cls=shape[1], p_ignore=0.15)
        wt = (0.5 + torch.rand(shape[1], generator=gen_o)) if use_w else None
        crit = ref.OhemCrossEntropy(ignore_label=255, thres=thres, min_kept=kept, weight=wt)
        loss = crit(sc, tg)
        (loss * 0.5).backward()
        save(tag, score=sc, target=tg, weight=(wt if use_w else torch.zeros(0)), thres=thres, min_kept=kept, upstream=0.5,
             loss=loss, grad=sc.grad)


if __name__ == "__main__":
    main()
