"""Pin the CPU oracle (oracle/diga_oracle.py) against outputs of the reference itself.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which runs the real
reference functions and the real reference script line ranges.  Bar: bit-exact (same torch ops on
the same CPU).  When /root/reference is present the live reference functions are checked too.
"""
import random

import numpy as np
import pytest
import torch

from oracle import diga_oracle as O
from oracle import ref_loader


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", ["kd_c19", "kd_c16_s025", "kd_saturated"])
def test_kd_matches_reference_bitwise(golden, name):
    g = golden(name)
    t, s = T(g["teacher"]), T(g["student"]).requires_grad_(True)
    loss = O.distillation_loss(t, s, float(g["scale"]))
    (loss * float(g["upstream"])).backward()
    assert np.array_equal(loss.detach().numpy(), g["loss"])
    assert np.array_equal(s.grad.numpy(), g["grad"])
    # closed-form gradient agrees to fp32 rounding
    cf = O.distillation_grad_closed_form(t, s.detach(), float(g["scale"]), float(g["upstream"])).numpy()
    assert np.abs(cf - g["grad"]).max() <= 1e-5 * np.abs(g["grad"]).max()


@pytest.mark.parametrize("name", ["ce_c19", "ce_weighted_sum"])
def test_cross_entropy2d_matches_reference_bitwise(golden, name):
    g = golden(name)
    x = T(g["input"]).requires_grad_(True)
    wt = T(g["weight"]) if g["weight"].size else None
    loss = O.cross_entropy2d(x, T(g["target"]), weight=wt, size_average=bool(g["size_average"]))
    (loss * float(g["upstream"])).backward()
    assert np.array_equal(loss.detach().numpy(), g["loss"])
    assert np.array_equal(x.grad.numpy(), g["grad"])
    assert (g["target"] < 0).any() and (g["target"] == 255).any()


def ema_nets(g):
    import torch.nn as nn

    def build(flat):
        net = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Conv2d(8, 5, 1), nn.Linear(7, 3))
        off = 0
        for p, n in zip(net.parameters(), g["sizes"]):
            p.data.copy_(T(flat[off:off + n]).reshape(p.shape))
            off += int(n)
        return net
    return build


def test_ema_teacher_update_matches_reference_bitwise(golden):
    g = golden("ema")
    build = ema_nets(g)
    cases = {"after_it0_stage0": (0, {}), "after_it7_stage0": (7, {}), "after_it5000_stage0": (5000, {}),
             "after_it3_stage0_mean": (3, {"stage0": False, "mean": True}), "after_it3_stage0": (3, {"stage0": False})}
    # key naming in the fixture: "_".join(kwargs) or "stage0"
    for key, (it, kw) in cases.items():
        fixture_key = f"after_it{it}_{'_'.join(k for k in kw) or 'stage0'}"
        teacher, student = build(g["teacher"]), build(g["student"])
        O.update_teacher_params(teacher, student, it, **kw)
        got = torch.cat([p.detach().reshape(-1) for p in teacher.parameters()]).numpy()
        assert np.array_equal(got, g[fixture_key]), key


def test_palette_png_format(golden):
    """f3: 'P'-mode PNG, palette index = trainId (pseudolabel_generator.py:38-49, 91-100) — oracle and product helper."""
    import io
    from PIL import Image
    g = golden("palette_png")
    from diga_b200.pseudolabel import CITYSCAPES_PALETTE, colorize_mask
    assert CITYSCAPES_PALETTE == g["palette"].tolist()
    for img in (O.colorize_mask(np.asarray(g["label"], dtype=np.uint8), g["palette"].tolist()),
                colorize_mask(np.asarray(g["label"], dtype=np.uint8))):
        bio = io.BytesIO()
        img.save(bio, format="PNG")
        back = Image.open(io.BytesIO(bio.getvalue()))
        assert back.mode == "".join(chr(c) for c in g["mode"]) == "P"
        assert np.array_equal(np.array(back), g["indices"])
        assert back.getpalette() == g["png_palette"].tolist()


def test_process_label(golden):
    g = golden("process_label")
    assert np.array_equal(O.process_label(T(g["label"])).numpy(), g["onehot"])


def test_mean_vector(golden):
    g = golden("mean_vector")
    cf = O.ClassFeaturesOracle(19, g["feat"].shape[1])
    for tag, labels in (("nolabel", None), ("label", T(g["labels"]))):
        vec, ids = cf.calculate_mean_vector(T(g["feat"]), T(g["out"]), labels)
        assert ids == g["ids_" + tag].tolist()
        assert vec[0].shape == (g["feat"].shape[1], 1, 1)
        assert np.array_equal(torch.stack(vec).reshape(len(ids), -1).numpy(), g["vec_" + tag])
    vec, ids = cf.calculate_mean_vector_by_output(T(g["feat"]), T(g["out"]))
    assert ids == g["ids_by_output"].tolist()
    assert np.array_equal(torch.stack(vec).reshape(len(ids), -1).numpy(), g["vec_by_output"])
    # the fixture exercises the "<5 pixels -> skipped" rule
    assert len(g["ids_nolabel"]) < 2 * 19 and len(g["ids_label"]) <= len(g["ids_nolabel"])


def replay_updates(cf, g):
    modes = ["mean", "moving_average"]
    at = int(g["clamp_inject_at"])
    for k, (cid, vec, m, sm) in enumerate(zip(g["ids"], g["vecs"], g["modes"], g["start_mean"])):
        if k == at:
            cf.objective_vectors_num[int(g["clamp_inject_class"])] = float(g["clamp_inject_value"])
        cf.update_objective_SingleVector(int(cid), T(vec).reshape(-1, 1, 1), modes[int(m)], start_mean=bool(sm))


def test_centroid_update_sequence(golden):
    g = golden("centroid_update")
    cf = O.ClassFeaturesOracle(19, g["vecs"].shape[1])
    replay_updates(cf, g)
    assert np.array_equal(cf.objective_vectors.numpy(), g["objective_vectors"])
    assert np.array_equal(cf.objective_vectors_num.numpy(), g["objective_vectors_num"])
    assert g["objective_vectors_num"][3] == 3000.0
    with pytest.raises(NotImplementedError):
        cf.update_objective_SingleVector(0, torch.ones(8, 1, 1), "median", start_mean=False)


@pytest.mark.parametrize("name", ["proto_d256", "proto_d64"])
def test_proto_distance(golden, name):
    g = golden(name)
    cf = O.ClassFeaturesOracle(19, g["feat"].shape[1])
    cf.objective_vectors = T(g["centroids"])
    assert np.array_equal(cf.feat_centroid_distance(T(g["feat"])).numpy(), g["dist"])
    assert np.array_equal(cf.get_centroid_weight(T(g["feat"])).numpy(), g["weight"])
    assert np.array_equal(cf.get_centroid_distance(T(g["feat"])).numpy(), g["negdist"])


@pytest.mark.parametrize("name", ["pseudo_label", "pseudo_label_sat"])
def test_pseudo_label(golden, name):
    g = golden(name)
    label, conf = O.pseudo_label_from_logits(T(g["output"]), T(g["output_ds"]))
    assert np.array_equal(label, g["label"])
    assert np.array_equal(conf, g["prob_hwc"].max(axis=2))


def test_pseudo_label_two_scale(golden):
    g = golden("pseudo_label_two_scale")
    label, _ = O.pseudo_label_two_scale(T(g["logits"]), T(g["logits_ds"]), g["size"].tolist())
    assert np.array_equal(label, g["label"])
    assert O.pseudo_label_to_uint8(label).dtype == np.uint8


def test_classmix(golden):
    g = golden("classmix")
    sl, a, b, tl = T(g["slabel"]), T(g["a"]), T(g["b"]), T(g["tlabel"])
    mask, mix = O.classmix(sl, a, b, rng=random.Random(int(g["seed_img"])))
    assert np.array_equal(mask.numpy(), g["mask_img"])
    assert np.array_equal(mix.numpy().view(np.uint32), g["mix_img"].view(np.uint32))
    mask, mix, ml = O.classmix(sl, a, b, tl, rng=random.Random(int(g["seed_dacs"])))
    assert np.array_equal(mask.numpy(), g["mask_dacs"])
    assert np.array_equal(mix.numpy().view(np.uint32), g["mix_dacs"].view(np.uint32))
    assert ml.dtype == torch.int64 and np.array_equal(ml.numpy(), g["mixlabel_dacs"])
    # every-label-255 batch: the reference never creates the mix tensor
    m2, mix2 = O.classmix(torch.full_like(sl, 255), a, b, rng=random.Random(0))
    assert mix2 is None and bool((m2 == 1).all())


def test_consensus(golden):
    g = golden("consensus")
    cf = O.ClassFeaturesOracle(19, g["t_feat"].shape[1])
    cf.objective_vectors = T(g["centroids"])
    w = cf.get_centroid_weight(T(g["t_feat"]))
    assert np.array_equal(w.numpy(), g["weights_lowres"])
    kept, fp = O.consensus_select(T(g["pseudo_prob"]), w, g["out_size"].tolist())
    assert np.array_equal(kept.numpy(), g["tlabelv_pseudo"])
    assert np.array_equal(fp.numpy(), g["feat_pseudo"])


def test_online_update_block(golden):
    """self_training.py:327-341 — nearest label down-sampling, label-gated means, EMA updates."""
    g = golden("online_update")
    d = g["t_feat"].shape[1]
    cf = O.ClassFeaturesOracle(19, d)
    cf.objective_vectors = T(g["centroids_before"]).clone()
    cf.objective_vectors_num = T(g["num_before"]).clone()
    for lab, feat, pred, key in ((g["tlabelv_pseudo"], g["t_feat"], g["t_pred"], "t"),
                                 (g["slabel"], g["s_feat"], g["s_pred"], "s")):
        nl = O.nearest_labels_to_feature_grid(T(lab), feat.shape[2:])
        assert np.array_equal(nl.numpy(), g["newlabels_" + key])
        vec, ids = cf.calculate_mean_vector(T(feat), T(pred), nl)
        assert ids == g["ids_" + key].tolist()
        for v, i in zip(vec, ids):
            cf.update_objective_SingleVector(i, v.detach(), start_mean=False)
    assert np.array_equal(cf.objective_vectors.numpy(), g["centroids_after"])
    assert np.array_equal(cf.objective_vectors_num.numpy(), g["num_after"])


@pytest.mark.skipif(not ref_loader.available("G"), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference():
    ref = ref_loader.load("G")
    gen = torch.Generator().manual_seed(5)
    t, s = 3 * torch.randn((6, 19, 12, 20), generator=gen), 3 * torch.randn((6, 19, 12, 20), generator=gen)
    assert torch.equal(ref.distillation_loss(t, s), O.distillation_loss(t, s))
    feat, out = torch.randn((3, 48, 17, 19), generator=gen), 3 * torch.randn((3, 19, 17, 19), generator=gen)
    rcf, ocf = ref.Class_Features(19), O.ClassFeaturesOracle(19, 48)
    rcf.objective_vectors = torch.zeros(19, 48)
    for _ in range(2):
        rv, ri = rcf.calculate_mean_vector(feat, out)
        ov, oi = ocf.calculate_mean_vector(feat, out)
        assert ri == oi and all(torch.equal(a, b) for a, b in zip(rv, ov))
        for v, i in zip(rv, ri):
            rcf.update_objective_SingleVector(i, v.detach().cpu().numpy(), "mean")
            ocf.update_objective_SingleVector(i, v.detach().cpu().numpy(), "mean")
    assert torch.equal(rcf.objective_vectors, ocf.objective_vectors)
    assert torch.equal(rcf.get_centroid_weight(feat), ocf.get_centroid_weight(feat))


def test_losses_upsampled_match_reference_bitwise(golden):
    """self_training.py:289,:344,:348-356,:382-385 — both seg losses + KD from the stride-8 logits, with backward."""
    g = golden("losses_up")
    stu = T(g["student_low"]).requires_grad_(True)
    mix = T(g["mix_low"]).requires_grad_(True)
    tea, sl, ml = T(g["teacher_low"]), T(g["slabel"]), T(g["mixlabel"])
    size = tuple(int(v) for v in g["size"])
    loss_src, loss_kd = O.seg_distillation_losses_upsampled(tea, stu, sl, float(g["kd_scale"]))
    loss_seg = loss_src + O.cross_entropy2d_upsampled(mix, ml)
    total = float(g["lambda_seg"]) * loss_seg + float(g["lambda_distil"]) * loss_kd
    total.backward()
    assert np.array_equal(loss_src.detach().numpy(), g["loss_semseg_src"])
    assert np.array_equal(loss_seg.detach().numpy(), g["loss_semseg"])
    assert np.array_equal(loss_kd.detach().numpy(), g["loss_distil"])
    assert np.array_equal(total.detach().numpy(), g["total_loss"])
    assert np.array_equal(stu.grad.numpy(), g["grad_student_low"])
    assert np.array_equal(mix.grad.numpy(), g["grad_mix_low"])
    # the stand-alone forms are the same op chains
    assert np.array_equal(O.distillation_loss_upsampled(tea, stu.detach(), size, float(g["kd_scale"])).numpy(), g["loss_distil"])


def test_running_score_matches_reference(golden):
    """runningScore (G/util/metrics.py:26-76): confusion matrix and every score bit for bit; live reference when present."""
    g = golden("running_score")
    rs = O.RunningScoreOracle(19)
    rs.update(g["gt"][:2], g["pred"][:2])
    rs.update(g["gt"][2:], g["pred"][2:])
    score, cls_iu = rs.get_scores()
    assert np.array_equal(rs.confusion_matrix, g["confusion_matrix"])
    assert score['Overall Acc: \t'] == g["overall_acc"] and score['Mean Acc : \t'] == g["mean_acc"]
    assert score['FreqW Acc : \t'] == g["fwavacc"] and score['Mean IoU : \t'] == g["mean_iu"]
    assert np.array_equal(np.array([cls_iu[k] for k in range(19)]), g["cls_iu"], equal_nan=True)
    assert (g["gt"] == 255).any()
    rs.reset()
    assert rs.confusion_matrix.sum() == 0
    if ref_loader.available("G"):
        import contextlib
        import io
        live = ref_loader.load("G").runningScore(19)
        live.update(g["gt"], g["pred"])
        with contextlib.redirect_stdout(io.StringIO()):
            lscore, _ = live.get_scores()
        assert np.array_equal(live.confusion_matrix, g["confusion_matrix"]) and lscore['Mean IoU : \t'] == g["mean_iu"]


def test_city_loader_labels_match_reference(golden):
    """CityLoader.py:92-96, :122-132 (label / pseudo-label maps): the oracle restatement and the host-side NEAREST index
    table the CUDA reader uses, both against what the reference's own statements produced."""
    from PIL import Image
    g = golden("city_loader_labels")
    crop = tuple(int(v) for v in g["crop_size"])
    lab, pl = O.city_loader_labels(Image.fromarray(g["ids"]), Image.fromarray(g["pseudo"]), crop)
    assert np.array_equal(lab, g["label_copy"]) and np.array_equal(pl, g["pseudo_label_copy"])
    assert lab.dtype == np.int64 and (g["pseudo_label_copy"] == 255).any() and (g["pseudo"] == 19).any()
    # the index tables of diga_b200.util.labels (pure host code, no CUDA) reproduce Pillow's NEAREST transform
    from diga_b200.util.labels import pil_nearest_table as table      # host code; the package loads without a GPU
    rng = np.random.default_rng(3)
    for h0, w0, hh, ww in [(1024, 2048, 512, 1024), (1024, 2048, 512, 896), (1052, 1914, 512, 896), (60, 104, 32, 56), (7, 9, 20, 31)] + \
            [tuple(int(v) for v in rng.integers(1, 200, 4)) for _ in range(60)]:
        img = rng.integers(0, 256, (h0, w0), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ww, hh), Image.NEAREST))
        assert np.array_equal(img[table(h0, hh)][:, table(w0, ww)], ref), (h0, w0, hh, ww)


@pytest.mark.parametrize("name", ["ohem_low_thresh", "ohem_low_minkept", "ohem_full"])
def test_ohem_matches_reference_bitwise(golden, name):
    """OhemCrossEntropy (G/util/loss.py:65-122), threshold from `thres` and from the min_kept-th order statistic."""
    g = golden(name)
    sc = T(g["score"]).requires_grad_(True)
    wt = T(g["weight"]) if g["weight"].size else None
    crit = O.OhemCrossEntropyOracle(255, float(g["thres"]), int(g["min_kept"]), wt)
    loss = crit(sc, T(g["target"]))
    (loss * float(g["upstream"])).backward()
    assert np.array_equal(loss.detach().numpy(), g["loss"]) and np.array_equal(sc.grad.numpy(), g["grad"])
