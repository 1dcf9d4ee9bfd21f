"""GPU parity tests: the CUDA path (through the Python mirror -> ctypes -> C ABI -> sm_100a kernels) against
the oracle, on the committed golden fixtures and on seeded synthetic inputs.

Bars (BASELINE.json north_star): labels / masks / int64 maps bit-exact; losses, gradients, centroids,
distances within 1e-5 relative in fp32.  "Relative" is norm-wise for tensors with cancellation
(SURVEY.md §7): max|a-b| <= 1e-5 * max|b|; scalars use plain relative error.

For paths that run through bilinear interpolation the oracle is evaluated ON THE GPU (same torch op chain
as the reference would run there): torch's CPU interpolation differs from its CUDA kernel by 1 ulp on a
quarter of the values, so only the GPU eager chain can arbitrate bit-exact labels.
"""
import random

import numpy as np
import pytest
import torch

from oracle import diga_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def dev():
    return torch.device("cuda", 0)


def T(a, d=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(d) if d is not None else t


def assert_normwise(got, want, rtol=RTOL, what=""):
    got, want = got.detach().double().cpu(), torch.as_tensor(want).detach().double().cpu()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= rtol * max(scale, 1e-30), f"{what}: max|diff| {err:.3e} > {rtol} * max|ref| {scale:.3e}"


def assert_same_bits(got, want, what=""):
    """Bit-exact fp32 comparison (distinguishes -0.0 from 0.0); NaNs must coincide, their payload may differ
    (x86 produces the negative quiet NaN for inf*0, the GPU the canonical 0x7fffffff)."""
    got, want = np.asarray(got, dtype=np.float32), np.asarray(want, dtype=np.float32)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN positions differ"
    assert np.array_equal(got.view(np.uint32)[~nan_g], want.view(np.uint32)[~nan_w]), f"{what}: bits differ"


def assert_rel(got, want, rtol=RTOL, what=""):
    got, want = float(got), float(want)
    assert abs(got - want) <= rtol * abs(want), f"{what}: {got!r} vs {want!r}"


@pytest.fixture(scope="module")
def D():
    import diga_b200
    return diga_b200


# ------------------------------------------------------------------------------------------------ a1 KD
@pytest.mark.parametrize("name", ["kd_c19", "kd_c16_s025", "kd_saturated"])
def test_kd_golden(D, golden, name):
    g = golden(name)
    t = T(g["teacher"], dev())
    s = T(g["student"], dev()).requires_grad_(True)
    loss = D.distillation_loss(t, s, float(g["scale"]))
    assert loss.dim() == 0 and loss.dtype == torch.float32
    (loss * float(g["upstream"])).backward()
    assert_rel(loss.item(), g["loss"], what="loss")
    assert_normwise(s.grad, g["grad"], what="grad")
    assert t.grad is None


@pytest.mark.parametrize("shape,scale", [((4, 19, 64, 128), 0.5), ((2, 19, 7, 9), 0.5), ((6, 16, 33, 65), 0.25),
                                         ((2, 7, 16, 16), 0.5), ((2, 32, 8, 10), 1.0), ((2, 1, 4, 4), 0.5)])
def test_kd_vs_oracle(D, shape, scale):
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    t, s = 3 * torch.randn(shape, generator=g), 3 * torch.randn(shape, generator=g)
    so = s.clone().requires_grad_(True)
    lo = O.distillation_loss(t, so, scale)
    (lo * 0.3).backward()
    sg = s.to(dev()).requires_grad_(True)
    lg = D.distillation_loss(t.to(dev()), sg, scale)
    (lg * 0.3).backward()
    assert_rel(lg.item(), lo.item(), what="loss")
    assert_normwise(sg.grad, so.grad, what="grad")
    # closed form in fp64
    assert_normwise(sg.grad, O.distillation_grad_closed_form(t, s, scale, 0.3), what="grad vs closed form")
    # single-pass variant
    l2, g2 = D.distillation_loss_and_grad(t.to(dev()), s.to(dev()), scale, 0.3)
    assert_rel(l2.item(), lo.item(), what="fused loss")
    assert_normwise(g2, so.grad, what="fused grad")


def test_kd_full_size_config2(D):
    """BASELINE config 2: [8,19,512,1024]; oracle = the reference op chain evaluated on the GPU."""
    g = torch.Generator(device=dev()).manual_seed(1234)
    shape = (8, 19, 512, 1024)
    t = 3 * torch.randn(shape, generator=g, device=dev())
    s = 3 * torch.randn(shape, generator=g, device=dev())
    so = s.clone().requires_grad_(True)
    lo = O.distillation_loss(t, so, 0.5)
    (lo * 0.25).backward()
    sg = s.clone().requires_grad_(True)
    lg = D.distillation_loss(t, sg, 0.5)
    (lg * 0.25).backward()
    assert_rel(lg.item(), lo.item(), what="loss")
    assert_normwise(sg.grad, so.grad, what="grad")
    # properties: deterministic, gradient of each pixel sums to zero over classes, linear in upstream
    lg2 = D.distillation_loss(t, s, 0.5)
    assert lg2.item() == lg.item()
    assert sg.grad.sum(1).abs().max().item() <= 1e-5 * sg.grad.abs().max().item()
    _, g1 = D.distillation_loss_and_grad(t, s, 0.5, 1.0)
    assert_normwise(g1 * 0.25, sg.grad, rtol=1e-6, what="upstream linearity")


def test_kd_errors(D):
    x = torch.zeros(3, 19, 4, 4, device=dev())
    with pytest.raises(ValueError):
        D.distillation_loss(x, x)
    with pytest.raises(RuntimeError):
        D.distillation_loss(torch.zeros(2, 19, 4, 4), torch.zeros(2, 19, 4, 4))
    with pytest.raises(RuntimeError):          # C > 32 is rejected by the C ABI, not silently mis-computed
        D.distillation_loss(torch.zeros(2, 40, 4, 4, device=dev()), torch.zeros(2, 40, 4, 4, device=dev()))


# ------------------------------------------------------------------------------------------------ a3 pseudo-label
def check_labels_modulo_softmax_ties(lab_gpu, logits_fused_cpu, lab_ref):
    """Labels must be identical except where the reference's softmax values tie exactly (float rounding),
    the one divergence SURVEY.md §7 permits; returns the number of such pixels."""
    lab_gpu = np.asarray(lab_gpu).astype(np.int64)
    diff = lab_gpu != lab_ref
    if not diff.any():
        return 0
    prob = torch.softmax(logits_fused_cpu, dim=1)[0].numpy()          # [C,H,W]
    ys, xs = np.nonzero(diff)
    for y, x in zip(ys, xs):
        assert prob[lab_gpu[y, x], y, x] == prob[lab_ref[y, x], y, x], (
            f"label mismatch at {(y, x)} not explained by a softmax tie: gpu {lab_gpu[y, x]} ref {lab_ref[y, x]}")
    return int(diff.sum())


@pytest.mark.parametrize("name", ["pseudo_label", "pseudo_label_sat"])
def test_pseudo_label_golden(D, golden, name):
    g = golden(name)
    z, zd = T(g["output"], dev()), T(g["output_ds"], dev())
    lab, conf, lab64 = D.pseudo_label(z, zd, want_int64=True)
    assert lab.dtype == torch.uint8 and lab64.dtype == torch.int64
    fused = torch.max(T(g["output_ds"]), T(g["output"]))
    check_labels_modulo_softmax_ties(lab[0].cpu().numpy(), fused, g["label"])
    assert torch.equal(lab.long(), lab64)
    assert_normwise(conf[0], g["prob_hwc"].max(axis=2), what="confidence")


@pytest.mark.parametrize("shape,two", [((1, 19, 64, 128), True), ((2, 19, 31, 37), True), ((3, 16, 16, 24), False),
                                       ((1, 5, 9, 9), True), ((1, 19, 6, 10), False)])
def test_pseudo_label_vs_oracle(D, shape, two):
    g = torch.Generator().manual_seed(11)
    z = 3 * torch.randn(shape, generator=g)
    zd = 3 * torch.randn(shape, generator=g) if two else None
    lab, conf = D.pseudo_label(z.to(dev()), zd.to(dev()) if two else None)
    for i in range(shape[0]):
        lo, co = O.pseudo_label_from_logits(z[i:i + 1], zd[i:i + 1] if two else None)
        fused = torch.max(zd[i:i + 1], z[i:i + 1]) if two else z[i:i + 1]
        check_labels_modulo_softmax_ties(lab[i].cpu().numpy(), fused, lo)
        assert_normwise(conf[i], co, what="confidence")


def test_pseudo_label_full_size(D):
    """2048x1024 two-scale: identical to the reference op chain run on the GPU (incl. first-index ties)."""
    g = torch.Generator(device=dev()).manual_seed(7)
    z = 3 * torch.randn((1, 19, 1024, 2048), generator=g, device=dev())
    zd = 3 * torch.randn((1, 19, 1024, 2048), generator=g, device=dev())
    zd[0, :, :8, :] = z[0, :, :8, :]                 # exact logit ties between scales
    z[0, 3, 8:16, :] = z[0, 11, 8:16, :] = 50.0      # exact class ties -> first index must win
    lab, conf = D.pseudo_label(z, zd)
    prob = torch.softmax(torch.max(zd, z), dim=1)
    ref_conf, ref_lab = prob.max(dim=1)
    mism = (lab.long() != ref_lab)
    # permitted only at exact softmax ties
    p_gpu = prob.gather(1, lab.long().unsqueeze(1)).squeeze(1)
    assert bool((p_gpu[mism] == ref_conf[mism]).all())
    assert int(mism[:, 16:].sum()) <= 8
    assert bool((lab[0, 8:16, :] == 3).all())
    assert_normwise(conf, ref_conf, what="confidence")
    # idempotence / determinism
    lab2, _ = D.pseudo_label(z, zd, want_conf=False)
    assert torch.equal(lab, lab2)


@pytest.mark.parametrize("lo,hi", [((9, 13), (64, 96)), ((65, 129), (512, 1024)), ((33, 65), (33, 65)), ((5, 7), (3, 4)),
                                   ((1, 1), (4, 4)), ((129, 257), (1024, 2048))])
def test_upsample_bitwise_vs_torch_cuda(D, lo, hi):
    from diga_b200.selection import upsample_bilinear
    g = torch.Generator(device=dev()).manual_seed(3)
    x = torch.randn((2, 19, *lo), generator=g, device=dev())
    got = upsample_bilinear(x, hi)
    want = O.upsample_bilinear_ac(x, hi)
    assert torch.equal(got, want), f"{int((got != want).sum())} of {got.numel()} values differ from torch's CUDA kernel"


def test_pseudo_label_two_scale_fused(D):
    g = torch.Generator(device=dev()).manual_seed(5)
    z1 = 3 * torch.randn((2, 19, 33, 65), generator=g, device=dev())
    z2 = 3 * torch.randn((2, 19, 17, 33), generator=g, device=dev())
    size = (256, 512)
    lab, conf = D.pseudo_label_two_scale(z1, z2, size)
    up = torch.max(O.upsample_bilinear_ac(z2, size), O.upsample_bilinear_ac(z1, size))
    prob = torch.softmax(up, dim=1)
    ref_conf, ref_lab = prob.max(dim=1)
    mism = lab.long() != ref_lab
    p_gpu = prob.gather(1, lab.long().unsqueeze(1)).squeeze(1)
    assert bool((p_gpu[mism] == ref_conf[mism]).all()) and int(mism.sum()) <= 4
    assert_normwise(conf, ref_conf, what="confidence")
    lab1, _ = D.pseudo_label_two_scale(z1, None, size)
    assert bool((lab1.long() == O.upsample_bilinear_ac(z1, size).argmax(1)).all())


def test_pseudo_label_two_scale_golden(D, golden):
    """Golden made with torch's CPU interpolation: 1-ulp differences may flip a near-tie; enumerate them."""
    g = golden("pseudo_label_two_scale")
    size = g["size"].tolist()
    lab, _ = D.pseudo_label_two_scale(T(g["logits"], dev()), T(g["logits_ds"], dev()), size)
    lab = lab[0].cpu().numpy().astype(np.int64)
    up = torch.max(O.upsample_bilinear_ac(T(g["logits_ds"]), size), O.upsample_bilinear_ac(T(g["logits"]), size))[0]
    for y, x in zip(*np.nonzero(lab != g["label"])):
        top2 = up[:, y, x].topk(2).values
        assert (top2[0] - top2[1]).item() <= 1e-5 * abs(top2[0].item()), "mismatch with a clear margin"
    assert (lab != g["label"]).sum() <= 2


# ------------------------------------------------------------------------------------------------ a2 ClassMix
def test_classmix_golden(D, golden):
    g = golden("classmix")
    sl, a, b, tl = (T(g[k], dev()) for k in ("slabel", "a", "b", "tlabel"))
    mask, mix = D.classmix(sl, a, b, rng=random.Random(int(g["seed_img"])))
    assert np.array_equal(mask.cpu().numpy(), g["mask_img"])
    assert np.array_equal(mix.cpu().numpy().view(np.uint32), g["mix_img"].view(np.uint32))     # incl. -0.0
    mask, mix, ml = D.classmix(sl, a, b, tl, rng=random.Random(int(g["seed_dacs"])))
    assert np.array_equal(mask.cpu().numpy(), g["mask_dacs"])
    assert np.array_equal(mix.cpu().numpy().view(np.uint32), g["mix_dacs"].view(np.uint32))
    assert ml.dtype == torch.int64 and np.array_equal(ml.cpu().numpy(), g["mixlabel_dacs"])


@pytest.mark.parametrize("b,h,w,block", [(8, 512, 1024, 32), (3, 37, 53, 8), (1, 16, 18, 4)])
def test_classmix_vs_oracle(D, b, h, w, block):
    from diga_b200 import synthetic as S
    g = S.gen(99)
    sl = S.block_labels(b, h, w, g, block)
    tl = S.perturb_labels(sl, g, block)
    a, bb = S.images((b, 3, h, w), g), S.images((b, 3, h, w), g)
    a[0, 0, 0, 0] = float("inf")
    bb[0, 1, 0, 1] = -0.0
    mo, xo, lo = O.classmix(sl, a, bb, tl, rng=random.Random(5))
    mg, xg, lg = D.classmix(sl.to(dev()), a.to(dev()), bb.to(dev()), tl.to(dev()), rng=random.Random(5))
    assert torch.equal(mg.cpu(), mo)
    assert_same_bits(xg.cpu().numpy(), xo.numpy(), "mix")
    assert bool(torch.isnan(xg[0, 0, 0, 0]))            # inf * 0 propagates exactly as in the reference
    assert torch.equal(lg.cpu(), lo)
    # presence bitmap == torch.unique
    from diga_b200.classmix import present_classes
    assert present_classes(sl.to(dev())) == [torch.unique(sl[i]).tolist() for i in range(b)]


def test_classmix_shared_async_presence(D):
    """One presence pass (present_classes_async) shared by both ClassMix blocks of a step == two independent calls,
    the seeded `random` stream consumed in the same order (self_training.py:265-266 then :310-311)."""
    g = torch.Generator(device=dev()).manual_seed(5)
    b, h, w = 3, 48, 80
    from diga_b200 import synthetic as S
    sl = S.block_labels(b, h, w, g, 8)
    tl = S.perturb_labels(sl, g, 8)
    xa, xb, xc, xd = (S.images((b, 3, h, w), g) for _ in range(4))
    r1 = random.Random(31)
    m1, mix1 = D.classmix(sl, xa, xb, rng=r1)
    m2, mix2, lab2 = D.classmix(sl, xc, xd, tl, rng=r1)
    r2 = random.Random(31)
    pres = D.present_classes_async(sl)
    _ = D.pseudo_label(S.logits((1, 19, 16, 16), g))                       # unrelated work queued behind the request
    n1, nix1 = D.classmix(sl, xa, xb, rng=r2, present=pres)
    n2, nix2, nab2 = D.classmix(sl, xc, xd, tl, rng=r2, present=pres)
    assert torch.equal(m1, n1) and torch.equal(mix1, nix1)
    assert torch.equal(m2, n2) and torch.equal(mix2, nix2) and torch.equal(lab2, nab2)


def test_classmix_all_ignore_and_bad_labels(D):
    sl = torch.full((2, 8, 8), 255, dtype=torch.int64, device=dev())
    a = torch.randn(2, 3, 8, 8, device=dev())
    mask, mix = D.classmix(sl, a, a.clone(), rng=random.Random(0))
    assert mix is None and bool((mask == 1).all())
    sl[0, 0, 0] = 300
    with pytest.raises(ValueError):
        D.classmix(sl, a, a.clone(), rng=random.Random(0))


# ------------------------------------------------------------------------------------------------ a6 / a7 centroids
def test_process_label_golden(D, golden):
    g = golden("process_label")
    assert np.array_equal(D.process_label(T(g["label"], dev())).cpu().numpy(), g["onehot"])


def test_mean_vector_golden(D, golden):
    g = golden("mean_vector")
    d = g["feat"].shape[1]
    cf = D.Class_Features(19, d)
    feat, out, labels = T(g["feat"], dev()), T(g["out"], dev()), T(g["labels"], dev())
    for tag, lab in (("nolabel", None), ("label", labels)):
        vec, ids = cf.calculate_mean_vector(feat, out, lab)
        assert ids == g["ids_" + tag].tolist()
        assert all(isinstance(i, int) for i in ids) and vec[0].shape == (d, 1, 1)
        got = torch.stack(vec).reshape(len(ids), d)
        for k in range(len(ids)):                                   # per class vector
            assert_normwise(got[k], g["vec_" + tag][k], what=f"{tag} vector {k}")
    vec, ids = cf.calculate_mean_vector_by_output(feat, out)
    assert ids == g["ids_by_output"].tolist()


def test_centroid_update_golden_bit_exact(D, golden):
    """Per-vector API, arithmetic mirrored op for op -> bit-exact against the reference sequence."""
    g = golden("centroid_update")
    cf = D.Class_Features(19, g["vecs"].shape[1])
    modes = ["mean", "moving_average"]
    at = int(g["clamp_inject_at"])
    for k, (cid, vec, m, sm) in enumerate(zip(g["ids"], g["vecs"], g["modes"], g["start_mean"])):
        if k == at:
            num = cf.objective_vectors_num.clone()
            num[int(g["clamp_inject_class"])] = float(g["clamp_inject_value"])
            cf.objective_vectors_num = num
        v = T(vec).reshape(-1, 1, 1)
        v = v.numpy() if k % 5 == 0 else (v.to(dev()) if k % 2 else v)         # numpy, CUDA and CPU tensors
        cf.update_objective_SingleVector(int(cid), v, modes[int(m)], start_mean=bool(sm))
    assert np.array_equal(cf.objective_vectors_num.cpu().numpy(), g["objective_vectors_num"])
    assert np.array_equal(cf.objective_vectors.cpu().numpy(), g["objective_vectors"])
    with pytest.raises(NotImplementedError):
        cf.update_objective_SingleVector(0, torch.ones(8), "median", start_mean=False)


def test_online_update_golden(D, golden):
    """self_training.py:327-341 — label-gated means + EMA, via the reference-shaped API and the fused path."""
    g = golden("online_update")
    d = g["t_feat"].shape[1]
    for fused in (False, True):
        cf = D.Class_Features(19, d)
        cf.objective_vectors = T(g["centroids_before"])
        cf.objective_vectors_num = T(g["num_before"])
        for lab, feat, pred, key in ((g["tlabelv_pseudo"], g["t_feat"], g["t_pred"], "t"),
                                     (g["slabel"], g["s_feat"], g["s_pred"], "s")):
            nl = O.nearest_labels_to_feature_grid(T(lab, dev()), feat.shape[2:])
            assert np.array_equal(nl.cpu().numpy(), g["newlabels_" + key])
            if fused:
                cf.update_from_features(T(feat, dev()), T(pred, dev()), nl, start_mean=False)
            else:
                vec, ids = cf.calculate_mean_vector(T(feat, dev()), T(pred, dev()), nl)
                assert ids == g["ids_" + key].tolist()
                for v, i in zip(vec, ids):
                    cf.update_objective_SingleVector(i, v.detach(), start_mean=False)
        assert np.array_equal(cf.objective_vectors_num.cpu().numpy(), g["num_after"])
        for c in range(19):
            assert_normwise(cf.objective_vectors[c], g["centroids_after"][c], what=f"centroid {c}")


@pytest.mark.parametrize("n,d,h,w,c,labels", [(1, 2048, 33, 65, 19, False), (2, 256, 65, 129, 19, True),
                                              (3, 512, 16, 16, 16, True), (1, 30, 5, 5, 19, False),
                                              (2, 64, 3, 2, 19, False)])
def test_centroid_pass_vs_oracle(D, n, d, h, w, c, labels):
    from diga_b200 import synthetic as S
    g = S.gen(4321)
    ocf, gcf = O.ClassFeaturesOracle(c, d), D.Class_Features(c, d)
    for it in range(3):
        feat = S.features((n, d, h, w), g)
        out = S.logits((n, c, h, w), g)
        out[:, : max(2, c // 3)] += 2.0
        lab = None
        if labels:
            lab = out.argmax(1, keepdim=True).float()
            flip = torch.rand((n, 1, h, w), generator=g) < 0.3
            lab[flip] = 255.0
        vec, ids = ocf.calculate_mean_vector(feat, out, lab)
        for v, i in zip(vec, ids):
            ocf.update_objective_SingleVector(i, v.detach().cpu().numpy(), "mean")
        gvec, gids = gcf.calculate_mean_vector(feat.to(dev()), out.to(dev()), None if lab is None else lab.to(dev()))
        assert gids == ids
        for a, b in zip(gvec, vec):
            assert_normwise(a.reshape(-1), b.reshape(-1), what="mean vector")
        gcf.update_from_features(feat.to(dev()), out.to(dev()), None if lab is None else lab.to(dev()), "mean")
    assert torch.equal(gcf.objective_vectors_num.cpu(), ocf.objective_vectors_num)
    for k in range(c):
        assert_normwise(gcf.objective_vectors[k], ocf.objective_vectors[k], what=f"centroid {k}")


def test_centroid_accum_variants_and_determinism(D):
    """Every launch shape of the accumulation kernel gives the same sums, run to run bit-identical."""
    from diga_b200 import _lib as L, synthetic as S
    g = S.gen(8, "cuda")
    feat = S.features((2, 256, 65, 129), g)
    out = S.logits((2, 19, 65, 129), g)
    ref = None
    try:
        for variant in range(21):
            L.set_tunable("accum_variant", variant)
            cf = D.Class_Features(19, 256)
            v1, s1, ok1 = cf._masked_means(feat, out, None)
            v2, _, _ = cf._masked_means(feat, out, None)
            assert torch.equal(v1, v2)
            if ref is None:
                ref = v1
            else:
                assert_normwise(v1, ref, rtol=1e-6, what=f"variant {variant}")
    finally:
        L.set_tunable("accum_variant", 0)
    # sum of class sums == plain sum over pixels (linearity / checksum of checksums)
    onehot = torch.nn.functional.one_hot(out.argmax(1), 19).permute(0, 3, 1, 2).float()
    want = torch.einsum("ndhw,nchw->ncd", feat.double(), onehot.double())
    cnt = onehot.sum((2, 3)).clamp(min=1).unsqueeze(2)
    assert_normwise(ref.double() * cnt, want, what="class sums")


@pytest.mark.parametrize("hi,lo", [((512, 1024), (65, 129)), ((512, 896), (65, 113)), ((64, 96), (8, 12)), ((37, 53), (37, 53))])
def test_update_from_full_resolution_labels(D, hi, lo):
    """labels_full= (the [B,H,W] int64 map, nearest down-sampling folded into the assign kernel) == the reference's
    .float() + F.interpolate(mode='nearest') (self_training.py:327-330) followed by labels_val=, bit for bit."""
    from diga_b200 import synthetic as S
    from diga_b200.calc_centroids import _labels_on_feature_grid
    g = S.gen(15, "cuda")
    n, c, d = 3, 19, 64
    feat, out = S.features((n, d, *lo), g), S.logits((n, c, *lo), g)
    out[:, :5] += 3
    lab = S.block_labels(n, hi[0], hi[1], g, 16)
    cen = S.centroids(c, d, g)
    a, b = D.Class_Features(c, d), D.Class_Features(c, d)
    for cf in (a, b):
        cf.objective_vectors = cen.clone()
        cf.objective_vectors_num = torch.full((c,), 150.0)
    a.update_from_features(feat, out, _labels_on_feature_grid(lab, lo), start_mean=False)
    b.update_from_features(feat, out, start_mean=False, labels_full=lab)
    assert torch.equal(a.objective_vectors, b.objective_vectors) and torch.equal(a.objective_vectors_num, b.objective_vectors_num)
    if lo[0] * lo[1] > 1000:
        assert not torch.equal(a.objective_vectors, cen)                  # some class did reach 5 agreeing pixels
    with pytest.raises(ValueError):
        b.update_from_features(feat, out, _labels_on_feature_grid(lab, lo), labels_full=lab)


# ------------------------------------------------------------------------------------------------ a5 distance
@pytest.mark.parametrize("name", ["proto_d256", "proto_d64"])
def test_proto_golden(D, golden, name):
    g = golden(name)
    cf = D.Class_Features(19, g["feat"].shape[1])
    cf.objective_vectors = T(g["centroids"])
    feat = T(g["feat"], dev())
    assert_normwise(cf.feat_centroid_distance(feat), g["dist"], what="dist")
    assert_normwise(cf.get_centroid_weight(feat), g["weight"], what="weight")
    assert_normwise(cf.get_centroid_distance(feat), g["negdist"], what="-dist")


@pytest.mark.parametrize("n,d,h,w,c", [(2, 2048, 65, 129, 19), (1, 256, 65, 113, 19), (3, 512, 33, 65, 16),
                                       (1, 2048, 16, 128, 19), (1, 100, 7, 5, 19), (2, 256, 64, 128, 19),
                                       (1, 2048, 129, 257, 19), (1, 256, 129, 257, 19),      # BASELINE config 5: one image
                                       (2, 2048, 129, 257, 19)])                             # ... and the pair the writer batches
def test_proto_vs_oracle(D, n, d, h, w, c):
    from diga_b200 import synthetic as S
    g = S.gen(17, "cuda")
    feat = S.features((n, d, h, w), g)
    ocf, gcf = O.ClassFeaturesOracle(c, d), D.Class_Features(c, d)
    cen = S.centroids(c, d, g, feat.mean().item())
    ocf.objective_vectors = cen                     # oracle evaluated on the GPU (torch eager)
    gcf.objective_vectors = cen
    dist_o, w_o = ocf.feat_centroid_distance(feat), ocf.get_centroid_weight(feat)
    dist_g, w_g = gcf.feat_centroid_distance(feat), gcf.get_centroid_weight(feat)
    # fp64 truth: the kernel must be at least as close to it as 1e-5 and the argmin must agree
    d64 = torch.cdist(feat.double().permute(0, 2, 3, 1).reshape(n, h * w, d), cen.double().unsqueeze(0).expand(n, -1, -1))
    d64 = d64.reshape(n, h, w, c).permute(0, 3, 1, 2)
    assert_normwise(dist_g, d64, what="dist vs fp64")
    assert_normwise(dist_g, dist_o, what="dist vs reference chain")
    # weight = softmax(-dist): d w = w * d(dist) (absolute), so a distance held to 1e-5 RELATIVE moves a weight by up
    # to 1e-5 * max(dist) relative.  That conditioning bound is the bar (at D=2048, dist ~ 64, two fp32 evaluations of
    # the reference's own formula already differ by more than 1e-5 * max|w|); on top, the kernel must be no further
    # from the fp64 truth than a small multiple of the reference chain's own fp32 error.
    w64 = torch.softmax(-d64, dim=1)
    err_g = (w_g.double() - w64).abs().max().item()
    err_o = (w_o.double() - w64).abs().max().item()
    tol = 1e-5 * d64.max().item() * w64.max().item()
    assert err_g <= tol, f"weight error {err_g:.3e} exceeds the conditioning bound {tol:.3e}"
    assert err_g <= max(0.1 * tol, 4 * err_o), f"weight error {err_g:.3e} vs reference chain's own {err_o:.3e}"
    mism = dist_g.argmin(1) != d64.argmin(1)
    if mism.any():
        top2 = d64.topk(2, dim=1, largest=False).values
        margin = (top2[:, 1] - top2[:, 0])[mism]
        assert margin.max().item() <= 1e-6 * d64.max().item(), f"argmin flips with margin {margin.max().item():.3e}"
    assert torch.allclose(w_g.sum(1), torch.ones_like(w_g.sum(1)), atol=1e-5)


# ------------------------------------------------------------------------------------------------ a4 selection
def test_consensus_vs_gpu_eager(D):
    """Bit-exact against the reference op chain run on the same GPU (the arbiter for interpolated labels)."""
    from diga_b200 import synthetic as S
    g = S.gen(23, "cuda")
    for (b, h, w, hh, ww) in ((8, 65, 129, 512, 1024), (3, 65, 113, 512, 896), (2, 9, 13, 63, 95), (1, 4, 4, 4, 4)):
        feat = S.features((b, 64, h, w), g)
        cf = D.Class_Features(19, 64)
        cf.objective_vectors = S.centroids(19, 64, g)
        wl = cf.get_centroid_weight(feat)
        pp = S.block_labels(b, hh, ww, g, 8)
        kept, fp = D.consensus_select(pp, wl, (hh, ww))
        kept_o, fp_o = O.consensus_select(pp, wl, (hh, ww))
        nm = int((fp != fp_o).sum())
        assert nm == 0, f"{nm} of {fp.numel()} arg-max labels differ from the GPU eager reference at {(b, h, w, hh, ww)}"
        assert torch.equal(kept, kept_o)
        assert kept.dtype == torch.int64 and fp.dtype == torch.int64
        kept2, none = D.consensus_select(pp, wl, want_feat_pseudo=False)
        assert none is None and torch.equal(kept2, kept)


def test_consensus_select_uint8_equals_int64(D):
    """The uint8 variant (offline pseudo-label path) returns the same labels as the int64 kernel, incl. odd widths."""
    from diga_b200 import synthetic as S
    g = S.gen(23, "cuda")
    for b, c, lo, hi in ((2, 19, (65, 129), (512, 1024)), (1, 16, (9, 13), (37, 53)), (1, 19, (129, 257), (1024, 2048))):
        wl = torch.softmax(S.logits((b, c, *lo), g), 1)
        pl = S.block_labels(b, hi[0], hi[1], g, 16, c)
        k64, f64 = D.consensus_select(pl, wl)
        k8, f8 = D.consensus_select(pl.to(torch.uint8), wl)
        assert k8.dtype == torch.uint8 and f8.dtype == torch.uint8
        assert torch.equal(k8.long(), k64) and torch.equal(f8.long(), f64)
        k8b, none = D.consensus_select(pl.to(torch.uint8), wl, want_feat_pseudo=False)
        assert none is None and torch.equal(k8b, k8)


def test_consensus_golden(D, golden):
    """Golden from torch's CPU interpolation: only near-ties of the up-sampled weights may differ."""
    g = golden("consensus")
    kept, fp = D.consensus_select(T(g["pseudo_prob"], dev()), T(g["weights_lowres"], dev()), g["out_size"].tolist())
    mism = fp.cpu().numpy() != g["feat_pseudo"]
    assert (g["top2_margin"][mism] <= 1e-6).all()
    ok = ~mism
    assert np.array_equal(kept.cpu().numpy()[ok], g["tlabelv_pseudo"][ok])
    assert mism.sum() <= 2


# ------------------------------------------------------------------------------------------------ next rows f2 / f4
@pytest.mark.parametrize("name", ["ce_c19", "ce_weighted_sum"])
def test_cross_entropy2d_golden(D, golden, name):
    g = golden(name)
    x = T(g["input"], dev()).requires_grad_(True)
    wt = T(g["weight"], dev()) if g["weight"].size else None
    loss = D.cross_entropy2d(x, T(g["target"], dev()), weight=wt, size_average=bool(g["size_average"]))
    (loss * float(g["upstream"])).backward()
    assert_rel(loss.item(), g["loss"], what="loss")
    assert_normwise(x.grad, g["grad"], what="grad")
    # ignored / dropped pixels receive exactly zero gradient
    dead = (g["target"] < 0) | (g["target"] == 255)
    assert float(x.grad.permute(0, 2, 3, 1)[T(dead, dev())].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(8, 19, 512, 1024), (2, 19, 33, 65), (1, 16, 7, 9), (2, 5, 4, 6)])
def test_cross_entropy2d_vs_oracle_on_gpu(D, shape):
    from diga_b200 import synthetic as S
    g = S.gen(31, "cuda")
    x = S.logits(shape, g)
    tgt = S.block_labels(shape[0], shape[2], shape[3], g, 8, shape[1], 0.15)
    xo = x.clone().requires_grad_(True)
    lo = O.cross_entropy2d(xo, tgt)
    (lo * 1.5).backward()
    xg = x.clone().requires_grad_(True)
    lg = D.cross_entropy2d(xg, tgt)
    (lg * 1.5).backward()
    assert_rel(lg.item(), lo.item(), what="loss")
    assert_normwise(xg.grad, xo.grad, what="grad")
    assert D.cross_entropy2d(x, tgt).item() == lg.item()           # deterministic reduction


def test_ema_teacher_update_golden_bit_exact(D, golden):
    import torch.nn as nn
    g = golden("ema")

    def build(flat):
        net = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Conv2d(8, 5, 1), nn.Linear(7, 3))
        off = 0
        for p, n in zip(net.parameters(), g["sizes"]):
            p.data.copy_(T(flat[off:off + n]).reshape(p.shape))
            off += int(n)
        return net.to(dev())
    for it, kw in ((0, {}), (7, {}), (5000, {}), (3, {"stage0": False, "mean": True}), (3, {"stage0": False})):
        key = f"after_it{it}_{'_'.join(k for k in kw) or 'stage0'}"
        teacher, student = build(g["teacher"]), build(g["student"])
        out = D.update_teacher_params(teacher, student, it, **kw)
        assert out is teacher
        got = torch.cat([p.detach().reshape(-1) for p in teacher.parameters()]).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), g[key].view(np.uint32)), key


def test_ema_many_tensors_vs_oracle(D):
    """More tensors than fit one launch (512), odd sizes, unaligned tails."""
    g = torch.Generator().manual_seed(3)
    sizes = [1, 3, 4, 5, 63, 64, 65, 1000, 4097, 300000] * 11 + [7, 129, 2048, 4096, 4100] * 90     # 560 tensors
    ts = [torch.randn(n, generator=g) for n in sizes]
    ss = [torch.randn(n, generator=g) for n in sizes]
    alpha = min(1 - 1 / (123 + 1), 0.999)
    want = [alpha * t + (1 - alpha) * s for t, s in zip(ts, ss)]
    from diga_b200.util.utils import ema_update_tensors
    tg, sg = [t.to(dev()) for t in ts], [s.to(dev()) for s in ss]
    ema_update_tensors(tg, sg, alpha)
    for a, b in zip(tg, want):
        assert np.array_equal(a.cpu().numpy().view(np.uint32), b.numpy().view(np.uint32))


# ------------------------------------------------------------------------------------------------ boundary behaviour
def test_no_cpu_fallback(D):
    with pytest.raises(RuntimeError):
        D.pseudo_label(torch.zeros(1, 19, 4, 4))
    with pytest.raises(RuntimeError):
        D.consensus_select(torch.zeros(1, 4, 4, dtype=torch.int64), torch.zeros(1, 19, 2, 2))
    from diga_b200 import _lib as L
    before = L.launch_count()
    D.pseudo_label(torch.zeros(1, 19, 4, 4, device=dev()))
    assert L.launch_count() == before + 1


def test_edge_shapes(D):
    """Empty batches, single pixels, the maximum class count, images with nothing to accumulate."""
    z = torch.zeros(0, 19, 8, 8, device=dev())
    lab, conf = D.pseudo_label(z)
    assert lab.shape == (0, 8, 8) and conf.shape == (0, 8, 8)
    assert D.classmix(torch.zeros(0, 4, 4, dtype=torch.int64, device=dev()), torch.zeros(0, 3, 4, 4, device=dev()),
                      torch.zeros(0, 3, 4, 4, device=dev()))[0].shape == (0, 4, 4)
    # one pixel, two views
    t, s = torch.randn(2, 19, 1, 1, device=dev()), torch.randn(2, 19, 1, 1, device=dev())
    assert_rel(D.distillation_loss(t, s).item(), O.distillation_loss(t, s).item())
    # 32 classes (the ABI maximum) through the padded instantiation
    z32 = 3 * torch.randn(2, 32, 9, 11, device=dev())
    lab32, conf32 = D.pseudo_label(z32)
    assert torch.equal(lab32.long(), z32.argmax(1))
    assert_normwise(conf32, torch.softmax(z32, 1).max(1).values)
    cf = D.Class_Features(32, 256)
    feat = torch.randn(1, 256, 9, 11, device=dev())
    cf.objective_vectors = torch.randn(32, 256)
    ocf = O.ClassFeaturesOracle(32, 256)
    ocf.objective_vectors = cf.objective_vectors.clone()
    assert_normwise(cf.feat_centroid_distance(feat), ocf.feat_centroid_distance(feat))
    # nothing to accumulate: every label disagrees with the prediction -> no vectors, state untouched
    cf19 = D.Class_Features(19, 64)
    before = cf19.objective_vectors.clone()
    f = torch.randn(2, 64, 6, 7, device=dev())
    out = 3 * torch.randn(2, 19, 6, 7, device=dev())
    labels = torch.full((2, 1, 6, 7), 255.0, device=dev())
    vec, ids = cf19.calculate_mean_vector(f, out, labels)
    assert vec == [] and ids == []
    cf19.update_from_features(f, out, labels)
    assert torch.equal(cf19.objective_vectors, before) and float(cf19.objective_vectors_num.sum()) == 0.0
    # empty batch through the fused chain and the exact-mode pass: a no-op
    cf19.update_from_features(f[:0], out[:0], None, "mean")
    assert torch.equal(cf19.objective_vectors, before)
    from diga_b200.parallel import ShardedCentroidPass
    sp = ShardedCentroidPass(cf19, 0, batch=4)
    assert list(sp.my_batches()) == []
    sp.finish()
    assert torch.equal(cf19.objective_vectors, before)


def test_streams_and_noncontiguous(D):
    """Launches follow torch's current stream; non-contiguous inputs are accepted like the reference's ops."""
    s = torch.cuda.Stream()
    z = 3 * torch.randn(1, 19, 32, 64, device=dev())
    with torch.cuda.stream(s):
        lab, _ = D.pseudo_label(z)
    s.synchronize()
    assert torch.equal(lab.long(), z.argmax(1))
    zt = z.permute(0, 1, 3, 2)                       # non-contiguous view
    lab_t, _ = D.pseudo_label(zt)
    assert torch.equal(lab_t.long(), zt.argmax(1))
