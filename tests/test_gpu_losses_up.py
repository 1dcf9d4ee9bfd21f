"""GPU parity of the fused up-sampling losses (csrc/loss_up.cu): KD and cross_entropy2d evaluated
from the stride-8 logits, against the golden fixture made by the reference's own statements and against the oracle
(`loss(upsample(low))` with autograd through F.interpolate) on seeded inputs.  Bar: 1e-5 relative (norm-wise for
gradients); the backward is additionally required to be bitwise deterministic."""
import numpy as np
import pytest
import torch

from oracle import diga_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5


def T(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def normwise(got, want, what, rtol=RTOL):
    got, want = got.detach().double().cpu(), torch.as_tensor(want).detach().double().cpu()
    assert got.shape == want.shape, what
    err, scale = (got - want).abs().max().item(), want.abs().max().item()
    assert err <= rtol * max(scale, 1e-30), f"{what}: max|diff| {err:.3e} > {rtol} * {scale:.3e}"


def rel(got, want, what, rtol=RTOL):
    got, want = float(got), float(want)
    assert abs(got - want) <= rtol * abs(want), f"{what}: {got!r} vs {want!r}"


@pytest.fixture(scope="module")
def D():
    import diga_b200
    return diga_b200


def test_losses_up_golden(D, golden):
    g = golden("losses_up")
    stu = T(g["student_low"]).requires_grad_(True)
    mix = T(g["mix_low"]).requires_grad_(True)
    tea, sl, ml = T(g["teacher_low"]), T(g["slabel"]), T(g["mixlabel"])
    loss_src, loss_kd = D.seg_distillation_losses_upsampled(tea, stu, sl, float(g["kd_scale"]))
    loss_seg = loss_src + D.cross_entropy2d_upsampled(mix, ml)
    total = float(g["lambda_seg"]) * loss_seg + float(g["lambda_distil"]) * loss_kd
    total.backward()
    rel(loss_src, g["loss_semseg_src"], "loss_semseg (source)")
    rel(loss_seg, g["loss_semseg"], "loss_semseg")
    rel(loss_kd, g["loss_distil"], "loss_s_distil")
    rel(total, g["total_loss"], "total_loss")
    normwise(stu.grad, g["grad_student_low"], "d total / d s_pred_cat_stu (stride 8)")
    normwise(mix.grad, g["grad_mix_low"], "d total / d cross_pred_mix (stride 8)")


GEOMS = [  # n2, C, (h, w), (H, W)
    (4, 19, (9, 13), (64, 96)),          # W < one tile
    (2, 19, (17, 33), (128, 256)),       # two full tiles, exact 8x cells
    (2, 19, (12, 21), (83, 301)),        # ragged strips / tiles, non-integer scale
    (2, 16, (7, 40), (50, 300)),         # Synthia class count
    (2, 7, (5, 6), (21, 37)),            # padded generic class count
    (2, 19, (6, 9), (6, 9)),             # identity geometry (scale 1)
    (2, 19, (1, 1), (5, 130)),           # single source pixel
    (2, 19, (3, 5), (40, 41)),           # 13x up-sampling: more than two strips per source row
]


def inputs(n2, c, lo, seed, sigma=3.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    tea = sigma * torch.randn((n2, c, *lo), generator=g, device=DEV)
    stu = sigma * torch.randn((n2, c, *lo), generator=g, device=DEV)
    return tea, stu, g


def labels(n, hi, c, g):
    t = torch.randint(0, c, (n, *hi), generator=g, device=DEV)
    r = torch.rand((n, *hi), generator=g, device=DEV)
    t[r < 0.2] = 255
    t[r > 0.97] = -1
    return t


@pytest.mark.parametrize("n2,c,lo,hi", GEOMS)
def test_kd_upsampled_vs_oracle(D, n2, c, lo, hi):
    tea, stu, _ = inputs(n2, c, lo, 11)
    for scale, up in ((0.5, 0.25), (0.25, 1.0)):
        s1 = stu.clone().requires_grad_(True)
        loss = D.distillation_loss_upsampled(tea, s1, hi, scale)
        (loss * up).backward()
        s2 = stu.clone().requires_grad_(True)
        ref = O.distillation_loss_upsampled(tea, s2, hi, scale)                # eager chain on the same GPU
        (ref * up).backward()
        rel(loss, ref, "kd_up loss")
        normwise(s1.grad, s2.grad, "kd_up grad")
        # fp64 evaluation of the same chain on the CPU: the fused result must be as close to it as the eager chain is
        s3 = stu.double().cpu().requires_grad_(True)
        ref64 = O.distillation_loss_upsampled(tea.double().cpu(), s3, hi, scale)
        (ref64 * up).backward()
        rel(loss, ref64, "kd_up loss vs fp64")
        normwise(s1.grad, s3.grad, "kd_up grad vs fp64")
        # single-pass variant
        l2, g2 = D.distillation_loss_upsampled_and_grad(tea, stu, hi, scale, up)
        assert torch.equal(l2, loss.detach())
        normwise(g2, s1.grad, "single pass vs autograd (unit gradient x upstream)", 1e-6)


@pytest.mark.parametrize("n2,c,lo,hi", GEOMS)
def test_ce_upsampled_vs_oracle(D, n2, c, lo, hi):
    _, x, g = inputs(n2, c, lo, 12)
    tgt = labels(n2, hi, c, g)
    wt = 0.5 + torch.rand((c,), generator=g, device=DEV)
    for weight, avg in ((None, True), (wt, False), (wt, True)):
        x1 = x.clone().requires_grad_(True)
        loss = D.cross_entropy2d_upsampled(x1, tgt, weight, avg)
        (loss * 0.7).backward()
        x2 = x.double().cpu().requires_grad_(True)
        ref = O.cross_entropy2d_upsampled(x2, tgt.cpu(), None if weight is None else weight.double().cpu(), avg)
        (ref * 0.7).backward()
        rel(loss, ref, "ce_up loss")
        normwise(x1.grad, x2.grad, "ce_up grad")


@pytest.mark.parametrize("n2,c,lo,hi", GEOMS[:5])
def test_seg_plus_kd_upsampled_is_the_sum_of_its_parts(D, n2, c, lo, hi):
    tea, stu, g = inputs(n2, c, lo, 13)
    tgt = labels(n2 // 2, hi, c, g)
    s1 = stu.clone().requires_grad_(True)
    l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s1, tgt, 0.5)
    (1.0 * l_ce + 0.25 * l_kd).backward()
    s2 = stu.double().cpu().requires_grad_(True)
    r_ce, r_kd = O.seg_distillation_losses_upsampled(tea.double().cpu(), s2, tgt.cpu(), 0.5)
    (1.0 * r_ce + 0.25 * r_kd).backward()
    rel(l_ce, r_ce, "fused ce")
    rel(l_kd, r_kd, "fused kd")
    normwise(s1.grad, s2.grad, "fused grad")
    # and bit-equal losses to the two stand-alone calls (same per-pixel expressions, same reduction order)
    assert torch.equal(l_kd.detach(), D.distillation_loss_upsampled(tea, stu, hi, 0.5))
    assert torch.equal(l_ce.detach(), D.cross_entropy2d_upsampled(stu[:n2 // 2].contiguous(), tgt))
    # only one of the two outputs used: the other upstream is None
    s3 = stu.clone().requires_grad_(True)
    D.seg_distillation_losses_upsampled(tea, s3, tgt, 0.5)[1].backward()
    s4 = stu.clone().requires_grad_(True)
    D.distillation_loss_upsampled(tea, s4, hi, 0.5).backward()
    normwise(s3.grad, s4.grad, "kd-only upstream")


def test_losses_up_config2_size_matches_materialised_path_and_is_deterministic(D):
    """[8,19,65,129] -> 512x1024 (config 2's logits before nn.Upsample): the fused path against diga's own materialised
    path (bit-identical interpolation + kd kernel) and torch's up-sampling backward; two runs are bit-equal."""
    n2, c, lo, hi = 8, 19, (65, 129), (512, 1024)
    tea, stu, g = inputs(n2, c, lo, 14)
    tgt = labels(n2 // 2, hi, c, g)
    s1 = stu.clone().requires_grad_(True)
    l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s1, tgt, 0.5)
    (l_ce + 0.25 * l_kd).backward()
    s2 = stu.clone().requires_grad_(True)
    up_s = O.upsample_bilinear_ac(s2, hi)
    m_kd = D.distillation_loss(O.upsample_bilinear_ac(tea, hi), up_s, 0.5)
    m_ce = D.cross_entropy2d(up_s[:n2 // 2], tgt)
    (m_ce + 0.25 * m_kd).backward()
    rel(l_kd, m_kd, "kd"), rel(l_ce, m_ce, "ce")
    normwise(s1.grad, s2.grad, "grad vs materialised path")
    s3 = stu.clone().requires_grad_(True)
    a_ce, a_kd = D.seg_distillation_losses_upsampled(tea, s3, tgt, 0.5)
    (a_ce + 0.25 * a_kd).backward()
    assert torch.equal(a_ce, l_ce) and torch.equal(a_kd, l_kd) and torch.equal(s3.grad, s1.grad)
    # linearity of the backward in the upstream scalars (size-independent property)
    s4 = stu.clone().requires_grad_(True)
    b_ce, b_kd = D.seg_distillation_losses_upsampled(tea, s4, tgt, 0.5)
    (2.0 * b_ce + 0.5 * b_kd).backward()
    normwise(s4.grad, 2.0 * s1.grad, "linearity")


def test_losses_up_errors(D):
    tea, stu, g = inputs(2, 19, (8, 8), 15)
    with pytest.raises(ValueError):
        D.distillation_loss_upsampled(tea, stu, (4, 4))                        # down-sampling
    with pytest.raises(ValueError):
        D.distillation_loss_upsampled(tea[:1], stu[:1], (16, 16))              # odd batch
    with pytest.raises(RuntimeError):
        D.distillation_loss_upsampled(tea.cpu(), stu.cpu(), (16, 16))          # no CPU fallback
    with pytest.raises(ValueError):
        D.cross_entropy2d_upsampled(stu, torch.zeros((3, 16, 16), dtype=torch.long, device=DEV))
    with pytest.raises(RuntimeError):
        big = torch.zeros((2, 40, 4, 4), device=DEV)
        D.distillation_loss_upsampled(big, big, (8, 8))                        # C > 32


@pytest.mark.parametrize("n2,c,lo,hi", [GEOMS[0], GEOMS[2], GEOMS[4]])
def test_two_pass_backward_entry_points(D, n2, c, lo, hi):
    """diga_loss_up_bwd for a single loss (the autograd wrapper takes the loss+gradient pass there, so the stand-alone
    backward of the C ABI is exercised directly): same gradient as the wrapper's."""
    from diga_b200 import _lib as L
    tea, stu, g = inputs(n2, c, lo, 16)
    tgt = labels(n2, hi, c, g)
    n, _, h, w = stu.shape
    ws = L.loss_up_workspace(n, c, h, w, hi[0], hi[1], stu.device)
    up = torch.tensor(0.3, device=DEV)
    # KD only
    s1 = stu.clone().requires_grad_(True)
    (D.distillation_loss_upsampled(tea, s1, hi, 0.5) * 0.3).backward()
    ds = torch.empty_like(stu)
    L.check(L.lib.diga_loss_up_bwd(tea.data_ptr(), stu.data_ptr(), None, None, n, 0, c, h, w, hi[0], hi[1], 0.5, 1,
                                   up.data_ptr(), None, None, ds.data_ptr(), ws.data_ptr(), L.stream()))
    normwise(ds, s1.grad, "kd two-pass bwd", 1e-6)
    # CE only (needs the forward's denominator)
    s2 = stu.clone().requires_grad_(True)
    (D.cross_entropy2d_upsampled(s2, tgt) * 0.3).backward()
    loss, kd_dummy, denom = (torch.empty((), device=DEV) for _ in range(3))
    L.check(L.lib.diga_loss_up_fwd(None, stu.data_ptr(), tgt.data_ptr(), None, n, n, c, h, w, hi[0], hi[1], 0.0, 1,
                                   None, loss.data_ptr(), denom.data_ptr(), ws.data_ptr(), L.stream()))
    L.check(L.lib.diga_loss_up_bwd(None, stu.data_ptr(), tgt.data_ptr(), None, n, n, c, h, w, hi[0], hi[1], 0.0, 1,
                                   None, up.data_ptr(), denom.data_ptr(), ds.data_ptr(), ws.data_ptr(), L.stream()))
    normwise(ds, s2.grad, "ce two-pass bwd", 1e-6)


def test_no_grad_forward_takes_the_loss_only_pass(D):
    from diga_b200 import _lib as L
    tea, stu, _ = inputs(2, 19, (9, 13), 17)
    before = L.launch_count()
    with torch.no_grad():
        a = D.distillation_loss_upsampled(tea, stu, (64, 96))
    assert L.launch_count() - before == 1                      # loss kernel only
    s = stu.clone().requires_grad_(True)
    before = L.launch_count()
    b = D.distillation_loss_upsampled(tea, s, (64, 96))
    assert L.launch_count() - before == 1                      # loss+gradient kernel; the patches wait for the upstream scalar
    assert torch.equal(a, b.detach())
    before = L.launch_count()
    (b * 0.37).backward()
    assert L.launch_count() - before == 1                      # the patch gather, scaled by the upstream scalar on the device
    s2 = stu.clone().requires_grad_(True)
    loss, grad = D.distillation_loss_upsampled_and_grad(tea, s2, (64, 96), 0.5, 0.37)   # upstream known: gathered at once
    normwise(s.grad, grad, "deferred gather x upstream vs upstream folded into the pass")


@pytest.mark.parametrize("n2,c,lo,hi", [GEOMS[0], GEOMS[2], GEOMS[3], GEOMS[7]])
def test_weighted_total_single_pass(D, n2, c, lo, hi):
    """seg_distillation_total_upsampled (loss weights known up front, one pass) == the weighted sum of the two-output form,
    with weights / sum reduction / dropped pixels, and the reference chain in fp64."""
    tea, stu, g = inputs(n2, c, lo, 18)
    tgt = labels(n2 // 2, hi, c, g)
    wt = 0.5 + torch.rand((c,), generator=g, device=DEV)
    for weight, avg, lam_s, lam_d in ((None, True, 1.0, 0.25), (wt, True, 0.7, 1.3), (wt, False, 1e-3, 0.5)):
        s1 = stu.clone().requires_grad_(True)
        total, l_ce, l_kd = D.seg_distillation_total_upsampled(tea, s1, tgt, lam_s, lam_d, 0.5, weight, avg)
        assert not l_ce.requires_grad and not l_kd.requires_grad
        (total * 0.5).backward()
        s2 = stu.double().cpu().requires_grad_(True)
        r_ce, r_kd = O.seg_distillation_losses_upsampled(tea.double().cpu(), s2, tgt.cpu(), 0.5,
                                                         None if weight is None else weight.double().cpu(), avg)
        ((lam_s * r_ce + lam_d * r_kd) * 0.5).backward()
        rel(l_ce, r_ce, "ce"), rel(l_kd, r_kd, "kd"), rel(total, lam_s * r_ce + lam_d * r_kd, "total")
        normwise(s1.grad, s2.grad, "weighted single-pass grad")
        with torch.no_grad():
            t2, _, _ = D.seg_distillation_total_upsampled(tea, stu, tgt, lam_s, lam_d, 0.5, weight, avg)
        assert torch.equal(t2, total.detach())


@pytest.mark.parametrize("sigma", [8.0, 20.0, 60.0, 300.0])
@pytest.mark.parametrize("n2,c,lo,hi", [GEOMS[1], GEOMS[2], GEOMS[4]])
def test_wide_logit_ranges_take_the_exact_paths(D, n2, c, lo, hi, sigma):
    """csrc/loss_up.cu advances the exponentials by one multiply per row only while |dif| <= 24 inside a warp's source cell;
    wider cells are evaluated with ex2 per row (sigma 8: a mix of both; 20: nearly every cell), and rows more than 41 logits
    apart re-base on the per-pixel max (sigma 60, 300).  Every regime against the fp64 chain of the reference statements."""
    tea, stu, g = inputs(n2, c, lo, 23, sigma)
    tgt = labels(n2 // 2, hi, c, g)
    s1 = stu.clone().requires_grad_(True)
    l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s1, tgt, 0.5)
    (1.0 * l_ce + 0.25 * l_kd).backward()
    s2 = stu.double().cpu().requires_grad_(True)
    r_ce, r_kd = O.seg_distillation_losses_upsampled(tea.double().cpu(), s2, tgt.cpu(), 0.5)
    (1.0 * r_ce + 0.25 * r_kd).backward()
    assert torch.isfinite(l_ce) and torch.isfinite(l_kd) and torch.isfinite(s1.grad).all()
    rel(l_ce, r_ce, f"ce, sigma {sigma}")
    rel(l_kd, r_kd, f"kd, sigma {sigma}")
    normwise(s1.grad, s2.grad, f"grad, sigma {sigma}")


def test_one_confident_class_next_to_flat_rows(D):
    """A cell whose top row is flat and whose bottom row has one class 100 logits up: the re-based exact pass inside an
    otherwise ordinary map (the warp votes, so its neighbours go through the same pass)."""
    n2, c, lo, hi = 2, 19, (9, 13), (64, 96)
    tea, stu, g = inputs(n2, c, lo, 29, 1.0)
    stu[:, 3, 4, :] += 100.0
    tea[:, 7, 5, 2:9] -= 150.0
    tgt = labels(1, hi, c, g)
    s1 = stu.clone().requires_grad_(True)
    l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s1, tgt, 0.25)
    (l_ce + l_kd).backward()
    s2 = stu.double().cpu().requires_grad_(True)
    r_ce, r_kd = O.seg_distillation_losses_upsampled(tea.double().cpu(), s2, tgt.cpu(), 0.25)
    (r_ce + r_kd).backward()
    rel(l_ce, r_ce, "ce")
    rel(l_kd, r_kd, "kd")
    normwise(s1.grad, s2.grad, "grad")


@pytest.mark.parametrize("n2,c,lo,hi", [GEOMS[0], GEOMS[2], GEOMS[4], GEOMS[5]])
def test_source_rows_from_global_memory_equal_the_shared_memory_tile(D, n2, c, lo, hi):
    """Geometries whose source tile would not fit in shared memory read the rows from global memory (forced here through
    the `lossup_tile` tunable): same expressions, bit-equal losses and gradients."""
    from diga_b200 import _lib as L
    tea, stu, g = inputs(n2, c, lo, 31)
    tgt = labels(n2 // 2, hi, c, g)

    def run():
        s = stu.clone().requires_grad_(True)
        l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s, tgt, 0.5)
        (l_ce + 0.3 * l_kd).backward()
        return l_ce.detach(), l_kd.detach(), s.grad

    a = run()
    L.set_tunable("lossup_tile", 0)
    try:
        b = run()
    finally:
        L.set_tunable("lossup_tile", 1)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_single_loss_outputs_and_unused_upstreams(D):
    """A switched-off loss is None inside the autograd function (no fill launch), an upstream that never arrives leaves the
    other loss's gradient untouched, and the result equals the two-loss call with an explicit zero weight."""
    tea, stu, g = inputs(2, 19, (12, 21), 41)
    tgt = labels(1, (83, 301), 19, g)
    s1 = stu.clone().requires_grad_(True)
    l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s1, tgt, 0.5)
    (0.7 * l_kd).backward()                                   # the CE upstream is None in backward
    s2 = stu.clone().requires_grad_(True)
    a_ce, a_kd = D.seg_distillation_losses_upsampled(tea, s2, tgt, 0.5)
    (0.0 * a_ce + 0.7 * a_kd).backward()                      # the CE upstream is a device zero
    assert torch.equal(s1.grad, s2.grad)
    s3 = stu.clone().requires_grad_(True)
    D.distillation_loss_upsampled(tea, s3, (83, 301), 0.5).mul(0.7).backward()
    normwise(s3.grad, s1.grad, "KD-only single pass vs KD part of the two-loss backward")
    s4 = stu[:1].clone().requires_grad_(True)
    l = D.cross_entropy2d_upsampled(s4, tgt)
    (g4,) = torch.autograd.grad(l * 3.0, s4)
    s5 = stu.clone().requires_grad_(True)
    b_ce, _ = D.seg_distillation_losses_upsampled(tea, s5, tgt, 0.5)
    (3.0 * b_ce).backward()
    rel(l, b_ce, "CE-only loss vs the CE part of the two-loss call")
    normwise(g4, s5.grad[:1], "CE-only single pass vs CE part of the two-loss backward")
    assert float(s5.grad[1:].abs().max()) == 0.0


@pytest.mark.parametrize("n2,c,lo,hi", [GEOMS[0], GEOMS[2], GEOMS[3]])
def test_promised_denominator_skips_the_count_and_is_checked(D, n2, c, lo, hi):
    """targets_nonnegative=True (loader labels: trainIds or 255) hands the size_average denominator to the kernel instead of
    counting it first: bit-equal losses and gradient when the promise holds, NaN losses when a negative target breaks it."""
    from diga_b200 import _lib as L
    tea, stu, g = inputs(n2, c, lo, 57)
    tgt = labels(n2 // 2, hi, c, g)
    clean = torch.where(tgt < 0, torch.full_like(tgt, 255), tgt)

    def run(target, promise):
        s = stu.clone().requires_grad_(True)
        n0 = L.launch_count()
        total, l_ce, l_kd = D.seg_distillation_total_upsampled(tea, s, target, 0.7, 1.3, 0.5, targets_nonnegative=promise)
        launches = L.launch_count() - n0
        total.backward()
        return total.detach(), l_ce, l_kd, s.grad, launches

    a, b = run(clean, False), run(clean, True)
    assert all(torch.equal(x, y) for x, y in zip(a[:4], b[:4]))
    assert a[4] == b[4] + 1                                    # the counting launch is gone
    assert (tgt < 0).any()
    bad = run(tgt, True)
    assert torch.isnan(bad[0]) and torch.isnan(bad[1]) and torch.isfinite(bad[2])
    good = run(tgt, False)
    assert torch.isfinite(good[0]) and torch.isfinite(good[1])


@pytest.mark.parametrize("n2,c,lo,hi", GEOMS)
def test_deferred_gather_reads_only_what_the_pass_wrote(D, n2, c, lo, hi, monkeypatch):
    """The loss+gradient pass leaves its patches in an UNINITIALISED scratch buffer and the backward gathers them: with the
    buffer poisoned with NaN beforehand every gradient must come out finite and bit-equal to the run on a zeroed buffer."""
    from diga_b200.util import loss as LM
    tea, stu, g = inputs(n2, c, lo, 77)
    tgt = labels(n2 // 2, hi, c, g)
    real = LM._new_scratch

    def run(fill):
        def poisoned(*a):
            buf = real(*a)
            buf.fill_(fill)
            return buf
        monkeypatch.setattr(LM, "_new_scratch", poisoned)
        out = []
        s = stu.clone().requires_grad_(True)
        D.seg_distillation_total_upsampled(tea, s, tgt, 0.7, 1.3, 0.5)[0].backward()
        out.append(s.grad)
        s = stu.clone().requires_grad_(True)
        (D.distillation_loss_upsampled(tea, s, hi, 0.5) * 0.3).backward()
        out.append(s.grad)
        s = stu[: n2 // 2].clone().requires_grad_(True)
        (D.cross_entropy2d_upsampled(s, tgt) * 1.7).backward()
        out.append(s.grad)
        return out

    zeroed, nan = run(0.0), run(float("nan"))
    for a, b in zip(zeroed, nan):
        assert torch.isfinite(b).all()
        assert torch.equal(a, b)
