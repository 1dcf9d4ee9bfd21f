"""``diga_b200.nn.Upsample``: the lazy stand-in for the scripts' ``nn.Upsample(bilinear, align_corners=True)`` modules
(train_DiGA_gta2city_self_training.py:190-192, pseudolabel_generator.py:55).  The reference's own statement sequence — up-sample,
then loss — run with the stand-in + this package's functions against torch's ``nn.Upsample`` + the oracle."""
import numpy as np
import pytest
import torch

from oracle import diga_oracle as O

DEV = "cuda:0"


def normwise(a, b, rtol=1e-5):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() <= rtol * max(b.abs().max().item(), 1e-30)


@pytest.mark.gpu
def test_unchanged_call_sites_with_lazy_upsample_match_the_reference_sequence():
    """self_training.py:289,:344,:348-356,:382 verbatim (only the Upsample class and the function modules swapped): losses
    1e-5, gradients with respect to the stride-8 logits 1e-5; the lazy objects are never materialised on this path."""
    import diga_b200 as D
    from diga_b200 import synthetic as S
    g = S.gen(5, DEV)
    B, C, h, w, H, W = 2, 19, 17, 33, 128, 256
    tea_low, stu_low = S.logits((2 * B, C, h, w), g), S.logits((2 * B, C, h, w), g)
    cross_low = S.logits((B, C, h, w), g)
    slabelv, mixlabel = S.block_labels(B, H, W, g, 16), S.block_labels(B, H, W, g, 16)

    def step(upsample_src, seg_loss, distillation_loss, stu, cpm):
        t_pred_cat_tea = upsample_src(tea_low)                                   # :289
        cross_pred_mix = upsample_src(cpm)                                       # :344
        s_pred_stu = upsample_src(stu[:B])                                       # :348
        loss_semseg = seg_loss(s_pred_stu, slabelv)                              # :349
        s_pred_cat_stu = upsample_src(stu)                                       # :351
        loss_distil = distillation_loss(t_pred_cat_tea, s_pred_cat_stu)          # :352
        loss_mix = seg_loss(cross_pred_mix, mixlabel)                            # :355
        total = (loss_semseg + loss_mix) + 0.25 * loss_distil                    # :356,:382
        return total, (t_pred_cat_tea, cross_pred_mix, s_pred_stu, s_pred_cat_stu)

    so, co = stu_low.clone().requires_grad_(True), cross_low.clone().requires_grad_(True)
    ref_total, _ = step(torch.nn.Upsample(size=(H, W), mode="bilinear", align_corners=True), O.cross_entropy2d, O.distillation_loss, so, co)
    ref_total.backward()
    sg, cg = stu_low.clone().requires_grad_(True), cross_low.clone().requires_grad_(True)
    total, lazies = step(D.nn.Upsample(size=(H, W), mode="bilinear", align_corners=True), D.cross_entropy2d, D.distillation_loss, sg, cg)
    total.backward()
    assert abs(total.item() - ref_total.item()) <= 1e-5 * abs(ref_total.item())
    assert normwise(sg.grad, so.grad) and normwise(cg.grad, co.grad)
    assert all(isinstance(z, D.nn.LazyUpsampled) and z._full is None for z in lazies), "a fused consumer materialised its input"
    # Synthia tree: OhemCrossEntropy as seg_loss
    ohem_o, ohem_g = O.OhemCrossEntropyOracle(255, 0.7, 1000), D.OhemCrossEntropy(255, 0.7, 1000)
    s1 = stu_low[:B].clone().requires_grad_(True)
    lo = ohem_o(torch.nn.Upsample(size=(H, W), mode="bilinear", align_corners=True)(s1), slabelv)
    lo.backward()
    s2 = stu_low[:B].clone().requires_grad_(True)
    lg = ohem_g(D.nn.Upsample(size=(H, W), mode="bilinear", align_corners=True)(s2), slabelv)
    lg.backward()
    assert abs(lg.item() - lo.item()) <= 1e-5 * abs(lo.item()) and normwise(s2.grad, s1.grad)


@pytest.mark.gpu
def test_lazy_upsample_materialises_like_nn_upsample():
    """Anything but the fused consumers sees exactly the tensor torch's nn.Upsample returns (values, autograd, methods)."""
    import diga_b200 as D
    from diga_b200 import synthetic as S
    g = S.gen(6, DEV)
    x = S.logits((3, 19, 9, 13), g)
    ref_up, lazy_up = torch.nn.Upsample(size=(40, 60), mode="bilinear", align_corners=True), D.nn.Upsample(size=(40, 60), mode="bilinear", align_corners=True)
    want, z = ref_up(x), lazy_up(x)
    assert isinstance(z, D.nn.LazyUpsampled) and z.shape == want.shape and z.size(2) == 40 and z.dim() == 4 and len(z) == 3
    assert torch.equal(z.materialize(), want)
    assert torch.equal(torch.max(z, z * 0.5), torch.max(want, want * 0.5))                    # pseudolabel_generator.py:80
    assert torch.equal(z.max(1, keepdim=True)[1].squeeze(1), want.max(1, keepdim=True)[1].squeeze(1))   # self_training.py:303
    assert torch.equal(torch.softmax(z, dim=1), torch.softmax(want, dim=1))
    assert torch.equal((z + 1.0), want + 1.0) and torch.equal(z[1:].materialize(), want[1:]) and torch.equal(z[0], want[0])
    assert torch.equal(torch.cat([z, z]), torch.cat([want, want])) and torch.equal(z.cpu(), want.cpu())
    a, b = z.chunk(3)[0], want.chunk(3)[0]
    assert isinstance(a, D.nn.LazyUpsampled) and torch.equal(a.materialize(), b)
    # autograd through the materialised tensor
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    (ref_up(x1) ** 2).sum().backward()
    (lazy_up(x2) ** 2).sum().backward()
    assert normwise(x2.grad, x1.grad, 1e-6)        # (ATen's up-sampling backward accumulates with atomics: not bit-reproducible)
    # one lazy, one real tensor: the loss materialises and still matches
    t = S.logits((4, 19, 40, 60), g)
    s_low = S.logits((4, 19, 9, 13), g)
    lo = O.distillation_loss(t, ref_up(s_low))
    lg = D.distillation_loss(t, lazy_up(s_low))
    assert abs(lg.item() - lo.item()) <= 1e-5 * abs(lo.item())
    # configurations the fused kernels do not cover fall through to F.interpolate
    for kw in (dict(scale_factor=2, mode="nearest"), dict(size=(40, 60), mode="bilinear", align_corners=False), dict(size=(5, 7), mode="bilinear", align_corners=True)):
        assert torch.equal(D.nn.Upsample(**kw)(x), torch.nn.Upsample(**kw)(x))
    assert torch.is_tensor(lazy_up(x.cpu())) and torch.equal(lazy_up(x.cpu()), ref_up(x.cpu()))


@pytest.mark.gpu
def test_pseudo_label_on_lazy_upsampled_logits():
    """pseudolabel_generator.py:77-85 with upsample_1024 swapped for the stand-in: same labels as the explicit fused call and as
    the reference op chain on the GPU (modulo exact softmax ties)."""
    import diga_b200 as D
    from diga_b200 import synthetic as S
    g = S.gen(8, DEV)
    out, out_ds = S.logits((1, 19, 33, 65), g), S.logits((1, 19, 17, 33), g)
    up = D.nn.Upsample(size=(256, 512), mode="bilinear", align_corners=True)
    lab, conf = D.pseudo_label(up(out), up(out_ds))
    lab2, conf2 = D.pseudo_label_two_scale(out, out_ds, (256, 512))
    assert torch.equal(lab, lab2) and torch.equal(conf, conf2)
    lab_o, _ = O.pseudo_label_two_scale(out, out_ds, (256, 512))
    diff = lab[0].cpu().numpy().astype(np.int64) != lab_o
    if diff.any():
        fused = torch.max(O.upsample_bilinear_ac(out, (256, 512)), O.upsample_bilinear_ac(out_ds, (256, 512)))
        prob = torch.softmax(fused, 1)[0].cpu().numpy()
        for y, x in zip(*np.nonzero(diff)):
            assert prob[lab[0, y, x].item(), y, x] == prob[lab_o[y, x], y, x]


def test_lazy_upsampled_generic_behaviour_cpu():
    """Host logic of LazyUpsampled (no CUDA): shape queries and batch slicing stay lazy, everything else materialises through
    F.interpolate with autograd intact."""
    from diga_b200.nn import LazyUpsampled, Upsample
    low = torch.randn(4, 3, 2, 5, requires_grad=True)
    z = LazyUpsampled(low, (6, 9))
    want = torch.nn.functional.interpolate(low, size=(6, 9), mode="bilinear", align_corners=True)
    assert z.shape == want.shape and z.dtype == torch.float32 and not z.is_cuda and z.requires_grad and z._full is None
    assert isinstance(z[:2], LazyUpsampled) and isinstance(z.chunk(2)[1], LazyUpsampled) and isinstance(z.detach(), LazyUpsampled)
    assert z._full is None
    assert torch.equal(torch.relu(z), torch.relu(want)) and torch.equal(z * 2, want * 2) and torch.equal(2 * z, 2 * want)
    assert torch.equal(z.permute(0, 2, 3, 1), want.permute(0, 2, 3, 1)) and torch.equal(z[:, 1], want[:, 1])
    z.sum().backward()
    g1 = low.grad.clone()
    low.grad = None
    want.sum().backward()
    assert torch.equal(g1, low.grad)
    m = Upsample(size=(6, 9), mode="bilinear", align_corners=True)
    assert torch.is_tensor(m(low.detach()))                      # CPU input: plain F.interpolate
