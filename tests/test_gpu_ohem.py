"""GPU parity of OhemCrossEntropy from the stride-8 scores (csrc/ohem_up.cu + the CE gradient pass of csrc/loss_up.cu)
against the reference-made goldens and the oracle in fp64.  Bars: loss 1e-5 relative, gradient 1e-5 norm-wise, the kept-
pixel count exact up to probabilities that sit within fp32 rounding of the threshold."""
import numpy as np
import pytest
import torch

from oracle import diga_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def normwise(got, want, what, rtol=1e-5):
    got, want = got.detach().double().cpu(), torch.as_tensor(want).detach().double().cpu()
    err, scale = (got - want).abs().max().item(), want.abs().max().item()
    assert err <= rtol * max(scale, 1e-30), f"{what}: max|diff| {err:.3e} > {rtol} * {scale:.3e}"


@pytest.mark.parametrize("name", ["ohem_low_thresh", "ohem_low_minkept", "ohem_full"])
def test_ohem_golden(golden, name):
    import diga_b200 as D
    g = golden(name)
    sc = T(g["score"]).requires_grad_(True)
    wt = T(g["weight"]) if g["weight"].size else None
    crit = D.OhemCrossEntropy(255, float(g["thres"]), int(g["min_kept"]), wt)
    loss = crit(sc, T(g["target"]))
    (loss * float(g["upstream"])).backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    normwise(sc.grad, g["grad"], "ohem grad")


@pytest.mark.parametrize("n,c,lo,hi,thres,kept", [(2, 19, (17, 33), (128, 256), 0.7, 100000),     # thresh decides
                                                  (2, 19, (17, 33), (128, 256), 0.01, 20000),     # order statistic decides
                                                  (1, 16, (12, 21), (83, 301), 0.5, 5000),        # ragged geometry
                                                  (2, 7, (6, 9), (6, 9), 0.9, 10),                # full-size score, padded C
                                                  (1, 19, (9, 13), (64, 96), 0.7, 10 ** 9)])      # min_kept > M: last element
def test_ohem_vs_oracle(n, c, lo, hi, thres, kept):
    import diga_b200 as D
    g = torch.Generator(device=DEV).manual_seed(31)
    sc = 3.0 * torch.randn((n, c, *lo), generator=g, device=DEV)
    tg = torch.randint(0, c, (n, *hi), generator=g, device=DEV)
    tg[torch.rand((n, *hi), generator=g, device=DEV) < 0.2] = 255
    wt = 0.5 + torch.rand((c,), generator=g, device=DEV)
    for weight in (None, wt):
        s1 = sc.clone().requires_grad_(True)
        loss = D.OhemCrossEntropy(255, thres, kept, weight)(s1, tg)
        (loss * 0.7).backward()
        s2 = sc.double().cpu().requires_grad_(True)
        ref = O.OhemCrossEntropyOracle(255, thres, kept, None if weight is None else weight.double().cpu())(s2, tg.cpu())
        (ref * 0.7).backward()
        assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), (loss.item(), ref.item())
        normwise(s1.grad, s2.grad, "ohem grad vs fp64")
    # deterministic
    s3 = sc.clone().requires_grad_(True)
    l3 = D.OhemCrossEntropy(255, thres, kept, wt)(s3, tg)
    (l3 * 0.7).backward()
    assert torch.equal(l3, loss) and torch.equal(s3.grad, s1.grad)


def test_ohem_order_statistic_is_exact():
    """The radix select returns exactly sorted(pred)[min(min_kept, M - 1)] of the kernel's own probabilities."""
    import diga_b200 as D
    from diga_b200 import _lib as L
    g = torch.Generator(device=DEV).manual_seed(32)
    n, c, lo, hi = 2, 19, (33, 65), (256, 512)
    sc = 3.0 * torch.randn((n, c, *lo), generator=g, device=DEV)
    tg = torch.randint(0, c, (n, *hi), generator=g, device=DEV)
    tg[torch.rand((n, *hi), generator=g, device=DEV) < 0.3] = 255
    pred = torch.empty((n, *hi), device=DEV)
    losspx = torch.empty((n, *hi), device=DEV)
    loss, count, thr = (torch.empty((), device=DEV) for _ in range(3))
    for kept in (0, 1, 777, 50000, 10 ** 8):
        L.check(L.lib.diga_ohem_up_fwd(sc.data_ptr(), tg.data_ptr(), None, n, c, lo[0], lo[1], hi[0], hi[1], 255, 0.0, kept,
                                       pred.data_ptr(), losspx.data_ptr(), loss.data_ptr(), count.data_ptr(), thr.data_ptr(),
                                       L.ohem_workspace(sc.device).data_ptr(), L.stream()))
        valid = pred[pred >= 0].sort().values
        want = valid[min(kept, valid.numel() - 1)]
        assert thr.item() == want.item(), (kept, thr.item(), want.item())
        assert count.item() == float((valid < want).sum().item())
        assert (pred < 0).sum().item() == (tg == 255).sum().item()
