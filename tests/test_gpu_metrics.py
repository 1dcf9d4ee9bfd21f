"""GPU parity of the evaluation confusion matrix (csrc/metrics.cu; util/metrics.py:26-76 of the reference): bit-exact
matrix and scores against the reference-made golden and the oracle, for every label dtype the producers deliver."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import diga_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def scores(rs):
    with contextlib.redirect_stdout(io.StringIO()) as out:
        s, iu = rs.get_scores()
    return s, iu, out.getvalue()


def test_running_score_golden(golden):
    from diga_b200.util.metrics import runningScore
    g = golden("running_score")
    rs = runningScore(19)
    rs.update(g["gt"][:2], g["pred"][:2])                                   # numpy in, like the reference call site
    rs.update(torch.from_numpy(g["gt"][2:]).to(DEV), torch.from_numpy(g["pred"][2:]).to(DEV))
    assert np.array_equal(rs.confusion_matrix, g["confusion_matrix"]) and rs.confusion_matrix.dtype == np.float64
    s, iu, printed = scores(rs)
    assert s['Overall Acc: \t'] == g["overall_acc"] and s['Mean Acc : \t'] == g["mean_acc"]
    assert s['FreqW Acc : \t'] == g["fwavacc"] and s['Mean IoU : \t'] == g["mean_iu"]
    assert np.array_equal(np.array([iu[k] for k in range(19)]), g["cls_iu"], equal_nan=True)
    assert printed.count("===>") == 19 and "===>road:" in printed
    rs.reset()
    assert rs.confusion_matrix.sum() == 0


@pytest.mark.parametrize("n_class,shape", [(19, (2, 37, 53)), (16, (1, 64, 64)), (5, (3, 7, 9)), (32, (1, 128, 130))])
@pytest.mark.parametrize("t_dtype,p_dtype", [(torch.int64, torch.int64), (torch.uint8, torch.uint8), (torch.int64, torch.uint8),
                                             (torch.uint8, torch.int64)])
def test_confusion_vs_oracle(n_class, shape, t_dtype, p_dtype):
    from diga_b200.util.metrics import runningScore
    g = torch.Generator().manual_seed(n_class)
    gt = torch.randint(0, n_class, shape, generator=g)
    gt[torch.rand(shape, generator=g) < 0.15] = 255
    pred = torch.randint(0, n_class, shape, generator=g)
    if t_dtype == torch.int64:
        gt[0, 0, :3] = -1                                                   # negative ground truth: masked out (:33)
    rs = runningScore(n_class)
    rs.update(gt.to(t_dtype).to(DEV), pred.to(p_dtype).to(DEV))
    ref = O.RunningScoreOracle(n_class)
    ref.update(gt.numpy(), pred.numpy())
    assert np.array_equal(rs.confusion_matrix, ref.confusion_matrix)


def test_confusion_full_size_properties_and_fused_eval_path():
    """The evaluation loop of train_DiGA_gta2city_self_training.py:428-442 on the GPU: two-scale logits -> fused
    up-sampling + max + argmax (uint8, never materialised) -> confusion matrix against an int64 ground truth, 4 images
    at 1024x2048.  Checks: equals the oracle on the same label maps; matrix total == counted pixels; row sums == class
    histogram of the ground truth; column sums == histogram of the prediction over counted pixels."""
    import diga_b200 as D
    from diga_b200 import synthetic as S
    from diga_b200.util.metrics import runningScore
    g = S.gen(21, DEV)
    n, hh, ww = 4, 1024, 2048
    gt = S.block_labels(n, hh, ww, g)
    pred_u8, _ = D.pseudo_label_two_scale(S.logits((n, 19, 129, 257), g), S.logits((n, 19, 65, 129), g), (hh, ww), want_conf=False)
    rs = runningScore(19)
    rs.update(gt, pred_u8)
    m = rs.confusion_matrix
    ref = O.RunningScoreOracle(19)
    ref.update(gt.cpu().numpy(), pred_u8.cpu().numpy().astype(np.int64))
    assert np.array_equal(m, ref.confusion_matrix)
    counted = gt < 19
    assert m.sum() == int(counted.sum())
    assert np.array_equal(m.sum(1), torch.bincount(gt[counted], minlength=19).cpu().numpy().astype(np.float64))
    assert np.array_equal(m.sum(0), torch.bincount(pred_u8[counted].long(), minlength=19).cpu().numpy().astype(np.float64))
    s, _, _ = scores(rs)
    rscore, _ = ref.get_scores()
    assert s == rscore


def test_confusion_bad_prediction_and_errors():
    from diga_b200.util.metrics import runningScore
    rs = runningScore(19)
    gt = torch.zeros((1, 4, 4), dtype=torch.int64, device=DEV)
    pred = torch.full((1, 4, 4), 19, dtype=torch.int64, device=DEV)         # out of range on counted pixels
    rs.update(gt, pred)
    with pytest.raises(ValueError):
        rs.confusion_matrix
    rs.reset()
    rs.update(torch.full_like(gt, 255), pred)                               # ... but fine where the ground truth is ignored
    assert rs.confusion_matrix.sum() == 0
    with pytest.raises(ValueError):
        rs.update(gt, pred[:, :2])
