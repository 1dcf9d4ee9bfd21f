"""CPU oracle for DiGA's per-pixel adaptation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diga_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do.  It is the checker, never the thing shipped or measured as product.

What it is
----------
The reference (fy-vision/DiGA) is pure Python whose arithmetic lives in stock ``torch`` /
``numpy`` calls.  This file restates, op for op, the *function-level* hot path
(``util/loss.py``, ``util/utils.py``, ``calc_centroids.py``) and the *inline script blocks*
(``pseudolabel_generator.py``, ``train_DiGA_gta2city_self_training.py``) as plain functions
over tensors.  Every function cites the reference file:line it follows; paths are relative to
``/root/reference/domain_adaptation/GTA5`` (``G/``) unless stated.

All functions are device agnostic: on the CPU they are the CPU oracle (and the CPU baseline
that ``bench.py`` times), on a CUDA tensor the same op chain is the "GPU eager reference"
that SURVEY.md §7 requires for the label paths that run through bilinear interpolation.

Pinning
-------
The reference has no tests, golden vectors or fixtures of its own.  The
oracle is pinned against *outputs of the reference itself*: ``tests/golden/make_golden.py``
imports the real reference functions from ``/root/reference`` in the build container
(``oracle/ref_loader.py``) and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this file against those fixtures bit for bit, and, when ``/root/reference`` is present,
against the live reference functions as well.
"""
from __future__ import annotations

import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

IGNORE = 255


# --------------------------------------------------------------------------------------
# a1  symmetric knowledge distillation   (G/util/loss.py:125-143, S/util/loss.py:52-70)
# --------------------------------------------------------------------------------------
def distillation_loss(teacher_out: torch.Tensor, student_out: torch.Tensor, scale: float = 0.5) -> torch.Tensor:
    """Cross-view soft-target cross entropy, G/util/loss.py:125-143.

    The batch holds two views stacked on dim 0.  Teacher view ``iq`` supervises the *other*
    student view; the pair whose teacher is view 1 is multiplied by ``scale``.
    """
    stu_views = student_out.chunk(2)                                  # loss.py:126
    tea_views = F.softmax(teacher_out, dim=1).detach().chunk(2)       # loss.py:127-128
    total = 0
    for iq, q in enumerate(tea_views):                                # loss.py:130
        for v, s in enumerate(stu_views):                             # loss.py:131
            if v == iq:                                               # loss.py:132-134
                continue
            per_px = torch.sum(-q * F.log_softmax(s, dim=1), dim=1)   # loss.py:138
            if iq == 1:                                               # loss.py:139-140
                per_px = per_px * scale
            total = total + per_px.mean()                             # loss.py:142
    return total


def distillation_grad_closed_form(teacher_out, student_out, scale=0.5, upstream=1.0):
    """Analytic d(loss)/d(student) of :func:`distillation_loss`, in float64.

    ``dL/ds_view1 = (softmax(s1) - p0) / (B*H*W)``; ``dL/ds_view0 = scale * (softmax(s0) - p1) / (B*H*W)``.
    """
    t = teacher_out.double()
    s = student_out.double()
    p0, p1 = F.softmax(t, dim=1).chunk(2)
    s0, s1 = s.chunk(2)
    denom = s0.shape[0] * s0.shape[2] * s0.shape[3]
    g0 = scale * (F.softmax(s0, dim=1) - p1) / denom
    g1 = (F.softmax(s1, dim=1) - p0) / denom
    return torch.cat([g0, g1]) * upstream


# --------------------------------------------------------------------------------------
# helper   one-hot with overflow channel   (G/util/utils.py:158-163)
# --------------------------------------------------------------------------------------
def process_label(label: torch.Tensor, class_numbers: int = 19) -> torch.Tensor:
    """``[B,1,H,W]`` float labels -> ``[B,C+1,H,W]`` one-hot; ids >= C land in channel C.

    G/util/utils.py:158-163 (the reference allocates with ``.cuda()``; here the label's device).
    """
    b, _, h, w = label.size()                                         # utils.py:159
    onehot = torch.zeros(b, class_numbers + 1, h, w, device=label.device)          # utils.py:160
    ids = torch.where(label < class_numbers, label,
                      torch.tensor([float(class_numbers)], device=label.device))   # utils.py:161
    return onehot.scatter_(1, ids.long(), 1)                          # utils.py:162


# --------------------------------------------------------------------------------------
# a5/a6/a7  prototypes   (G/calc_centroids.py:84-180)
# --------------------------------------------------------------------------------------
class ClassFeaturesOracle:
    """Restatement of ``Class_Features`` (G/calc_centroids.py:84-180).

    State: ``objective_vectors [C,D]`` and ``objective_vectors_num [C]`` (both fp32), momentum 1e-4
    (calc_centroids.py:90-92).  ``feat_dim`` is 256 in ``G/``/``S/`` and 512 in ``SS/``.
    """

    def __init__(self, numbers: int = 19, feat_dim: int = 256):
        self.class_numbers = numbers
        self.objective_vectors = torch.zeros([numbers, feat_dim])     # calc_centroids.py:90
        self.objective_vectors_num = torch.zeros([numbers])           # calc_centroids.py:91
        self.centroid_momentum = 0.0001                               # calc_centroids.py:92

    # -- a6 ---------------------------------------------------------------------------
    def _selection_onehot(self, outputs, labels_val):
        """calc_centroids.py:121-127: one-hot of argmax(softmax(out)), optionally AND one-hot(labels)."""
        am = F.softmax(outputs, dim=1).argmax(dim=1, keepdim=True)    # :121-122
        pred = process_label(am.float(), self.class_numbers)          # :123
        if labels_val is not None:                                    # :124-127
            pred = process_label(labels_val, self.class_numbers) * pred
        return pred

    def calculate_mean_vector(self, feat_cls, outputs, labels_val=None, model=None):
        """Per image, per class masked feature mean; calc_centroids.py:120-145.

        Returns ``(vectors, ids)``: vectors are ``[D,1,1]`` tensors ordered by (image, class) with
        classes of <5 selected pixels skipped.
        """
        pred = self._selection_onehot(outputs, labels_val)
        frac = F.adaptive_avg_pool2d(pred, 1)                         # :129
        vectors, ids = [], []
        for n in range(feat_cls.size(0)):                             # :132
            for t in range(self.class_numbers):                       # :133
                if frac[n][t].item() == 0:                            # :134
                    continue
                if (pred[n][t] > 0).sum() < 5:                        # :136
                    continue
                masked = feat_cls[n] * pred[n][t]                     # :138
                vectors.append(F.adaptive_avg_pool2d(masked, 1) / frac[n][t])   # :141
                ids.append(t)
        return vectors, ids

    def calculate_mean_vector_by_output(self, feat_cls, outputs):
        """calc_centroids.py:97-118 — identical to the label-free branch above."""
        return self.calculate_mean_vector(feat_cls, outputs, None)

    # -- a7 ---------------------------------------------------------------------------
    def update_objective_SingleVector(self, id, vector, name="moving_average", start_mean=True):
        """Running mean / EMA of one class centroid; calc_centroids.py:147-164."""
        if isinstance(vector, np.ndarray):
            vector = torch.from_numpy(vector).to(self.objective_vectors.device)
        if vector.sum().item() == 0:                                  # :148
            return
        if start_mean and self.objective_vectors_num[id].item() < 100:   # :150
            name = "mean"
        if name == "moving_average":                                  # :152-156
            m = self.centroid_momentum
            self.objective_vectors[id] = self.objective_vectors[id] * (1 - m) + m * vector.squeeze()
            self.objective_vectors_num[id] += 1
            self.objective_vectors_num[id] = min(self.objective_vectors_num[id], 3000)
        elif name == "mean":                                          # :157-161
            self.objective_vectors[id] = self.objective_vectors[id] * self.objective_vectors_num[id] + vector.squeeze()
            self.objective_vectors_num[id] += 1
            self.objective_vectors[id] = self.objective_vectors[id] / self.objective_vectors_num[id]
            self.objective_vectors_num[id] = min(self.objective_vectors_num[id], 3000)
        else:                                                         # :163-164
            raise NotImplementedError("no such updating way of objective vectors {}".format(name))

    # -- a5 ---------------------------------------------------------------------------
    def feat_centroid_distance(self, feat):
        """``dist[n,c,y,x] = || centroid_c - feat[n,:,y,x] ||_2``; calc_centroids.py:166-171."""
        n, _, h, w = feat.shape
        out = -torch.ones((n, self.class_numbers, h, w), device=feat.device)      # :168
        for i in range(self.class_numbers):                                        # :169
            proto = self.objective_vectors[i].detach().to(feat.device).reshape(-1, 1, 1).expand(-1, h, w)
            out[:, i, :, :] = torch.norm(proto - feat, 2, dim=1)                   # :170
        return out

    def get_centroid_weight(self, feat):
        """``softmax(-dist)`` over classes; calc_centroids.py:173-176."""
        return F.softmax(-self.feat_centroid_distance(feat), dim=1)

    def get_centroid_distance(self, feat):
        """calc_centroids.py:178-180."""
        return -self.feat_centroid_distance(feat)


def nearest_labels_to_feature_grid(labels_i64: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """``[B,H,W]`` int64 -> ``[B,1,h,w]`` fp32 by nearest interpolation.

    G/train_DiGA_gta2city_self_training.py:327-330 (target) and :336-337 (source).
    """
    b, h, w = labels_i64.size()
    lab = labels_i64.clone().reshape([b, 1, h, w]).float()
    return F.interpolate(lab, size=tuple(size), mode="nearest")


# --------------------------------------------------------------------------------------
# a3  pseudo-label generation   (G/pseudolabel_generator.py:77-85)
# --------------------------------------------------------------------------------------
def upsample_bilinear_ac(x: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """``nn.Upsample(size, mode='bilinear', align_corners=True)``; pseudolabel_generator.py:55,
    train_DiGA_gta2city_self_training.py:190-192."""
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=True)


def pseudo_label_from_logits(output: torch.Tensor, output_ds: Optional[torch.Tensor] = None
                             ) -> Tuple[np.ndarray, np.ndarray]:
    """Two-scale fusion by elementwise max, softmax, then numpy argmax / max on the host.

    ``output`` / ``output_ds`` are the already up-sampled ``[1,C,H,W]`` logits.  Returns
    ``(label int64 [H,W], confidence fp32 [H,W])`` for image 0, exactly as
    pseudolabel_generator.py:80-85 does (the reference discards the confidence).
    """
    if output_ds is not None:
        output = torch.max(output_ds, output)                         # :80
    prob = F.softmax(output, dim=1)                                   # :81
    prob = prob.cpu().data[0].numpy().transpose(1, 2, 0)              # :82-83
    return np.argmax(prob, axis=2), np.max(prob, axis=2)              # :85


def pseudo_label_two_scale(logits_full: torch.Tensor, logits_ds: torch.Tensor, size=(1024, 2048)):
    """pseudolabel_generator.py:77-85 including the two bilinear up-samplings (:77-78)."""
    return pseudo_label_from_logits(upsample_bilinear_ac(logits_full, size), upsample_bilinear_ac(logits_ds, size))


def pseudo_label_to_uint8(label: np.ndarray) -> np.ndarray:
    """pseudolabel_generator.py:92 — labels go through a float64 staging array then to uint8."""
    return np.asarray(label.astype(np.float64), dtype=np.uint8)


def colorize_mask(mask: np.ndarray, palette):
    """pseudolabel_generator.py:45-49 — 'P'-mode PIL image, palette index = trainId."""
    from PIL import Image
    new_mask = Image.fromarray(mask.astype(np.uint8)).convert('P')    # :47
    new_mask.putpalette(palette)                                       # :48
    return new_mask


# --------------------------------------------------------------------------------------
# a4  bilateral-consensus ("threshold-free dynamic") selection
#     (G/train_DiGA_gta2city_self_training.py:298-304)
# --------------------------------------------------------------------------------------
def consensus_select(pseudo_prob: torch.Tensor, feat_weights_lowres: torch.Tensor, out_size: Sequence[int]
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Keep a stored pseudo-label only where it equals the argmax of the up-sampled prototype weights.

    ``pseudo_prob [B,H,W]`` int64, ``feat_weights_lowres [B,C,h,w]`` fp32 (output of
    ``get_centroid_weight``).  Returns ``(tlabelv_pseudo, feat_pseudo)`` both int64 ``[B,H,W]``.
    """
    kept = pseudo_prob.clone()                                        # :299
    up = upsample_bilinear_ac(feat_weights_lowres, out_size)          # :302
    feat_pseudo = up.max(1, keepdim=True)[1].squeeze(1)               # :303
    kept[pseudo_prob != feat_pseudo] = IGNORE                         # :304
    return kept, feat_pseudo


# --------------------------------------------------------------------------------------
# a2  cross-domain ClassMix
#     (G/train_DiGA_gta2city_self_training.py:259-275 image only, :306-325 DACS image + label)
# --------------------------------------------------------------------------------------
def classmix_select_classes(slabel: torch.Tensor, rng=_random) -> List[List[int]]:
    """Per image: half of the classes present (``random.sample`` on the sorted unique list), plus 255.

    self_training.py:264-268.  ``rng`` needs a ``sample`` method (the ``random`` module by default).
    """
    chosen = []
    for idx in range(slabel.size(0)):
        present = torch.unique(slabel[idx]).tolist()                  # :265
        sel = rng.sample(present, len(present) // 2)                  # :266
        if IGNORE not in sel:                                         # :267-268
            sel.append(IGNORE)
        chosen.append(sel)
    return chosen


def classmix(slabel: torch.Tensor, a: torch.Tensor, b: torch.Tensor, tlabel: Optional[torch.Tensor] = None,
             rng=_random, classes: Optional[List[List[int]]] = None):
    """ClassMix mask build + blend.  ``a`` is pasted onto (mask 0), ``b`` supplies the selected classes (mask 1).

    Image-only variant: self_training.py:259-275 (``a = rec_s2t``, ``b = sdatav_aug``), also
    warm_up.py:240-259 and calc_centroids.py:47-58.  With ``tlabel`` the DACS variant :306-325
    (``a = tdatav_aug``, ``b = sdatav``, labels mixed too).  Returns ``(mask, mix)`` or
    ``(mask, mix, mixlabel)``; ``mix`` is ``None`` when every source label is 255 (:271 / :321 —
    the reference never creates the tensor in that case).
    """
    if classes is None:
        classes = classmix_select_classes(slabel, rng)
    mask = torch.zeros(slabel.size(), device=slabel.device)           # :263 / :311
    mixlabel = tlabel.clone() if tlabel is not None else None         # :308
    for idx, sel in enumerate(classes):
        for cls_m in sel:                                             # :269-270 / :318-320
            if mixlabel is not None:
                mixlabel[idx][slabel[idx] == cls_m] = cls_m
            mask[idx][slabel[idx] == cls_m] = 1
    mix = None
    if not torch.all(torch.eq(slabel, IGNORE)):                       # :271 / :321
        mix = torch.zeros(a.size(), device=a.device)
        for idx in range(a.size(0)):                                  # :274-275 / :323-324
            mix[idx] = torch.mul(a[idx], 1 - mask[idx]) + torch.mul(b[idx], mask[idx])
        if mixlabel is not None:
            mixlabel = mixlabel.long()                                # :325
    if tlabel is None:
        return mask, mix
    return mask, mix, mixlabel


# --------------------------------------------------------------------------------------
# f2  cross_entropy2d   (G/util/loss.py:48-62)        f4  EMA teacher   (G/util/utils.py:103-116)
# --------------------------------------------------------------------------------------
def cross_entropy2d(input, target, weight=None, size_average=True):
    """Pixel-wise CE with ignore_index 255; negative targets dropped; the mean divides by #(target >= 0).
    G/util/loss.py:48-62."""
    n, c, h, w = input.size()
    log_p = F.log_softmax(input, dim=1)                                               # :50
    log_p = log_p.transpose(1, 2).transpose(2, 3).contiguous().view(-1, c)            # :51
    keep = target.contiguous().view(n * h * w, 1).repeat(1, c) >= 0                   # :52
    log_p = log_p[keep].view(-1, c)                                                   # :52-54
    mask = target >= 0                                                                # :56
    loss = F.nll_loss(log_p, target[mask], ignore_index=IGNORE, weight=weight, reduction="sum")   # :58-59
    if size_average:                                                                  # :60-61
        loss = loss / mask.data.sum()
    return loss


def distillation_loss_upsampled(teacher_low, student_low, size, scale=0.5):
    """``distillation_loss(upsample_src(tea), upsample_src(stu))``: train_DiGA_gta2city_self_training.py:289,:351-352
    (bilinear ``align_corners=True`` up-sampling in front of G/util/loss.py:125-143)."""
    return distillation_loss(upsample_bilinear_ac(teacher_low, size), upsample_bilinear_ac(student_low, size), scale)


def cross_entropy2d_upsampled(input_low, target, weight=None, size_average=True):
    """``seg_loss(upsample(pred), label)``: train_DiGA_gta2city_self_training.py:344,:348-349,:355."""
    return cross_entropy2d(upsample_bilinear_ac(input_low, target.shape[-2:]), target, weight, size_average)


def seg_distillation_losses_upsampled(teacher_low, student_low, target, scale=0.5, weight=None, size_average=True):
    """The two losses that share ``s_pred_cat_stu`` (train_DiGA_gta2city_self_training.py:348-352):
    ``seg_loss(upsample_src(s_pred_cat_stu[:B]), slabelv)`` and ``distillation_loss(upsample_src(tea), upsample_src(stu))``."""
    size = target.shape[-2:]
    s_pred_stu = upsample_bilinear_ac(student_low[:target.shape[0]], size)            # :348 (its own up-sampling)
    loss_semseg = cross_entropy2d(s_pred_stu, target, weight, size_average)           # :349
    s_pred_cat_stu = upsample_bilinear_ac(student_low, size)                          # :351
    return loss_semseg, distillation_loss(upsample_bilinear_ac(teacher_low, size), s_pred_cat_stu, scale)   # :289,:352


class OhemCrossEntropyOracle(torch.nn.Module):
    """``OhemCrossEntropy`` restated (G/util/loss.py:65-122; single-score form, ``weights = [1]``)."""

    def __init__(self, ignore_label=255, thres=0.7, min_kept=100000, weight=None):
        super().__init__()
        self.thresh = thres                                                                   # :69
        self.min_kept = max(1, min_kept)                                                      # :70
        self.ignore_label = ignore_label
        self.criterion = torch.nn.CrossEntropyLoss(weight=weight, ignore_index=ignore_label, reduction="none")   # :72-76

    def forward(self, score, target):
        ph, pw = score.size(2), score.size(3)                                                 # :91
        h, w = target.size(1), target.size(2)
        if ph != h or pw != w:                                                                # :93-95
            score = F.interpolate(input=score, size=(h, w), mode="bilinear", align_corners=True)
        pred = F.softmax(score, dim=1)                                                        # :96
        pixel_losses = self.criterion(score, target).contiguous().view(-1)                    # :97
        mask = target.contiguous().view(-1) != self.ignore_label                              # :98
        tmp_target = target.clone()                                                           # :100-101
        tmp_target[tmp_target == self.ignore_label] = 0
        pred = pred.gather(1, tmp_target.unsqueeze(1))                                        # :102
        pred, ind = pred.contiguous().view(-1,)[mask].contiguous().sort()                     # :103
        min_value = pred[min(self.min_kept, pred.numel() - 1)]                                # :104
        threshold = max(min_value, self.thresh)                                               # :105
        pixel_losses = pixel_losses[mask][ind]                                                # :107
        pixel_losses = pixel_losses[pred < threshold]                                         # :108
        return pixel_losses.mean()                                                            # :109


def ema_alpha(iteration, stage0=True, mean=False, replace=False):
    """G/util/utils.py:105-112."""
    if stage0 == True:      # noqa: E712
        return min(1 - 1 / (iteration + 1), 0.999)
    if mean == True:        # noqa: E712
        return 0.9
    if replace == True:     # noqa: E712
        return 0.0
    return 0.999


def update_teacher_params(teacher, student, iteration, stage0=True, mean=False, replace=False):
    """EMA of the student's parameters into the teacher; G/util/utils.py:103-116."""
    alpha = ema_alpha(iteration, stage0, mean, replace)
    for tp, sp in zip(teacher.parameters(), student.parameters()):                     # :113
        tp.data[:] = alpha * tp[:].data[:] + (1 - alpha) * sp[:].data[:]               # :115
    return teacher


# --------------------------------------------------------------------------------------
# f5  evaluation confusion matrix   (G/util/metrics.py:26-76)
# --------------------------------------------------------------------------------------
class RunningScoreOracle:
    """``runningScore`` restated (G/util/metrics.py:26-76) without the per-class ``print`` of :62-63."""

    def __init__(self, n_classes):
        self.n_classes = n_classes
        self.confusion_matrix = np.zeros((n_classes, n_classes))                       # :30

    def _fast_hist(self, label_true, label_pred, n_class):
        mask = (label_true >= 0) & (label_true < n_class)                                # :33
        return np.bincount(n_class * label_true[mask].astype(int) + label_pred[mask],   # :34-36
                           minlength=n_class ** 2).reshape(n_class, n_class)

    def update(self, label_trues, label_preds):
        for lt, lp in zip(label_trues, label_preds):                                     # :40-41
            self.confusion_matrix += self._fast_hist(lt.flatten(), lp.flatten(), self.n_classes)

    def get_scores(self):
        hist = self.confusion_matrix
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(hist).sum() / hist.sum()                                       # :51
            acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))                       # :52-53
            iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))   # :54
            mean_iu = np.nanmean(iu)                                                     # :57
            freq = hist.sum(axis=1) / hist.sum()                                         # :58
            fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()                              # :59
        cls_iu = dict(zip(range(self.n_classes), iu))                                    # :60
        return {'Overall Acc: \t': acc, 'Mean Acc : \t': acc_cls, 'FreqW Acc : \t': fwavacc, 'Mean IoU : \t': mean_iu}, cls_iu

    def reset(self):
        self.confusion_matrix = np.zeros((self.n_classes, self.n_classes))               # :76


# --------------------------------------------------------------------------------------
# f3 reader half: label maps as the loader delivers them   (G/util/loader/CityLoader.py:86-95, :113-131)
# --------------------------------------------------------------------------------------
CITY_ID_TO_TRAINID = {7: 0, 8: 1, 11: 2, 12: 3, 13: 4, 17: 5, 19: 6, 20: 7, 21: 8, 22: 9, 23: 10, 24: 11, 25: 12, 26: 13,
                      27: 14, 28: 15, 31: 16, 32: 17, 33: 18}                       # CityLoader.py:49-56


def city_loader_labels(label_img, pseudo_img, crop_size=None, n_classes=19, id_to_trainid=None):
    """``CityLoader.__getitem__`` restricted to the two label maps, PIL images in, int64 numpy maps out: NEAREST resize
    to ``crop_size`` = (H, W) (:92-95), ground-truth ids -> trainIds (:113-118), pseudo-label ids >= n_classes -> 255
    (:123, :129-131).  ``np.compat.long`` of the reference is int64."""
    from PIL import Image
    id_to_trainid = id_to_trainid or CITY_ID_TO_TRAINID
    if crop_size is not None:                                                          # :92
        label_img = label_img.resize((crop_size[1], crop_size[0]), Image.NEAREST)     # :94
        pseudo_img = pseudo_img.resize((crop_size[1], crop_size[0]), Image.NEAREST)   # :96
    label = np.asarray(label_img, np.int64)                                            # :122
    pseudo_label = np.asarray(pseudo_img, np.int64)                                    # :123
    label_copy = 255 * np.ones(label.shape, dtype=np.int64)                            # :125
    for k, v in id_to_trainid.items():                                                 # :126-127
        label_copy[label == k] = v
    pseudo_label_copy = 255 * np.ones(pseudo_label.shape, dtype=np.int64)              # :129
    for v in range(n_classes):                                                         # :130-131
        pseudo_label_copy[pseudo_label == v] = v
    return label_copy, pseudo_label_copy


# --------------------------------------------------------------------------------------
# drivers used as CPU baseline / multi-rank checks
# --------------------------------------------------------------------------------------
def centroid_pass(feats: Sequence[torch.Tensor], outs: Sequence[torch.Tensor], numbers=19, feat_dim=256,
                  cf: Optional[ClassFeaturesOracle] = None) -> ClassFeaturesOracle:
    """The target-domain loop of ``calc_centroids`` (G/calc_centroids.py:67-78) over in-memory batches."""
    cf = cf or ClassFeaturesOracle(numbers, feat_dim)
    for feat, out in zip(feats, outs):
        vectors, ids = cf.calculate_mean_vector(feat, out)            # :75
        for t in range(len(ids)):                                     # :77-78
            cf.update_objective_SingleVector(ids[t], vectors[t].detach().cpu().numpy(), "mean")
    return cf
