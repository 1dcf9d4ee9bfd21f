"""Drop-in for ``util/metrics.py`` of the reference (``runningScore``, :26-76), confusion matrix kept on the GPU.

Same class name, constructor, ``update(label_trues, label_preds)``, ``get_scores()`` (same dictionary keys, same per-class
``print`` lines) and ``reset()``.  The reference moves every prediction and ground-truth map to the host and runs
``np.bincount`` per image (:32-41); here ``update`` takes CUDA tensors as they come out of ``argmax`` /
``pseudo_label_two_scale`` (uint8 or int64; numpy arrays are accepted and uploaded) and one kernel per call accumulates
the ``n x n`` int64 matrix on the device (csrc/metrics.cu).  Only ``get_scores`` reads it back (n*n*8 bytes).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L

label = ['road', 'sidewalk', 'building', 'wall', 'fence', 'pole', 'light', 'sign', 'vegetation', 'terrain', 'sky',
         'person', 'rider', 'car', 'truck', 'bus', 'train', 'motorcycle', 'bycycle']        # metrics.py:6-24 (spelling kept)


class runningScore(object):
    def __init__(self, n_classes, device=None):
        self.n_classes = n_classes
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._hist = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=self.device)
        self._flags = torch.zeros((1,), dtype=torch.int32, device=self.device)

    def _as_labels(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if not torch.is_tensor(x):
            x = torch.stack([self._as_labels(v) for v in x])       # a list of per-image maps
        if not x.is_cuda:
            x = x.to(self.device, non_blocking=True)
        if x.dtype not in (torch.uint8, torch.int64):
            x = x.long()                                            # :35  .astype(int)
        return x.contiguous()

    def update(self, label_trues, label_preds):
        """``confusion_matrix += sum_images _fast_hist(true, pred)`` (:39-41); one launch for the whole batch."""
        lt, lp = self._as_labels(label_trues), self._as_labels(label_preds)
        if lt.numel() != lp.numel():
            raise ValueError(f"runningScore.update: {lt.numel()} ground-truth labels vs {lp.numel()} predictions")
        L.require_cuda(lt, lp, self._hist, what="runningScore.update input")     # one device: the matrix's
        with torch.cuda.device(self.device):
            L.check(L.lib.diga_confusion_matrix(lt.data_ptr(), int(lt.dtype == torch.uint8), lp.data_ptr(),
                                                int(lp.dtype == torch.uint8), lt.numel(), self.n_classes, self._hist.data_ptr(),
                                                self._flags.data_ptr(), L.stream()))

    @property
    def confusion_matrix(self):
        """The matrix as the reference holds it: float64 numpy ``[n, n]`` (:30).  One D2H copy + sync."""
        if int(self._flags.item()):
            raise ValueError("runningScore: a prediction outside [0, n_classes) met a counted pixel "
                             "(np.bincount(...).reshape fails there in the reference)")
        return self._hist.cpu().numpy().astype(np.float64)

    def get_scores(self):
        """metrics.py:43-73, same arithmetic on the same matrix."""
        hist = self.confusion_matrix
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(hist).sum() / hist.sum()
            acc_cls = np.diag(hist) / hist.sum(axis=1)
            acc_cls = np.nanmean(acc_cls)
            iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
            for id in range(min(19, self.n_classes)):
                print('===>' + label[id] + ':' + str(iu[id]))
            mean_iu = np.nanmean(iu)
            freq = hist.sum(axis=1) / hist.sum()
            fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
        cls_iu = dict(zip(range(self.n_classes), iu))
        return {'Overall Acc: \t': acc, 'Mean Acc : \t': acc_cls, 'FreqW Acc : \t': fwavacc, 'Mean IoU : \t': mean_iu}, cls_iu

    def reset(self):
        self._hist.zero_()
        self._flags.zero_()


def evaluate_val(student, val_loader, n_classes=19, size=(1024, 2048), running_metrics=None, preprocess=None):
    """The evaluation loop of ``evaluate_val.py:72-90`` (and ``train_DiGA_gta2city_self_training.py:428-449``) with the
    per-pixel math kept on the GPU: two forward passes (full and half resolution, :78-81), fused up-sampling + max + arg-max
    (:82-86, never materialising the two ``[1,19,1024,2048]`` tensors), confusion matrix on the device (:87-89).
    ``student(x)`` returns ``(_, _, logits, _)`` like the reference ``SegModel``; ``val_loader`` yields ``(images, labels)``.
    Returns ``running_metrics.get_scores()``.  ``preprocess``: applied to both image batches before the model — the
    semi-supervised tree needs ``lambda x: x[:, [2, 1, 0]]`` (its evaluate_val.py:76-77 flips BGR -> RGB at the call site);
    the GTA5 / Synthia trees, which this mirrors, feed the loader's image as is."""
    import torch.nn.functional as F

    from ..pseudolabel import pseudo_label_two_scale
    rs = running_metrics if running_metrics is not None else runningScore(n_classes)
    with torch.no_grad():
        for images_val, labels_val in val_loader:
            images_val = images_val.cuda(non_blocking=True)
            images_ds = F.interpolate(images_val, (size[0] // 2, size[1] // 2), mode='bilinear', align_corners=True)   # :78
            if preprocess is not None:
                images_val, images_ds = preprocess(images_val), preprocess(images_ds)
            _, _, pred, _ = student(images_val)                                                                       # :80
            _, _, pred_ds, _ = student(images_ds)                                                                     # :81
            label, _ = pseudo_label_two_scale(pred, pred_ds, size, want_conf=False)                                    # :82-86
            rs.update(labels_val.cuda(non_blocking=True), label)                                                       # :87-89
    return rs.get_scores()
