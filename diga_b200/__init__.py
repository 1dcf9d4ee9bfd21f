"""diga_b200 — B200-native (sm_100a) implementation of DiGA's per-pixel adaptation hot path.

Host side mirrors the reference's interface for this path:

* ``diga_b200.util.loss.distillation_loss``          <- ``util/loss.py:125``
* ``diga_b200.util.utils.process_label``             <- ``util/utils.py:158``
* ``diga_b200.util.loss.cross_entropy2d``            <- ``util/loss.py:48``   (next row f2)
* ``diga_b200.util.utils.update_teacher_params``     <- ``util/utils.py:103`` (next row f4)
* ``diga_b200.util.metrics.runningScore``              <- ``util/metrics.py:26``  (next row f5)
* ``diga_b200.calc_centroids.Class_Features`` / ``calc_centroids`` <- ``calc_centroids.py:84`` / ``:17``
* ``diga_b200.classmix.classmix``                    <- inline block ``train_DiGA_gta2city_self_training.py:259-275, 306-325``
* ``diga_b200.selection.consensus_select``           <- inline block ``:298-304``
* ``diga_b200.pseudolabel.pseudo_label``             <- inline block ``pseudolabel_generator.py:77-85``
* ``diga_b200.nn.Upsample``                          <- the ``nn.Upsample(bilinear, align_corners=True)`` modules of
  ``train_DiGA_gta2city_self_training.py:190-192`` / ``pseudolabel_generator.py:55``: returns a lazy tensor that the
  functions above consume at stride-8 resolution (fused up-sampling), and that materialises for anything else

Everything runs through ``libdiga_b200.so`` (C ABI in ``include/diga_b200.h``); importing this package fails
loudly when the library has not been built.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from . import nn  # noqa: F401  (diga_b200.nn.Upsample: lazy stand-in for the scripts' nn.Upsample modules)
from .calc_centroids import Class_Features, calc_centroids
from .classmix import classmix, present_classes_async
from .pseudolabel import pseudo_label, pseudo_label_two_scale
from .selection import consensus_select
from .util.loss import (OhemCrossEntropy, cross_entropy2d, cross_entropy2d_upsampled, distillation_loss, distillation_loss_and_grad,
                        distillation_loss_upsampled, distillation_loss_upsampled_and_grad,
                        seg_distillation_losses_upsampled, seg_distillation_total_upsampled)
from .util.utils import process_label, update_teacher_params

__all__ = ["OhemCrossEntropy", "present_classes_async", "Class_Features", "calc_centroids", "classmix", "pseudo_label", "pseudo_label_two_scale",
           "consensus_select", "distillation_loss", "distillation_loss_and_grad", "process_label", "cross_entropy2d",
           "update_teacher_params", "distillation_loss_upsampled", "cross_entropy2d_upsampled",
           "seg_distillation_losses_upsampled", "seg_distillation_total_upsampled", "distillation_loss_upsampled_and_grad"]
__version__ = "0.1.0"
