"""Drop-in for the hot-path helper of ``util/utils.py`` of the reference (``process_label``, :158-163)."""
from __future__ import annotations

import torch

from .. import _lib as L


def process_label(label, class_numbers=19):
    """``[B,1,H,W]`` float labels -> ``[B,class_numbers+1,H,W]`` fp32 one-hot; ids >= class_numbers go to the
    last channel (G/util/utils.py:158-163; the Synthia tree's default is 16)."""
    L.require_cuda(label, what="process_label input")
    batch, channel, w, h = label.size()
    if channel != 1:
        raise ValueError("process_label: expected a [B,1,H,W] label tensor")
    lab = L.f32c(label)
    out = torch.empty((batch, class_numbers + 1, w, h), dtype=torch.float32, device=label.device)
    L.check(L.lib.diga_onehot_labels(lab.data_ptr(), batch, class_numbers, w * h, out.data_ptr(), L.stream()))
    return out
