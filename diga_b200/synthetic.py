"""Seeded synthetic Cityscapes-shaped inputs for the tests and ``bench.py``.

No dataset or checkpoint exists offline, so every workload is generated: logits ``3*randn``, features
``randn``, centroids ``0.5*randn + mean(feat)``, piecewise-constant int64 label maps (32x32 blocks, 10 %
ignore), images ``randn`` clamped to [-1, 1].  Generators take a device so the same tensors can be made on
the GPU (bench) or on the CPU (oracle side of a parity test) from the same seed *per device type*; parity
tests generate once and copy.
"""
from __future__ import annotations

import torch


def feature_hw(h: int, w: int):
    """Stride-8 feature resolution of the reference ResNet: 512x1024 -> 65x129."""
    def one(x):
        s1 = (x - 1) // 2 + 1
        s2 = -(-(s1 - 1) // 2) + 1
        return (s2 - 1) // 2 + 1
    return one(h), one(w)


def gen(seed: int, device="cpu") -> torch.Generator:
    return torch.Generator(device=device).manual_seed(seed)


def logits(shape, g: torch.Generator, sigma: float = 3.0) -> torch.Tensor:
    return sigma * torch.randn(shape, generator=g, device=g.device, dtype=torch.float32)


def logits_peaked(shape, g: torch.Generator, block: int = 6, margin: float = 8.0, sigma: float = 1.0) -> torch.Tensor:
    """Network-like logits ``[N,C,h,w]``: a piecewise-constant arg-max map (``block`` x ``block`` regions at the logit
    resolution) leads by ``margin`` over Gaussian noise of scale ``sigma`` — confident inside regions, contested along their
    borders, which is what a trained segmentation head emits.  The i.i.d. ``logits`` above are the worst case of
    the arg-max kernels' candidate pruning (nearly every class can win in every source cell); this is the realistic case."""
    n, c, h, w = shape
    lab = block_labels(n, h, w, g, block, c, 0.0)
    onehot = torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).to(torch.float32)
    return sigma * torch.randn(shape, generator=g, device=g.device, dtype=torch.float32) + margin * onehot


def features(shape, g: torch.Generator) -> torch.Tensor:
    return torch.randn(shape, generator=g, device=g.device, dtype=torch.float32)


def centroids(c: int, d: int, g: torch.Generator, feat_mean: float = 0.0) -> torch.Tensor:
    return 0.5 * torch.randn((c, d), generator=g, device=g.device, dtype=torch.float32) + feat_mean


def images(shape, g: torch.Generator) -> torch.Tensor:
    return torch.randn(shape, generator=g, device=g.device, dtype=torch.float32).clamp_(-1, 1)


def block_labels(b: int, h: int, w: int, g: torch.Generator, block: int = 32, n_cls: int = 19, p_ignore: float = 0.1
                 ) -> torch.Tensor:
    """Segmentation-like int64 label maps: constant ``block`` x ``block`` tiles, ``p_ignore`` of them 255."""
    gh, gw = -(-h // block), -(-w // block)
    coarse = torch.randint(0, n_cls, (b, gh, gw), generator=g, device=g.device)
    coarse[torch.rand((b, gh, gw), generator=g, device=g.device) < p_ignore] = 255
    lab = coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :h, :w]
    return lab.contiguous().long()


def perturb_labels(lab: torch.Tensor, g: torch.Generator, block: int = 32, n_cls: int = 19, p: float = 0.2) -> torch.Tensor:
    """Pseudo-labels = labels with a fraction ``p`` of the tiles re-drawn."""
    b, h, w = lab.shape
    other = block_labels(b, h, w, g, block, n_cls, 0.0)
    gh, gw = -(-h // block), -(-w // block)
    flip = (torch.rand((b, gh, gw), generator=g, device=g.device) < p)
    flip = flip.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :h, :w]
    return torch.where(flip, other, lab)
