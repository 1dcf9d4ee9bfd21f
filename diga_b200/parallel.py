"""Image-sharded data parallelism for the hot path.

One process per GPU (``torchrun``).  a1–a6 have no cross-image dependency, so ranks work on their own images
independently; the only exchange is a7, the running update of the class centroids (calc_centroids.py:147-164), which the
reference applies image after image.  Two modes, and every result says which one produced it:

**exact** (:class:`ShardedCentroidPass`, ``Class_Features.update_from_features_sharded``) — every rank computes the
per-image class vectors of its own loader batches (``batches[rank::world]``), ONE ``all_gather`` moves the ``[images, C, D]``
vectors (2975 x 19 x 2048 fp32 = 463 MB once per pass, over NVLink) and every rank replays the reference recurrence
over all images in GLOBAL loader order on the device (``diga_centroid_update_sharded``).  Bit-identical to the single-
process sequence — beyond the 3000 clamp (:156,:161), across the reference's five passes (:20-23), and for the online EMA.

**sum** (``accumulate_mean_pass`` + :func:`finish_mean_pass`) — every rank accumulates ``acc [C, D+1]`` = (sum of its
per-image class vectors, their number) and ONE ``all_reduce`` of 19 x 2049 fp32 finishes the pass.  Equals the reference's
running mean while ``num + n <= 3000`` (the first pass over 2975 images); beyond the clamp the reference recursion
``obj = (3000 obj + v) / 3001`` weights later images more, and this mode replaces those weights by their average
(documented approximation; the exact mode exists for parity).

Collectives run through ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

CLAMP = 3000.0          # calc_centroids.py:156,161


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(n_items: int, rank: int, world: int):
    """Indices of the items rank ``rank`` owns: ``range(n_items)[rank::world]``."""
    return list(range(n_items))[rank::world]


def global_row_order(n_total: int, batch: int, world: int, per_shard: int):
    """Host mirror of ``order_row`` in ``csrc/centroid.cu``: row of the gathered ``[world * per_shard]`` buffer that holds the
    g-th image of the global loader sequence, when rank r processed batches r, r + world, ... of ``batch`` images each."""
    rows = []
    for g in range(n_total):
        k, j = divmod(g, batch)
        rows.append((k % world) * per_shard + (k // world) * batch + j)
    return rows


# ------------------------------------------------------------------------------------------------------------------
# sum mode: one all-reduce of [C, D+1]
# ------------------------------------------------------------------------------------------------------------------
def new_mean_accumulator(class_numbers: int, feat_dim: int, device) -> torch.Tensor:
    return torch.zeros((class_numbers, feat_dim + 1), dtype=torch.float32, device=device)


def finish_mean_pass(acc: torch.Tensor, class_features=None, group=None):
    """All-reduce ``acc`` (in place, SUM) and return ``(objective_vectors [C,D], objective_vectors_num [C])``.

    Without ``class_features`` the pass starts from empty centroids: ``objective_vectors[c] = sum / max(n, 1)``,
    ``objective_vectors_num[c] = min(n, 3000)`` — the reference's first pass.  With ``class_features`` the pass continues
    from its state ``(obj, num)`` and the result is written back into it: the first ``m = min(n, 3000 - num)`` vectors enter
    as the running mean (``(obj num + m mean) / (num + m)``, exact), the remaining ``n - m`` as the clamped recursion with
    uniform weights (``obj rho^k + (1 - rho^k) mean``, ``rho = 3000/3001``, ``k = n - m``) — see the module docstring.
    Works on CUDA tensors with the NCCL backend and on CPU tensors with gloo (host-logic tests).  No host sync.
    """
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    d = acc.shape[1] - 1
    n = acc[:, d]
    mean = acc[:, :d] / n.clamp(min=1.0).unsqueeze(1)
    if class_features is None:
        return mean, n.clamp(max=CLAMP)
    obj = class_features.objective_vectors.to(acc.device)
    num = class_features.objective_vectors_num.to(acc.device)
    m = torch.minimum(n, (CLAMP - num).clamp(min=0.0))                  # vectors that still enter as a plain mean
    k = n - m                                                            # vectors applied at the clamp
    num1 = num + m
    obj1 = (obj * num.unsqueeze(1) + mean * m.unsqueeze(1)) / num1.clamp(min=1.0).unsqueeze(1)
    decay = torch.pow(torch.full_like(k, CLAMP / (CLAMP + 1.0)), k).unsqueeze(1)
    obj2 = torch.where((n > 0).unsqueeze(1), obj1 * decay + (1.0 - decay) * mean, obj)
    num2 = (num + n).clamp(max=CLAMP)
    class_features.objective_vectors = obj2
    class_features.objective_vectors_num = num2
    return class_features.objective_vectors, class_features.objective_vectors_num


# ------------------------------------------------------------------------------------------------------------------
# exact mode: all-gather of the per-image vectors + ordered replay
# ------------------------------------------------------------------------------------------------------------------
class ShardedCentroidPass:
    """One pass of ``calc_centroids`` (calc_centroids.py:67-78) over ``n_images`` images in loader batches of ``batch``,
    sharded ``batches[rank::world]``, finished by an all-gather and an ordered device replay that leaves EVERY rank's
    ``class_features`` in exactly the state the single-process loop reaches.  Re-usable: ``finish()`` resets the buffers,
    the centroid state carries over to the next pass (the reference runs five, :20-23).

        sp = ShardedCentroidPass(cf, n_images=2975, batch=1)
        for k in sp.my_batches():                     # k = global batch index
            sp.add(feat_k, out_k)                     # [b, D, h, w], [b, C, h, w] of batch k
        sp.finish()                                   # all ranks now hold identical centroids

    ``feat_dim`` is taken from ``class_features.objective_vectors``.  Memory: ``ceil(batches / world) * batch`` rows of
    ``C x D`` fp32 per rank plus the gathered ``world`` x that (463 MB at 2975 x 19 x 2048).
    """

    def __init__(self, class_features, n_images: int, batch: int = 1, name: str = "mean", start_mean: bool = True, group=None,
                 device=None, rank=None, world=None):
        self.cf, self.group = class_features, group
        self.rank, self.world = _world(group)
        if rank is not None or world is not None:         # explicit placement (tests build the shards of several ranks in one
            self.rank, self.world = int(rank), int(world)   # process; gather() then needs a matching process group)
        self.n_images, self.batch = int(n_images), int(batch)
        if self.n_images < 0 or self.batch < 1:
            raise ValueError("ShardedCentroidPass: n_images >= 0 and batch >= 1 required")
        self.name, self.start_mean = name, start_mean
        self.batches = -(-self.n_images // self.batch)
        self.per_shard = -(-self.batches // self.world) * self.batch
        c, d = class_features.class_numbers, int(class_features.objective_vectors.shape[1])
        dev = class_features.objective_vectors.device if device is None else torch.device(device)
        rows = max(self.per_shard, 1)
        self.vec = torch.empty((rows, c, d), dtype=torch.float32, device=dev)
        self.vecsum = torch.zeros((rows, c), dtype=torch.float32, device=dev)      # rows never filled must read as "skip"
        self.valid = torch.zeros((rows, c), dtype=torch.uint8, device=dev)
        self._gathered = None
        self._local_batches = 0

    def my_batches(self):
        """Global indices of the loader batches this rank processes, in the order ``add`` expects them."""
        return range(self.batches)[self.rank::self.world]

    def batch_size_of(self, k: int) -> int:
        """Images in global batch ``k`` (the last batch of the set may be short)."""
        return min(self.batch, self.n_images - k * self.batch)

    def _next_rows(self, n: int):
        k = self.rank + self._local_batches * self.world
        if k >= self.batches:
            raise RuntimeError("ShardedCentroidPass: more batches added than this rank owns")
        if n != self.batch_size_of(k):
            raise ValueError(f"ShardedCentroidPass: batch {k} must hold {self.batch_size_of(k)} images, got {n}")
        r0 = self._local_batches * self.batch
        self._local_batches += 1
        return slice(r0, r0 + n)

    def add_rows(self, vec, vecsum, valid):
        """Append precomputed per-image rows ``vec [n,C,D]``, ``vecsum [n,C]``, ``valid [n,C]`` of this rank's next batch."""
        rows = self._next_rows(vec.shape[0])
        self.vec[rows].copy_(vec)
        self.vecsum[rows].copy_(vecsum)
        self.valid[rows].copy_(valid)

    def add(self, feat_cls, outputs, labels_val=None, labels_full=None):
        """a6 of this rank's next batch (``calculate_mean_vector``, calc_centroids.py:120-145), written straight into the
        pass buffer; no host sync."""
        rows = self._next_rows(feat_cls.shape[0])
        self.cf._masked_means(feat_cls, outputs, labels_val, labels_full, out_rows=(self.vec[rows], self.vecsum[rows], self.valid[rows]))

    def gather(self):
        """The exchange: all-gather of the three row buffers -> ``[world * per_shard, ...]`` on every rank."""
        if self.world == 1:
            return self.vec, self.vecsum, self.valid
        if self._gathered is None:
            self._gathered = tuple(torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                                   for t in (self.vec, self.vecsum, self.valid))
        for dst, src in zip(self._gathered, (self.vec, self.vecsum, self.valid)):
            dist.all_gather_into_tensor(dst, src, group=self.group)
        return self._gathered

    def finish(self):
        """All-gather + ordered replay into ``class_features`` (identical on every rank), then reset for the next pass."""
        gvec, gsum, gvalid = self.gather()
        self.cf._update_sharded(gvec, gsum, gvalid, self.n_images, self.batch, self.world, max(self.per_shard, 1), self.name,
                                self.start_mean)
        self.reset()
        return self.cf.objective_vectors, self.cf.objective_vectors_num

    def reset(self):
        self.vecsum.zero_()
        self.valid.zero_()
        self._local_batches = 0
