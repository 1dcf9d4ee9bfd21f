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

Collectives run through ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests of the host logic).  On
NVLink-connected GPUs the exact mode does not call a collective for its data at all: the row buffers are ONE symmetric
allocation per rank (``torch.distributed._symmetric_memory``), and the kernel that turns class sums into per-image vectors
(``diga_centroid_means_scatter``) stores every row directly into the row block its rank owns in EVERY rank's buffer — peer
stores over NVLink, issued while the accumulation of the next batch runs.  The pass then ends with a barrier and the ordered
replay; the 463 MB all-gather is gone (``DIGA_SYMM=0`` in the environment, or a failed rendezvous, falls back to it).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib as L

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
                 device=None, rank=None, world=None, symmetric: bool = True):
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
        self._gathered = None
        self._local_batches = 0
        self._symm = None
        if symmetric and self.world > 1 and dev.type == "cuda" and rank is None and os.environ.get("DIGA_SYMM", "1") != "0":
            self._symm = self._try_symmetric(rows, c, d, dev)
        if self._symm is None:
            self.vec = torch.empty((rows, c, d), dtype=torch.float32, device=dev)
            self.vecsum = torch.zeros((rows, c), dtype=torch.float32, device=dev)      # rows never filled must read as "skip"
            self.valid = torch.zeros((rows, c), dtype=torch.uint8, device=dev)

    def _try_symmetric(self, rows, c, d, dev):
        """One symmetric allocation [vec | vecsum | valid] for the GATHERED rows of all ranks; returns None when symmetric
        memory cannot be set up here (then the all-gather path is used).  Collective: every rank of the group calls it."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            total = self.world * rows
            off_vec, n_vec = 0, total * c * d * 4
            off_sum = (n_vec + 255) // 256 * 256
            off_val = (off_sum + total * c * 4 + 255) // 256 * 256
            nbytes = (off_val + total * c + 255) // 256 * 256
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            if hdl.world_size != self.world or self.world > 16:
                return None
            buf.zero_()                                   # rows no rank ever writes (shorter shards) must read as "skip"
            hdl.barrier(channel=0)
            ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            mc = int(hdl.multicast_ptr) if getattr(hdl, "has_multicast_support", False) and os.environ.get("DIGA_MULTICAST", "1") != "0" else 0
            gvec = buf[off_vec:off_vec + n_vec].view(torch.float32).view(total, c, d)
            gsum = buf[off_sum:off_sum + total * c * 4].view(torch.float32).view(total, c)
            gval = buf[off_val:off_val + total * c].view(total, c)
            return {"buf": buf, "hdl": hdl, "ptrs": ptrs, "mc": mc or None, "off": (off_vec, off_sum, off_val), "views": (gvec, gsum, gval)}
        except Exception as e:                            # noqa: BLE001  (no NVLink peer access, old torch, ...): documented fallback
            self._symm_error = f"{type(e).__name__}: {e}"
            return None

    @property
    def exchange(self) -> str:
        """How the rows travel: 'peer-stores' (fused into the means kernel over NVLink), 'all-gather', or 'local' (one rank)."""
        if self.world == 1:
            return "local"
        if self._symm is None:
            return "all-gather"
        return "multicast-stores" if self._symm["mc"] else "peer-stores"

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
        if self._symm is not None:
            raise RuntimeError("ShardedCentroidPass.add_rows needs the all-gather exchange: construct with symmetric=False "
                               "(add() writes the rows of a batch through the peer-store kernel)")
        rows = self._next_rows(vec.shape[0])
        self.vec[rows].copy_(vec)
        self.vecsum[rows].copy_(vecsum)
        self.valid[rows].copy_(valid)

    def add(self, feat_cls, outputs, labels_val=None, labels_full=None):
        """a6 of this rank's next batch (``calculate_mean_vector``, calc_centroids.py:120-145), written straight into the
        pass buffer — with symmetric memory into the row block of this rank in EVERY rank's buffer; no host sync."""
        rows = self._next_rows(feat_cls.shape[0])
        if self._symm is None:
            self.cf._masked_means(feat_cls, outputs, labels_val, labels_full, out_rows=(self.vec[rows], self.vecsum[rows], self.valid[rows]))
            return
        with torch.cuda.device(feat_cls.device):
            sums_p, counts_p, (n, c, d, hw), ws = self.cf._class_sums(feat_cls, outputs, labels_val, labels_full)
            off = self._symm["off"]
            # the scatter kernel runs on a side stream: its NVLink stores (and their acknowledgement latency) overlap the
            # assign / accumulation kernels of the next batch instead of sitting between them
            main = torch.cuda.current_stream()
            side = self._symm.setdefault("side", torch.cuda.Stream())
            side.wait_stream(main)
            with torch.cuda.stream(side):
                L.check(L.lib.diga_centroid_means_scatter(sums_p, counts_p, n, c, d, hw, self._symm["ptrs"],
                                                          self._symm["mc"], self.world, off[0], off[1], off[2],
                                                          self.rank * max(self.per_shard, 1) + rows.start, side.cuda_stream))
            ws.record_stream(side)

    def gather(self):
        """The exchange: all-gather of the three row buffers -> ``[world * per_shard, ...]`` on every rank.  With symmetric
        memory the rows are already everywhere (peer stores of the means kernel): a barrier makes them visible."""
        if self._symm is not None:
            if "side" in self._symm:
                torch.cuda.current_stream().wait_stream(self._symm["side"])
            self._symm["hdl"].barrier(channel=0)
            return self._symm["views"]
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
        if self._symm is not None:
            # nobody may start writing the next pass into a buffer another rank is still replaying from; the rows written per
            # pass are the same every pass, so there is nothing to clear
            self._symm["hdl"].barrier(channel=1)
        else:
            self.vecsum.zero_()
            self.valid.zero_()
        self._local_batches = 0
