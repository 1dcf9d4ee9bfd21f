"""Drop-in for ``calc_centroids.py`` of the reference: ``Class_Features`` and the ``calc_centroids`` driver.

Reference: ``domain_adaptation/GTA5/calc_centroids.py`` — class ``Class_Features`` :84-180, driver :17-81.
Names, arguments and return structures are the reference's; the work is done by the sm_100a kernels in
``csrc/centroid.cu`` (a6/a7) and ``csrc/proto.cu`` / ``csrc/proto_umma.cu`` (a5).  State lives on the GPU.

What changes for the caller: nothing in the signatures.  ``calculate_mean_vector`` costs one host sync
(the list of class ids is a Python list) instead of ~95; ``update_objective_SingleVector`` costs none.
The fused entry points ``update_from_features`` / ``accumulate_mean_pass`` (not in the reference) do the
whole a6 -> a7 chain without any host sync and are what the ``calc_centroids`` driver below uses.
"""
from __future__ import annotations

import ctypes as C
import os
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L
from .classmix import classmix

_MODES = {"mean": L.UPDATE_MEAN, "moving_average": L.UPDATE_MOVING_AVERAGE}


class Class_Features:
    """Per-class feature centroids (prototypes); reference ``calc_centroids.py:84-95``.

    Attributes kept from the reference: ``class_numbers``, ``objective_vectors [C,D]`` (D = 256 until
    re-assigned; the self-training script assigns a loaded tensor from outside, self_training.py:168-170),
    ``objective_vectors_num [C]``, ``centroid_momentum``.  Both tensors are fp32 CUDA tensors here; assigning
    a CPU tensor or array moves it to the device.
    """

    def __init__(self, numbers=19, feat_dim=256, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("diga_b200.Class_Features needs a CUDA device (there is no CPU fallback)")
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.class_numbers = numbers
        self.objective_vectors = torch.zeros([numbers, feat_dim])
        self.objective_vectors_num = torch.zeros([numbers])
        self.centroid_momentum = 0.0001
        self.valid_classes = list(range(numbers))
        self._proto_ws = None
        self._proto_key = None      # (data_ptr, tensor version, device) of the centroids the workspace was prepared from

    # -- state ----------------------------------------------------------------------------------
    def _to_state(self, value):
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(value)
        return torch.as_tensor(value).detach().to(device=self._device, dtype=torch.float32).contiguous()

    @property
    def objective_vectors(self):
        return self._objective_vectors

    @objective_vectors.setter
    def objective_vectors(self, value):
        self._objective_vectors = self._to_state(value)
        self._proto_key = None

    @property
    def objective_vectors_num(self):
        return self._objective_vectors_num

    @objective_vectors_num.setter
    def objective_vectors_num(self, value):
        self._objective_vectors_num = self._to_state(value)

    @property
    def feat_dim(self):
        return int(self._objective_vectors.shape[1])

    def _require_state_device(self, t):
        """The update kernels read the features and write the centroid state in one launch: both must be on one device."""
        if t.device != self._device:
            raise RuntimeError(f"Class_Features: the centroids live on {self._device}, the features on {t.device}; "
                               "construct Class_Features(device=...) on the device that produces the features")

    # -- a6 -------------------------------------------------------------------------------------
    def _prep(self, feat_cls, outputs, labels_val, labels_full):
        """Validated, contiguous fp32 / int64 views of the inputs of the a6 chain + its one scratch tensor."""
        L.require_cuda(feat_cls, outputs, labels_val, labels_full, what="Class_Features input")
        feat, out = L.f32c(feat_cls.detach()), L.f32c(outputs.detach())
        n, d, h, w = feat.shape
        c = self.class_numbers
        if out.shape != (n, c, h, w):
            raise ValueError(f"outputs must be [{n},{c},{h},{w}], got {tuple(out.shape)}")
        if labels_val is not None and labels_full is not None:
            raise ValueError("pass either labels_val (down-sampled fp32) or labels_full (int64), not both")
        lab = full = None
        hh = ww = 0
        if labels_val is not None:
            lab = L.f32c(labels_val.detach())
            if lab.shape != (n, 1, h, w):
                raise ValueError(f"labels_val must be [{n},1,{h},{w}], got {tuple(lab.shape)}")
        if labels_full is not None:
            full = L.i64c(labels_full)
            if full.dim() != 3 or full.shape[0] != n:
                raise ValueError(f"labels_full must be [{n},H,W] int64, got {tuple(full.shape)}")
            hh, ww = full.shape[1], full.shape[2]
        # ONE scratch tensor per call: class map, phase-shifted class words, counts, sums (and means where a path needs them)
        ws = torch.empty((max(int(L.lib.diga_centroid_chain_workspace_bytes(n, c, d, h * w)), 256),), dtype=torch.uint8, device=feat.device)
        return feat, out, lab, full, hh, ww, (n, c, d, h, w), ws

    def _class_sums(self, feat_cls, outputs, labels_val, labels_full=None):
        """assign -> accum on the current stream in one library call.  Returns ``(sums_ptr, counts_ptr, (n, c, d, hw), ws)``:
        raw device pointers of ``sums [N,C,D]`` fp32 and ``counts [N,C]`` int32 inside the scratch tensor ``ws`` (keep it alive)."""
        feat, out, lab, full, hh, ww, (n, c, d, h, w), ws = self._prep(feat_cls, outputs, labels_val, labels_full)
        sums_p, counts_p = C.c_void_p(), C.c_void_p()
        L.check(L.lib.diga_centroid_chain_sums(feat.data_ptr(), out.data_ptr(), L.ptr(lab), L.ptr(full), hh, ww, n, c, d, h, w,
                                               ws.data_ptr(), C.byref(sums_p), C.byref(counts_p), L.stream()))
        return sums_p.value, counts_p.value, (n, c, d, h * w), ws

    @L.on_device
    def _masked_means(self, feat_cls, outputs, labels_val, labels_full=None, out_rows=None):
        """assign -> accum -> means on the current stream.  Returns (vec [N,C,D], vecsum [N,C], valid [N,C]).
        ``labels_full``: the full-resolution int64 map ``[N,H,W]`` instead of its nearest-down-sampled fp32 copy.
        ``out_rows``: (vec, vecsum, valid) tensors to write into (rows of a pass buffer) instead of fresh ones."""
        sums_p, counts_p, (n, c, d, hw), ws = self._class_sums(feat_cls, outputs, labels_val, labels_full)
        dev, st = ws.device, L.stream()
        if out_rows is not None:
            vec, vecsum, valid = out_rows
            if (tuple(vec.shape), tuple(vecsum.shape), tuple(valid.shape)) != ((n, c, d), (n, c), (n, c)) or \
                    not (vec.is_contiguous() and vecsum.is_contiguous() and valid.is_contiguous()) or vec.device != dev:
                raise ValueError(f"out_rows buffers must be contiguous [{n},{c},{d}] / [{n},{c}] / [{n},{c}] tensors on {dev}")
        else:
            vec = torch.empty((n, c, d), dtype=torch.float32, device=dev)
            vecsum = torch.empty((n, c), dtype=torch.float32, device=dev)
            valid = torch.empty((n, c), dtype=torch.uint8, device=dev)
        if n and hw and d:
            L.check(L.lib.diga_centroid_means(sums_p, counts_p, n, c, d, hw, vec.data_ptr(), vecsum.data_ptr(), valid.data_ptr(), st))
        else:
            vec.zero_(), vecsum.zero_(), valid.zero_()
        return vec, vecsum, valid

    def calculate_mean_vector(self, feat_cls, outputs, labels_val=None, model=None):
        """Reference ``calc_centroids.py:120-145``: per image and class the mean feature over the pixels whose
        arg-max prediction (and, if given, label) is that class; classes with < 5 pixels are skipped.
        Returns ``(vectors, ids)`` — ``[D,1,1]`` tensors and Python ints, ordered by (image, class)."""
        vec, _, valid = self._masked_means(feat_cls, outputs, labels_val)
        keep = valid.cpu().numpy().astype(bool)                      # the single host sync of this call
        d = vec.shape[2]
        vectors, ids = [], []
        for n, t in zip(*np.nonzero(keep)):
            vectors.append(vec[n, t].view(d, 1, 1))
            ids.append(int(t))
        return vectors, ids

    def calculate_mean_vector_by_output(self, feat_cls, outputs):
        """Reference ``calc_centroids.py:97-118`` (prediction only)."""
        return self.calculate_mean_vector(feat_cls, outputs, None)

    # -- a7 -------------------------------------------------------------------------------------
    def _mode(self, name):
        if name not in _MODES:
            raise NotImplementedError('no such updating way of objective vectors {}'.format(name))
        return _MODES[name]

    def update_objective_SingleVector(self, id, vector, name='moving_average', start_mean=True):
        """Reference ``calc_centroids.py:147-164``.  ``vector`` may be a torch tensor (any device) or a numpy
        array of D elements.  All conditions (zero vector, ``num < 100``) are evaluated on the GPU."""
        mode = self._mode(name)
        if isinstance(vector, np.ndarray):
            vector = torch.from_numpy(vector)
        v = vector.detach().to(device=self._device, dtype=torch.float32).reshape(-1).contiguous()
        if v.numel() != self.feat_dim:
            raise ValueError(f"vector has {v.numel()} elements, centroids have {self.feat_dim}")
        self._proto_key = None
        with torch.cuda.device(self._device):      # the state's device, whatever device is current or held the vector
            L.check(L.lib.diga_centroid_update_single(v.data_ptr(), int(id), self.class_numbers, self.feat_dim,
                                                      self._objective_vectors.data_ptr(),
                                                      self._objective_vectors_num.data_ptr(), mode, int(bool(start_mean)),
                                                      float(self.centroid_momentum), L.stream()))

    @L.on_device
    def update_from_features(self, feat_cls, outputs, labels_val=None, name='moving_average', start_mean=True, labels_full=None):
        """Fused a6 -> a7 (not in the reference): equals ``calculate_mean_vector`` followed by
        ``update_objective_SingleVector`` on every returned vector in order, with no host sync."""
        mode = self._mode(name)
        self._require_state_device(feat_cls)
        feat, out, lab, full, hh, ww, (n, c, d, h, w), ws = self._prep(feat_cls, outputs, labels_val, labels_full)
        if d != self.feat_dim:
            raise ValueError(f"features have {d} channels, centroids have {self.feat_dim}")
        if n == 0 or h * w == 0:
            return                                        # nothing to accumulate (the reference's loops simply do not run)
        self._proto_key = None
        # ONE scratch tensor and ONE library call queue the whole chain (assign -> accum -> finish): the reference drives this
        # path one image per call (calc_centroids.py:67-78), where per-kernel FFI calls and scratch tensors cost more host
        # time than the kernels take
        L.check(L.lib.diga_centroid_chain(feat.data_ptr(), out.data_ptr(), L.ptr(lab), L.ptr(full), hh, ww, n, c, d, h, w,
                                          ws.data_ptr(), self._objective_vectors.data_ptr(),
                                          self._objective_vectors_num.data_ptr(), mode, int(bool(start_mean)),
                                          float(self.centroid_momentum), L.stream()))

    def _update_sharded(self, gvec, gsum, gvalid, n_total, batch, world, per_shard, name='mean', start_mean=True):
        """Ordered replay of an all-gathered row buffer (``diga_b200.parallel``): the reference recurrence over the
        ``n_total`` images of the global loader sequence, rank r holding batches r, r + world, ... of ``batch`` images."""
        mode = self._mode(name)
        L.require_cuda(gvec, gsum, gvalid, self._objective_vectors, what="sharded centroid update")
        c, d = self.class_numbers, self.feat_dim
        rows = world * per_shard
        if tuple(gvec.shape) != (rows, c, d) or tuple(gsum.shape) != (rows, c) or tuple(gvalid.shape) != (rows, c):
            raise ValueError(f"gathered buffers must be [{rows},{c},{d}] / [{rows},{c}] / [{rows},{c}]")
        self._proto_key = None
        with torch.cuda.device(self._device):
            L.check(L.lib.diga_centroid_update_sharded(gvec.data_ptr(), gsum.data_ptr(), gvalid.data_ptr(), int(n_total), int(batch),
                                                       int(world), int(per_shard), c, d, self._objective_vectors.data_ptr(),
                                                       self._objective_vectors_num.data_ptr(), mode, int(bool(start_mean)),
                                                       float(self.centroid_momentum), L.stream()))

    def update_from_features_sharded(self, feat_cls, outputs, labels_val=None, name='moving_average', start_mean=True,
                                     labels_full=None, group=None):
        """Multi-rank form of :meth:`update_from_features` for the ONLINE update of the self-training step
        (self_training.py:327-341): every rank computes the class vectors of its own batch, one all-gather of
        ``[B, C, D]`` per rank (1.2 MB at B=8, D=2048) exchanges them, and every rank replays the updates in the order
        rank 0's images, rank 1's images, ... — all ranks keep IDENTICAL centroids, equal to what a single process reaches on
        the concatenated global batch.  Without an initialised process group this is ``update_from_features``."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.update_from_features(feat_cls, outputs, labels_val, name, start_mean, labels_full)
        self._mode(name)
        self._require_state_device(feat_cls)
        world = dist.get_world_size(group)
        vec, vecsum, valid = self._masked_means(feat_cls, outputs, labels_val, labels_full)
        n = vec.shape[0]
        gathered = []
        for t in (vec, vecsum, valid):
            g = torch.empty((world * n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(g, t, group=group)
            gathered.append(g)
        self._update_sharded(*gathered, world * n, n, world, n, name, start_mean)

    @L.on_device
    def accumulate_mean_pass(self, acc, feat_cls, outputs, labels_val=None):
        """Multi-GPU 'mean' pass, sum mode: ``acc [C, D+1]`` += (sum of per-image class means, image count).
        ``diga_b200.parallel.finish_mean_pass`` all-reduces ``acc`` and writes the centroids.  One library call per batch."""
        feat, out, lab, full, hh, ww, (n, c, d, h, w), ws = self._prep(feat_cls, outputs, labels_val, None)
        if tuple(acc.shape) != (c, d + 1) or acc.dtype != torch.float32 or not acc.is_cuda or acc.device != feat.device:
            raise ValueError(f"acc must be a CUDA fp32 [{c},{d + 1}] tensor on {feat.device}")
        L.check(L.lib.diga_centroid_chain_reduce(feat.data_ptr(), out.data_ptr(), L.ptr(lab), None, 0, 0, n, c, d, h, w, ws.data_ptr(),
                                                 acc.data_ptr(), L.stream()))

    # -- a5 -------------------------------------------------------------------------------------
    @L.on_device
    def _proto(self, feat, want_dist, want_weight):
        L.require_cuda(feat, what="Class_Features input")
        f = L.f32c(feat.detach())
        n, d, h, w = f.shape
        c = self.class_numbers
        if d != self.feat_dim:
            raise ValueError(f"features have {d} channels, centroids have {self.feat_dim}")
        cen = self._objective_vectors
        if cen.device != f.device:
            cen = cen.to(f.device)
        need = int(L.lib.diga_proto_workspace_bytes(c, d))
        if self._proto_ws is None or self._proto_ws.numel() < need or self._proto_ws.device != f.device:
            self._proto_ws = torch.empty(need, dtype=torch.uint8, device=f.device)
            self._proto_key = None
        dist = torch.empty((n, c, h, w), dtype=torch.float32, device=f.device) if want_dist else None
        weight = torch.empty((n, c, h, w), dtype=torch.float32, device=f.device) if want_weight else None
        # the split centroid operands are rebuilt only when the centroids changed (assignment, in-place torch op, or one
        # of our own update kernels, which reset the key)
        key = (cen.data_ptr(), cen._version, f.device, L.stream())
        if key != self._proto_key:
            L.check(L.lib.diga_proto_prepare(cen.data_ptr(), c, d, self._proto_ws.data_ptr(), L.stream()))
            self._proto_key = key
        L.check(L.lib.diga_proto_distance_prepared(f.data_ptr(), cen.data_ptr(), n, d, c, h * w, L.ptr(dist), L.ptr(weight),
                                                   self._proto_ws.data_ptr(), L.stream()))
        return dist, weight

    def feat_centroid_distance(self, feat):
        """``[N,C,H,W]`` L2 distance of every feature pixel to every centroid; reference :166-171."""
        return self._proto(feat, True, False)[0]

    def get_centroid_weight(self, feat):
        """``softmax(-distance)`` over classes; reference :173-176."""
        return self._proto(feat, False, True)[1]

    def get_centroid_distance(self, feat):
        """Negated distance; reference :178-180."""
        return self._proto(feat, True, False)[0].neg_()


def _labels_on_feature_grid(labels_i64, size):
    """self_training.py:329-330 / calc_centroids.py:60-62: [B,H,W] int64 -> [B,1,h,w] fp32, nearest."""
    b, h, w = labels_i64.size()
    return F.interpolate(labels_i64.reshape([b, 1, h, w]).float(), size=tuple(size), mode='nearest')


def _sharded_target_pass(class_features, model, target_loader, rank, world, first, preprocess=None, state=None):
    """One pass of the target loop (calc_centroids.py:67-78) under ``torchrun``: every rank runs the backbone only on the
    loader batches ``rank, rank + world, ...`` (all ranks iterate the SAME un-sharded loader, like the reference's), the
    per-image class vectors are all-gathered once and replayed in loader order (exact mode of ``diga_b200.parallel``): all
    ranks end the pass with the centroids the single-process loop produces."""
    from .parallel import ShardedCentroidPass
    state = {} if state is None else state
    sp = state.get("pass")                       # the row buffers (a symmetric allocation) are set up once and re-used by every pass
    n_images = len(target_loader.dataset) if hasattr(target_loader, "dataset") else state.get("n_images")
    bsz = getattr(target_loader, "batch_size", None) or state.get("bsz")
    pending = []
    for index, batch in enumerate(target_loader):
        if index % world != rank:
            continue
        tdatav = batch[0].cuda()
        if preprocess is not None:
            tdatav = preprocess(tdatav)
        with torch.no_grad():
            _, _, out, feature_t = model(tdatav)
            if first:
                class_features.objective_vectors = torch.zeros([19, feature_t.shape[1]])
                first = False
            if sp is None and n_images is not None and bsz:
                sp = state["pass"] = ShardedCentroidPass(class_features, n_images, bsz, 'mean')
            if sp is not None:
                sp.add(feature_t, out)
            else:
                pending.append(class_features._masked_means(feature_t, out, None))
    if sp is None:
        # a loader without len()/batch_size: sizes are only known now; exchange them, then fill the pass buffer
        sizes = torch.tensor([sum(p[0].shape[0] for p in pending), max([p[0].shape[0] for p in pending] + [0])], device="cuda")
        tot = sizes.clone()
        torch.distributed.all_reduce(tot[:1])
        torch.distributed.all_reduce(sizes[1:], op=torch.distributed.ReduceOp.MAX)
        if first and int(tot[0]) > 0:
            raise RuntimeError("calc_centroids: this rank received no batch, cannot infer the feature dimension")
        state["n_images"], state["bsz"] = int(tot[0]), max(int(sizes[1]), 1)      # known from now on: later passes use add()
        sp = ShardedCentroidPass(class_features, state["n_images"], state["bsz"], 'mean', symmetric=False)
        for p in pending:
            sp.add_rows(*p)
    sp.finish()


def calc_centroids(opt, model, enc_s, dec_s2t, source_loader, source_loader_full, target_loader, preprocess=None):
    """Reference driver ``calc_centroids.py:17-81``: five passes over the loader, running 'mean' update of the
    class centroids, ``feat_centroids`` written next to ``opt.centroid_dir`` after every pass.

    ``model(x)`` must return ``(_, _, logits, feat)`` like the reference ``SegModel``.  As in the reference the
    target branch is always taken (``opt.source`` is overwritten with ``False``, :27); the source branch is kept
    for completeness (:29-65).  Returns the ``Class_Features`` object (the reference returns ``None``).

    ``preprocess`` (not in the reference signature, default ``None``): applied to the image batch before ``model`` — the
    semi-supervised tree flips BGR -> RGB at the call site (``model(sdatav[:, [2, 1, 0], :, :])``,
    semi-supervised_segmentation/calc_centroids.py:39); the GTA5 tree mirrored here feeds the image as is.

    Under ``torchrun`` (an initialised ``torch.distributed`` group of more than one rank) the target loop is image-sharded
    in the exact mode of ``diga_b200.parallel``: identical ``feat_centroids`` on every rank, bit-equal to the single-process run.
    """
    class_features = Class_Features(numbers=19)
    first = True
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    sharded_state = {}
    for epoch in range(5):
        model.eval(), enc_s.eval(), dec_s2t.eval()
        opt.source = False
        if opt.source:
            for index, (batch, batch_full) in enumerate(zip(source_loader, source_loader_full)):
                sdatav = torch.cat([batch[0].cuda(), batch_full[0].cuda()], dim=0)
                slabelv = torch.cat([batch[1].cuda(), batch_full[1].cuda()], dim=0)
                with torch.no_grad():
                    rec_s2t = dec_s2t(enc_s(sdatav))
                    _, transmix = classmix(slabelv, rec_s2t, sdatav, rng=random)
                    _, _, out, feature_s = model(transmix)
                    if first:
                        class_features.objective_vectors = torch.zeros([19, feature_s.shape[1]])
                        first = False
                    newlabels = _labels_on_feature_grid(slabelv, out.size()[2:])
                    class_features.update_from_features(feature_s, out, newlabels, 'mean')
        elif world > 1:
            _sharded_target_pass(class_features, model, target_loader, rank, world, first, preprocess, sharded_state)
            first = False
        else:
            for index, batch in enumerate(target_loader):
                if index % 100 == 0:
                    print('epoch', epoch)
                    print('%d processd' % index)
                tdatav = batch[0].cuda()
                if preprocess is not None:
                    tdatav = preprocess(tdatav)
                with torch.no_grad():
                    _, _, out, feature_t = model(tdatav)
                    if first:
                        class_features.objective_vectors = torch.zeros([19, feature_t.shape[1]])
                        first = False
                    class_features.update_from_features(feature_t, out, None, 'mean')
        save_path = os.path.join(os.path.dirname(opt.centroid_dir), "feat_centroids")
        if rank == 0:                                       # every rank holds the same tensor; one writer
            torch.save(class_features.objective_vectors.cpu(), save_path)
    return class_features
