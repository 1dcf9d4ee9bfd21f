"""GPU parity of the multi-GPU centroid path and of the BASELINE config shapes round 1 left untested.

* sum mode  : ``accumulate_mean_pass`` (``diga_centroid_reduce_images``) + ``finish_mean_pass`` vs the sequential reference
  loop (calc_centroids.py:67-78, :147-164) restated by ``oracle.centroid_pass``;
* exact mode: ``ShardedCentroidPass`` / ``diga_centroid_update_sharded`` — the all-gathered row buffer of W "virtual ranks"
  built on one GPU must replay to the bits of the sequential single-process update, beyond the 3000 clamp and over several
  passes; the real 2-rank NCCL run of both modes is spawned through ``torch.distributed.run`` when two GPUs are visible;
* config shapes: a5 at [1,2048,129,257] / [1,256,129,257], a6 at [1,2048,65,129], u8 consensus selection at
  129x257 -> 1024x2048 against the GPU-eager oracle, KD at the reference's own [6,19,512,896].
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import diga_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5


def dev():
    return torch.device("cuda", 0)


def assert_normwise(got, want, rtol=RTOL, what=""):
    got, want = got.detach().double().cpu(), torch.as_tensor(want).detach().double().cpu()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= rtol * max(scale, 1e-30), f"{what}: max|diff| {err:.3e} > {rtol} * max|ref| {scale:.3e}"


def assert_rel(got, want, rtol=RTOL, what=""):
    got, want = float(got), float(want)
    assert abs(got - want) <= rtol * abs(want), f"{what}: {got!r} vs {want!r}"


@pytest.fixture(scope="module")
def D():
    import diga_b200
    return diga_b200


def _inputs(n, d, h, w, c, g, labels=False, boost=2.0):
    from diga_b200 import synthetic as S
    feat = S.features((n, d, h, w), g)
    out = S.logits((n, c, h, w), g)
    out[:, : max(2, c // 3)] += boost
    lab = None
    if labels:
        lab = out.argmax(1, keepdim=True).float()
        flip = torch.rand((n, 1, h, w), generator=g, device=g.device) < 0.3
        lab[flip] = 255.0
    return feat, out, lab


def _oracle_pass(batches, c, d, cf=None):
    """calc_centroids.py:67-78 over in-memory batches with optional labels (oracle evaluated where the tensors live)."""
    cf = cf or O.ClassFeaturesOracle(c, d)
    for feat, out, lab in batches:
        vectors, ids = cf.calculate_mean_vector(feat, out, lab)
        for v, i in zip(vectors, ids):
            cf.update_objective_SingleVector(i, v.detach().cpu().numpy(), "mean")
    return cf


# ------------------------------------------------------------------------------------------------ sum mode (one all-reduce)
@pytest.mark.parametrize("n,d,h,w,labels", [(8, 2048, 65, 129, False), (1, 2048, 65, 129, False), (3, 256, 65, 113, True),
                                            (2, 64, 9, 11, True)])
def test_mean_pass_sum_mode_vs_sequential_oracle(D, n, d, h, w, labels):
    """accumulate_mean_pass over several batches + finish_mean_pass == the reference's sequential 'mean' updates."""
    from diga_b200 import parallel as P, synthetic as S
    g = S.gen(99, "cuda")
    c = 19
    batches = [_inputs(n, d, h, w, c, g, labels) for _ in range(3)]
    cf = D.Class_Features(c, d)
    acc = P.new_mean_accumulator(c, d, dev())
    for feat, out, lab in batches:
        cf.accumulate_mean_pass(acc, feat, out, lab)
    vectors, num = P.finish_mean_pass(acc)
    ref = _oracle_pass(batches, c, d)
    assert torch.equal(num.cpu(), ref.objective_vectors_num)
    for k in range(c):
        assert_normwise(vectors[k], ref.objective_vectors[k], what=f"centroid {k}")
    # with the Class_Features handle the result is written back and a second pass continues from it (num + n <= 3000: exact)
    cf2 = D.Class_Features(c, d)
    for half in (batches[:1], batches[1:]):
        acc.zero_()
        for feat, out, lab in half:
            cf2.accumulate_mean_pass(acc, feat, out, lab)
        P.finish_mean_pass(acc, cf2)
    assert torch.equal(cf2.objective_vectors_num.cpu(), ref.objective_vectors_num)
    for k in range(c):
        assert_normwise(cf2.objective_vectors[k], ref.objective_vectors[k], what=f"continued centroid {k}")


def test_mean_pass_all_invalid_image_and_zero_sum_vector(D):
    """An image whose labels gate every pixel out contributes nothing; a class whose mean vector sums to exactly zero is
    skipped (calc_centroids.py:148) and does not advance the count."""
    from diga_b200 import parallel as P, synthetic as S
    g = S.gen(5, "cuda")
    n, d, h, w, c = 3, 64, 17, 19, 19
    feat, out, lab = _inputs(n, d, h, w, c, g, labels=True)
    lab[1] = 255.0                                             # image 1: nothing survives the label gate
    am = out.argmax(1)
    feat[0].masked_fill_((am[0] == 2).unsqueeze(0), 0.0)       # image 0, class 2: every selected feature is 0 -> vector sum 0
    assert int(((am[0] == 2) & (lab[0, 0] == 2)).sum()) >= 5
    cf = D.Class_Features(c, d)
    acc = P.new_mean_accumulator(c, d, dev())
    cf.accumulate_mean_pass(acc, feat, out, lab)
    vectors, num = P.finish_mean_pass(acc)
    ref = _oracle_pass([(feat, out, lab)], c, d)
    assert torch.equal(num.cpu(), ref.objective_vectors_num)
    for k in range(c):
        assert_normwise(vectors[k], ref.objective_vectors[k], what=f"centroid {k}")
    _, ids = cf.calculate_mean_vector(feat[1:2], out[1:2], lab[1:2])
    assert ids == []


# ------------------------------------------------------------------------------------------------ exact mode
def _sequential(D, batches, c, d, passes, name="mean", start_mean=True, start=None):
    cf = D.Class_Features(c, d)
    if start is not None:
        cf.objective_vectors, cf.objective_vectors_num = start[0].clone(), start[1].clone()
    for _ in range(passes):
        for feat, out, lab in batches:
            cf.update_from_features(feat, out, lab, name, start_mean)
    return cf


@pytest.mark.parametrize("world,batch,n_images", [(1, 1, 37), (2, 1, 37), (8, 1, 37), (4, 3, 37), (8, 8, 30), (3, 2, 5)])
def test_sharded_replay_equals_sequential_bitwise(D, world, batch, n_images):
    """W virtual ranks fill their own ShardedCentroidPass buffers (batches[r::W]); the concatenation of the buffers is what
    the all-gather delivers; diga_centroid_update_sharded over it must equal the sequential update BIT FOR BIT."""
    from diga_b200 import parallel as P, synthetic as S
    g = S.gen(2024, "cuda")
    c, d, h, w = 19, 96, 9, 13
    nb = -(-n_images // batch)
    batches = [_inputs(min(batch, n_images - k * batch), d, h, w, c, g, labels=(k % 2 == 0)) for k in range(nb)]
    seq = _sequential(D, batches, c, d, passes=2)
    cf = D.Class_Features(c, d)
    shards = []
    for r in range(world):
        # a virtual rank: the same buffers a real rank r of `world` would fill
        shards.append(P.ShardedCentroidPass(cf, n_images, batch, rank=r, world=world))
    for _ in range(2):
        for sp in shards:
            assert list(sp.my_batches()) == P.shard_indices(nb, sp.rank, world)
            for k in sp.my_batches():
                sp.add(*batches[k][:2], batches[k][2])
        per = shards[0].per_shard
        gathered = [torch.cat([getattr(sp, name) for sp in shards]) for name in ("vec", "vecsum", "valid")]
        cf._update_sharded(*gathered, n_images, batch, world, max(per, 1), "mean", True)
        for sp in shards:
            sp.reset()
    assert torch.equal(cf.objective_vectors_num, seq.objective_vectors_num)
    assert torch.equal(cf.objective_vectors, seq.objective_vectors), "sharded replay differs from the sequential update"
    # the host mirror of the device row order
    rows = P.global_row_order(n_images, batch, world, max(per, 1))
    assert len(set(rows)) == n_images and max(rows) < world * max(per, 1)


def test_exact_mode_beyond_clamp_five_passes_vs_oracle(D):
    """The reference runs 5 passes (calc_centroids.py:20-23) and clamps the count at 3000 (:156,:161): after the clamp the
    'mean' update is an order-dependent recursion.  Exact mode must track the sequential oracle through it; sum mode is the
    documented approximation (its error against the exact result is reported, and must be small but non-zero)."""
    from diga_b200 import parallel as P, synthetic as S
    g = S.gen(7, "cuda")
    c, d, h, w, n_img = 19, 32, 6, 8, 640
    feats = S.features((n_img, d, h, w), g)
    outs = S.logits((n_img, c, h, w), g)
    outs[:, :3] += 4.0                                                   # classes 0..2 are valid in (nearly) every image
    cf = D.Class_Features(c, d)
    sp = P.ShardedCentroidPass(cf, n_img, batch=8)
    cf_sum = D.Class_Features(c, d)
    acc = P.new_mean_accumulator(c, d, dev())
    ref = O.ClassFeaturesOracle(c, d)
    for _ in range(5):
        acc.zero_()
        for k in sp.my_batches():
            sl = slice(8 * k, 8 * k + 8)
            sp.add(feats[sl], outs[sl])
            cf_sum.accumulate_mean_pass(acc, feats[sl], outs[sl])
        sp.finish()
        P.finish_mean_pass(acc, cf_sum)
        ref = O.centroid_pass([feats], [outs], c, d, cf=ref)             # one call = all images in order
    assert ref.objective_vectors_num.max().item() == 3000.0              # the clamp was reached (5 x 640 > 3000)
    assert torch.equal(cf.objective_vectors_num.cpu(), ref.objective_vectors_num)
    for k in range(c):
        assert_normwise(cf.objective_vectors[k], ref.objective_vectors[k], what=f"exact-mode centroid {k} after 5 passes")
    # and bit-equal to the sequential device update
    seq = _sequential(D, [(feats[8 * k:8 * k + 8], outs[8 * k:8 * k + 8], None) for k in range(n_img // 8)], c, d, passes=5)
    assert torch.equal(cf.objective_vectors, seq.objective_vectors) and torch.equal(cf.objective_vectors_num, seq.objective_vectors_num)
    # sum mode: same counts, vectors close (uniform instead of recency weights beyond the clamp)
    assert torch.equal(cf_sum.objective_vectors_num, cf.objective_vectors_num)
    err = (cf_sum.objective_vectors - cf.objective_vectors).abs().max().item() / cf.objective_vectors.abs().max().item()
    assert err < 5e-2, f"sum-mode approximation error {err:.3e}"


def test_online_sharded_update_without_process_group_is_the_plain_update(D):
    from diga_b200 import synthetic as S
    g = S.gen(11, "cuda")
    feat, out, lab = _inputs(2, 64, 9, 11, 19, g, labels=True)
    a, b = D.Class_Features(19, 64), D.Class_Features(19, 64)
    a.update_from_features(feat, out, lab, "moving_average", True)
    b.update_from_features_sharded(feat, out, lab, "moving_average", True)
    assert torch.equal(a.objective_vectors, b.objective_vectors) and torch.equal(a.objective_vectors_num, b.objective_vectors_num)


def test_long_update_kernel_matches_short_kernel(D):
    """n > 32 rows per call takes the list-compaction kernel; on the SAME per-image vectors it must agree bit for bit with
    one-row calls of the short kernel (EMA mode with start_mean: the rule switches from 'mean' to EMA at num == 100 inside
    the sequence)."""
    from diga_b200 import _lib as L, synthetic as S
    g = S.gen(3, "cuda")
    n, d, h, w, c = 150, 48, 9, 11, 19
    feat, out = S.features((n, d, h, w), g), S.logits((n, c, h, w), g)
    out[:, :3] += 6.0                                   # three classes own (nearly) every image: their count passes 100
    one, many = D.Class_Features(c, d), D.Class_Features(c, d)
    vec, vecsum, valid = many._masked_means(feat, out, None)

    def update(cf, rows):
        L.check(L.lib.diga_centroid_update(vec[rows].data_ptr(), vecsum[rows].data_ptr(), valid[rows].data_ptr(), rows.stop - rows.start,
                                           c, d, cf.objective_vectors.data_ptr(), cf.objective_vectors_num.data_ptr(),
                                           L.UPDATE_MOVING_AVERAGE, 1, 1e-4, L.stream()))

    for i in range(n):
        update(one, slice(i, i + 1))
    update(many, slice(0, n))
    assert one.objective_vectors_num.max().item() > 100
    assert torch.equal(one.objective_vectors_num, many.objective_vectors_num)
    assert torch.equal(one.objective_vectors, many.objective_vectors)


@pytest.mark.parametrize("start", [0.0, 2990.0, 17.5])
def test_long_update_division_is_ieee_for_every_count_and_magnitude(D, start):
    """The whole-pass kernel against the short kernel on the same rows: 3100 rows in 'mean' mode walk every divisor 1..3001
    (then the clamp), on vectors from 1e-30 to 1e30 with zeros, negative zeros and infinities mixed in, from integer and
    non-integer start counts.  Bit-equal centroids and counts.  (Written for an experiment that took the reciprocal of
    `num + 1` off the per-channel chain — exact, but not faster, DESIGN.md §8 — and kept as a guard of the replay.)"""
    from diga_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(41)
    n, c, d = 3100, 3, 320
    mag = torch.exp(torch.empty((n, c, d), device="cuda").uniform_(-69.0, 69.0, generator=g))
    vec = mag * torch.where(torch.rand((n, c, d), device="cuda", generator=g) < 0.5, -1.0, 1.0)
    r = torch.rand((n, c, d), device="cuda", generator=g)
    vec[r < 0.02] = 0.0
    vec[(r >= 0.02) & (r < 0.03)] = -0.0
    vec[:, 2, :7] = float("inf")                                    # one class sees infinities in a few channels
    vec[:, 1] = torch.randn((n, d), device="cuda", generator=g)      # and one class ordinary values
    vecsum = torch.ones((n, c), device="cuda")
    vecsum[torch.rand((n, c), device="cuda", generator=g) < 0.05] = 0.0          # skipped rows
    valid = torch.ones((n, c), dtype=torch.uint8, device="cuda")

    def run(chunks):
        obj = torch.zeros((c, d), device="cuda")
        num = torch.full((c,), start, device="cuda")
        for lo, hi in chunks:
            L.check(L.lib.diga_centroid_update(vec[lo:hi].data_ptr(), vecsum[lo:hi].data_ptr(), valid[lo:hi].data_ptr(), hi - lo, c, d,
                                               obj.data_ptr(), num.data_ptr(), L.UPDATE_MEAN, 0, 1e-4, L.stream()))
        return obj, num

    short = run([(i, min(i + 32, n)) for i in range(0, n, 32)])      # <= 32 rows per call: the short kernel
    long_ = run([(0, n)])
    assert torch.equal(short[1], long_[1])
    same = (short[0] == long_[0]) | (torch.isnan(short[0]) & torch.isnan(long_[0]))
    assert bool(same.all()), f"{int((~same).sum())} of {same.numel()} centroid entries differ"
    assert torch.equal(short[0].view(torch.int32)[~torch.isnan(short[0])], long_[0].view(torch.int32)[~torch.isnan(long_[0])])


NCCL_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import diga_b200 as D
from diga_b200 import parallel as P, synthetic as S
from oracle import diga_oracle as O
c, d, h, w, n_img, batch = 19, 256, 33, 65, 46, 4
g = S.gen(77, dev)                                   # same seed on every rank: every rank can see the whole synthetic set
nb = -(-n_img // batch)
feats = [S.features((min(batch, n_img - k * batch), d, h, w), g) for k in range(nb)]
outs = [S.logits((f.shape[0], c, h, w), g) for f in feats]
for o in outs: o[:, :6] += 3
# single-rank sequential oracle over the union (the reference loop, calc_centroids.py:67-78, :157-161)
ref = O.ClassFeaturesOracle(c, d)
for _ in range(2):
    ref = O.centroid_pass(feats, outs, c, d, cf=ref)
# exact mode, two passes: rows exchanged by peer stores fused into the means kernel (symmetric memory) ...
cf = D.Class_Features(c, d)
sp = P.ShardedCentroidPass(cf, n_img, batch)
print("rank", rank, "exchange:", sp.exchange, getattr(sp, "_symm_error", ""))
assert sp.exchange == os.environ.get("DIGA_EXPECT_EXCHANGE", sp.exchange)
for _ in range(2):
    for k in sp.my_batches():
        sp.add(feats[k], outs[k])
    sp.finish()
# ... and by the all-gather: the same bits
cf_ag = D.Class_Features(c, d)
sp_ag = P.ShardedCentroidPass(cf_ag, n_img, batch, symmetric=False)
assert sp_ag.exchange == "all-gather"
for _ in range(2):
    for k in sp_ag.my_batches():
        sp_ag.add(feats[k], outs[k])
    sp_ag.finish()
assert torch.equal(cf_ag.objective_vectors, cf.objective_vectors) and torch.equal(cf_ag.objective_vectors_num, cf.objective_vectors_num)
num = cf.objective_vectors_num.cpu()
assert torch.equal(num, ref.objective_vectors_num), (num, ref.objective_vectors_num)
err = (cf.objective_vectors.cpu() - ref.objective_vectors).abs().max().item()
scale = ref.objective_vectors.abs().max().item()
assert err <= 1e-5 * scale, ("exact", err, scale)
# ... and bit-equal to the sequential device path, and identical on all ranks
seq = D.Class_Features(c, d)
for _ in range(2):
    for f, o in zip(feats, outs):
        seq.update_from_features(f, o, None, "mean")
assert torch.equal(seq.objective_vectors, cf.objective_vectors)
other = [torch.empty_like(cf.objective_vectors) for _ in range(world)]
dist.all_gather(other, cf.objective_vectors)
assert all(torch.equal(o, other[0]) for o in other)
# sum mode, first pass from empty centroids: exact below the clamp
cf2 = D.Class_Features(c, d)
acc = P.new_mean_accumulator(c, d, dev)
for k in P.shard_indices(nb, rank, world):
    cf2.accumulate_mean_pass(acc, feats[k], outs[k])
vectors, num2 = P.finish_mean_pass(acc, cf2)
ref1 = O.centroid_pass(feats, outs, c, d)
assert torch.equal(num2.cpu(), ref1.objective_vectors_num)
err = (vectors.cpu() - ref1.objective_vectors).abs().max().item()
assert err <= 1e-5 * ref1.objective_vectors.abs().max().item(), ("sum", err)
# online EMA, sharded: every rank's own batch, replayed rank-major on all ranks == one process over the concatenated batch
gl = S.gen(500 + rank, dev)
f_r, o_r = S.features((2, d, h, w), gl), S.logits((2, c, h, w), gl)
cf3 = D.Class_Features(c, d)
cf3.objective_vectors, cf3.objective_vectors_num = ref.objective_vectors.clone(), torch.full((c,), 150.0)
cf3.update_from_features_sharded(f_r, o_r, None, "moving_average", False)
fs, os_ = [torch.empty_like(f_r) for _ in range(world)], [torch.empty_like(o_r) for _ in range(world)]
dist.all_gather(fs, f_r); dist.all_gather(os_, o_r)
one = D.Class_Features(c, d)
one.objective_vectors, one.objective_vectors_num = ref.objective_vectors.clone(), torch.full((c,), 150.0)
one.update_from_features(torch.cat(fs), torch.cat(os_), None, "moving_average", False)
assert torch.equal(one.objective_vectors, cf3.objective_vectors) and torch.equal(one.objective_vectors_num, cf3.objective_vectors_num)
# the calc_centroids driver itself under torchrun (calc_centroids.py:17-81): every rank iterates the same loader, runs the
# model on its own batches only, and all ranks end every pass with the single-process centroids; rank 0 writes the file
import types, torch.nn as nn
class TinySeg(nn.Module):
    def __init__(self, d=64, c=19):
        super().__init__()
        torch.manual_seed(0)
        self.f = nn.Conv2d(3, d, 8, stride=8)
        self.g = nn.Conv2d(d, c, 1)
    def forward(self, x):
        feat = self.f(x)
        return None, None, 4.0 * self.g(feat), feat
model = TinySeg().to(dev)
gen = torch.Generator().manual_seed(1)
imgs = torch.randn((7, 3, 64, 96), generator=gen)
for loader in (torch.utils.data.DataLoader(torch.utils.data.TensorDataset(imgs, torch.zeros(7)), batch_size=2, shuffle=False),
               [(imgs[i:i + 2], torch.zeros(1)) for i in range(0, 7, 2)]):          # DataLoader (len / batch_size known) and a bare list
    opt = types.SimpleNamespace(source=True, centroid_dir=os.path.join(sys.argv[2], "centroids", "x"))
    os.makedirs(os.path.dirname(opt.centroid_dir), exist_ok=True)
    cfd = D.calc_centroids(opt, model, nn.Identity(), nn.Identity(), [], [], loader)
    seq2 = D.Class_Features(19, 64)
    with torch.no_grad():
        for _ in range(5):
            for batch in loader:
                _, _, out_b, feat_b = model(batch[0].to(dev))
                seq2.update_from_features(feat_b, out_b, None, "mean")
    assert torch.equal(cfd.objective_vectors_num, seq2.objective_vectors_num), (cfd.objective_vectors_num, seq2.objective_vectors_num)
    err = (cfd.objective_vectors - seq2.objective_vectors).abs().max().item()
    assert err <= 1e-6 * seq2.objective_vectors.abs().max().item(), ("driver", err)
    dist.barrier()
    if rank == 0:
        saved = torch.load(os.path.join(os.path.dirname(opt.centroid_dir), "feat_centroids"))
        assert torch.equal(saved, cfd.objective_vectors.cpu())
    dist.barrier()
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("world", [2])
def test_two_rank_nccl_exact_and_sum_modes(tmp_path, world):
    """Real NCCL run: `world` ranks over disjoint shards == the single-rank sequential oracle (counts equal, vectors 1e-5;
    exact mode bit-equal to the sequential device path).  Needs `world` visible GPUs."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "nccl_worker.py"
    script.write_text(NCCL_WORKER)
    port = 29700 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT, str(tmp_path)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:]
    assert res.stdout.count("ok") >= world


def test_device_guard_non_current_device(D):
    """Tensors on cuda:1 while cuda:0 is current: the wrappers switch to the tensors' device for the call."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from diga_b200 import synthetic as S
    d1 = torch.device("cuda", 1)
    g = S.gen(1, d1)
    assert torch.cuda.current_device() == 0
    t, s = S.logits((2, 19, 16, 24), g), S.logits((2, 19, 16, 24), g)
    so = s.clone().requires_grad_(True)
    lo = O.distillation_loss(t, so)
    lo.backward()
    sg = s.clone().requires_grad_(True)
    lg = D.distillation_loss(t, sg)
    lg.backward()
    assert lg.device == d1 and torch.cuda.current_device() == 0
    assert_rel(lg.item(), lo.item())
    assert_normwise(sg.grad, so.grad)
    cf = D.Class_Features(19, 64, device=d1)
    feat, out, _ = _inputs(2, 64, 9, 11, 19, g)
    cf.update_from_features(feat, out, None, "mean")
    ref = _oracle_pass([(feat, out, None)], 19, 64)
    assert torch.equal(cf.objective_vectors_num.cpu(), ref.objective_vectors_num)
    with pytest.raises(RuntimeError):
        D.distillation_loss(t, s.to("cuda:0"))
    with pytest.raises(RuntimeError):
        D.Class_Features(19, 64).update_from_features(feat, out, None, "mean")      # state on cuda:0, features on cuda:1


# ------------------------------------------------------------------------------------------------ BASELINE config shapes
def test_kd_reference_shape_6x19x512x896(D):
    """The reference's own training shape (B=3 per view at 512x896); oracle = its op chain on the GPU."""
    from diga_b200 import synthetic as S
    g = S.gen(6, "cuda")
    t, s = S.logits((6, 19, 512, 896), g), S.logits((6, 19, 512, 896), g)
    so = s.clone().requires_grad_(True)
    lo = O.distillation_loss(t, so, 0.5)
    (lo * 0.25).backward()
    sg = s.clone().requires_grad_(True)
    lg = D.distillation_loss(t, sg, 0.5)
    (lg * 0.25).backward()
    assert_rel(lg.item(), lo.item(), what="loss")
    assert_normwise(sg.grad, so.grad, what="grad")
    assert_normwise(sg.grad, O.distillation_grad_closed_form(t, s, 0.5, 0.25), what="grad vs fp64 closed form")


@pytest.mark.parametrize("n,d,h,w", [(1, 2048, 65, 129), (8, 2048, 65, 129)])
def test_centroid_pass_config4_shape(D, n, d, h, w):
    """a6 + a7 at config 4's shape; the oracle runs on the GPU tensors (same op chain the reference would run there)."""
    from diga_b200 import synthetic as S
    g = S.gen(44, "cuda")
    c = 19
    batches = [_inputs(n, d, h, w, c, g) for _ in range(2)]
    gcf = D.Class_Features(c, d)
    ref = O.ClassFeaturesOracle(c, d)
    for feat, out, _ in batches:
        vec, ids = ref.calculate_mean_vector(feat, out)
        gvec, gids = gcf.calculate_mean_vector(feat, out)
        assert gids == ids
        for a, b in zip(gvec, vec):
            assert_normwise(a.reshape(-1), b.reshape(-1), what="mean vector")
        for v, i in zip(vec, ids):
            ref.update_objective_SingleVector(i, v.detach().cpu().numpy(), "mean")
        gcf.update_from_features(feat, out, None, "mean")
    assert torch.equal(gcf.objective_vectors_num.cpu(), ref.objective_vectors_num)
    for k in range(c):
        assert_normwise(gcf.objective_vectors[k], ref.objective_vectors[k], what=f"centroid {k}")


def test_consensus_select_uint8_vs_gpu_eager_oracle(D):
    """The uint8 kernel (config 5 and the PNG stage) against the reference op chain on the GPU — not against the int64
    kernel — at 129x257 -> 1024x2048 and at an odd geometry."""
    from diga_b200 import synthetic as S
    g = S.gen(31, "cuda")
    for b, c, lo, hi in ((1, 19, (129, 257), (1024, 2048)), (2, 19, (65, 129), (512, 1024)), (1, 16, (9, 13), (37, 53))):
        wl = torch.softmax(S.logits((b, c, *lo), g), 1)
        pl = S.block_labels(b, hi[0], hi[1], g, 16, c)
        kept_o, fp_o = O.consensus_select(pl, wl, hi)
        k8, f8 = D.consensus_select(pl.to(torch.uint8), wl)
        assert k8.dtype == torch.uint8 and f8.dtype == torch.uint8
        assert int((f8.long() != fp_o).sum()) == 0, "u8 feat_pseudo differs from the GPU eager reference"
        assert torch.equal(k8.long(), kept_o)


def test_ema_more_than_512_tensors_with_empty_ones(D):
    """> 512 tensors with empty ones in between: every tensor is updated exactly once."""
    from diga_b200.util.utils import ema_update_tensors
    g = torch.Generator().manual_seed(5)
    sizes = ([0, 5, 0, 130, 7] * 130)[:640]
    ts = [torch.randn(n, generator=g) for n in sizes]
    ss = [torch.randn(n, generator=g) for n in sizes]
    alpha = 0.999
    want = [alpha * t + (1 - alpha) * s for t, s in zip(ts, ss)]
    tg, sg = [t.to(dev()) for t in ts], [s.to(dev()) for s in ss]
    ema_update_tensors(tg, sg, alpha)
    for a, b in zip(tg, want):
        assert np.array_equal(a.cpu().numpy().view(np.uint32), b.numpy().view(np.uint32))


# ------------------------------------------------------------------------------------------------ fused means + update
@pytest.mark.parametrize("n,d,h,w,name,start_mean", [(8, 2048, 33, 65, "moving_average", True), (4, 4096, 9, 11, "mean", True),
                                                     (3, 100, 9, 11, "moving_average", False), (16, 256, 17, 19, "mean", True),
                                                     (1, 8192, 9, 11, "moving_average", True), (2, 30, 5, 5, "mean", True)])
def test_finish_kernel_equals_means_plus_update_bitwise(D, n, d, h, w, name, start_mean):
    """diga_centroid_finish (one cluster launch: means, vector.sum() over distributed shared memory, recurrence) against
    diga_centroid_means + diga_centroid_update on the same class sums: centroids and counts bit-equal, incl. a zero-sum
    vector, an all-gated image and the start_mean switch at num == 100."""
    from diga_b200 import _lib as L, synthetic as S
    g = S.gen(71, "cuda")
    c = 19
    feat, out, lab = _inputs(n, d, h, w, c, g, labels=True)
    if n > 1:
        lab[1] = 255.0
    feat[0].masked_fill_((out.argmax(1)[0] == 1).unsqueeze(0), 0.0)
    assert L.lib.diga_centroid_finish_supported(n, d)
    a, b = D.Class_Features(c, d), D.Class_Features(c, d)
    start = (S.centroids(c, d, g), torch.full((c,), 98.0, device=dev()))
    for cf in (a, b):
        cf.objective_vectors, cf.objective_vectors_num = start[0].clone(), start[1].clone()
    for _ in range(3):                                         # crosses num == 100
        a.update_from_features(feat, out, lab, name, start_mean)                       # fused path
        vec, vecsum, valid = b._masked_means(feat, out, lab)                           # separate kernels
        L.check(L.lib.diga_centroid_update(vec.data_ptr(), vecsum.data_ptr(), valid.data_ptr(), n, c, d, b.objective_vectors.data_ptr(),
                                           b.objective_vectors_num.data_ptr(), {"mean": 0, "moving_average": 1}[name], int(start_mean),
                                           1e-4, L.stream()))
    assert torch.equal(a.objective_vectors_num, b.objective_vectors_num)
    assert torch.equal(a.objective_vectors, b.objective_vectors)
    # the optional outputs of the fused kernel equal the means kernel's (vecsum: another summation order, 1e-6)
    sums_p, counts_p, (_, _, _, hw), ws = a._class_sums(feat, out, lab)        # pointers into the scratch tensor `ws`
    v2, s2, ok2 = torch.empty_like(vec), torch.empty_like(vecsum), torch.empty_like(valid)
    L.check(L.lib.diga_centroid_finish(sums_p, counts_p, n, c, d, hw, v2.data_ptr(), s2.data_ptr(), ok2.data_ptr(),
                                       None, None, 0, 1, 1e-4, L.stream()))
    assert torch.equal(v2, vec) and torch.equal(ok2, valid)
    assert torch.allclose(s2, vecsum, rtol=1e-5, atol=1e-5 * float(vec.abs().max()) * 4)
    assert not L.lib.diga_centroid_finish_supported(17, 2048) and not L.lib.diga_centroid_finish_supported(2, 8 * 2048 + 1)


def test_accum_from_plain_class_map_and_built_class_words(D):
    """The C-ABI pieces a caller with its own class map uses: diga_centroid_clsw_build + diga_centroid_accum without the
    per-image counts (every class is then cleared, reduced and written) against the einsum of the one-hot map."""
    from diga_b200 import _lib as L, synthetic as S
    g = S.gen(12, "cuda")
    for n, d, h, w, c in ((2, 256, 65, 129, 19), (1, 128, 9, 12, 16), (2, 64, 5, 7, 19)):
        hw = h * w
        feat = S.features((n, d, h, w), g)
        cls = torch.randint(0, c, (n, hw), device=dev(), dtype=torch.uint8, generator=g)
        cls[:, ::7] = 255                                                   # gated-out pixels
        clsw = torch.empty((int(L.lib.diga_centroid_clsw_bytes(n, hw)) // 4,), dtype=torch.int32, device=dev())
        sums = torch.full((n, c, d), float("nan"), device=dev())
        L.check(L.lib.diga_centroid_clsw_build(cls.data_ptr(), n, c, hw, clsw.data_ptr(), L.stream()))
        L.check(L.lib.diga_centroid_accum(feat.data_ptr(), cls.data_ptr(), None, clsw.data_ptr(), n, d, c, hw, sums.data_ptr(), L.stream()))
        onehot = torch.nn.functional.one_hot(cls.long(), 256)[:, :, :c].double()               # [n, hw, c]; 255 -> no class
        want = torch.einsum("ndp,npc->ncd", feat.reshape(n, d, hw).double(), onehot)
        assert_normwise(sums, want, what=f"class sums {(n, d, h, w)}")
        # without class words the round-1 byte-map kernel answers; same sums
        sums2 = torch.empty_like(sums)
        L.check(L.lib.diga_centroid_accum(feat.data_ptr(), cls.data_ptr(), None, None, n, d, c, hw, sums2.data_ptr(), L.stream()))
        assert_normwise(sums2, want, what="byte-map kernel")
