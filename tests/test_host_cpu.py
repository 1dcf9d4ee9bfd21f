"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header declares, host logic
(class choice, sharding, the multi-rank centroid reduction under gloo) and the product/oracle separation."""
import ctypes
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "diga_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(diga_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from diga_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/diga_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.diga_version() >= 100
    assert _lib.launch_count() == 0


def test_ctypes_table_matches_header_prototypes():
    """Arity and scalar / pointer kind of every ctypes signature against the C prototype in include/diga_b200.h (a drifted
    table would pass pointers in float slots without any error)."""
    import ctypes as C
    from diga_b200 import _lib
    src = open(os.path.join(ROOT, "include", "diga_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//.*", "", src)
    decls = re.findall(r"\b(diga_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert {n for n, _ in decls} == set(_lib.SIGNATURES)
    kinds = {"ptr": (C.c_void_p, C.c_char_p), "int64_t": (C.c_int64,), "float": (C.c_float,), "double": (C.c_double,),
             "int": (C.c_int,), "size_t": (C.c_size_t,), "uint32_t": (C.c_uint32, C.c_uint)}

    def kind(param):
        param = param.strip()
        if "*" in param or "diga_stream_t" in param:
            return "ptr"
        return param.replace("const", "").split()[0]

    for name, params in decls:
        plist = [] if params.strip() in ("", "void") else params.split(",")
        argtypes = _lib.SIGNATURES[name][1]
        assert len(plist) == len(argtypes), f"{name}: header has {len(plist)} parameters, ctypes table {len(argtypes)}"
        for i, (prm, at) in enumerate(zip(plist, argtypes)):
            assert at in kinds[kind(prm)], f"{name} parameter {i} ({prm.strip()}): ctypes {at.__name__}"


def test_c_abi_argument_validation_without_gpu():
    """Validation happens before any CUDA call, so bad arguments are reported even on a GPU-less host."""
    from diga_b200 import _lib as L
    assert L.lib.diga_kd_fwd(None, None, 2, 19, 16, 0.5, None, None, None) == -1
    assert "null" in L.last_error()
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    assert L.lib.diga_kd_fwd(p, p, 3, 19, 1, 0.5, p, p, None) == -1 and "even" in L.last_error()
    assert L.lib.diga_pseudo_label(p, None, 1, 40, 1, None, None, p, None) == -1
    assert L.lib.diga_pseudo_label(p + 2, None, 1, 19, 1, None, None, p, None) == -2
    assert L.lib.diga_centroid_update(p, p, None, 1, 19, 2, p, p, 7, 1, 1e-4, None) == -1
    assert "no such updating way" in L.last_error()
    with pytest.raises(RuntimeError):
        L.check(-1)


def test_wrappers_refuse_cpu_tensors():
    import diga_b200 as D
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D.distillation_loss(torch.zeros(2, 19, 4, 4), torch.zeros(2, 19, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D.classmix(torch.zeros(1, 4, 4, dtype=torch.int64), torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4))
    with pytest.raises(RuntimeError):
        D.process_label(torch.zeros(1, 1, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diga_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in txt, f"{f} reads the reference tree"


def test_class_choice_matches_oracle_rng_stream():
    from diga_b200.classmix import select_classes
    from oracle import diga_oracle as O
    g = torch.Generator().manual_seed(1)
    sl = torch.randint(0, 19, (4, 16, 16), generator=g)
    sl[1, :4] = 255
    sl[3] = 255
    present = [torch.unique(sl[i]).tolist() for i in range(4)]
    assert select_classes(present, random.Random(9)) == O.classmix_select_classes(sl, random.Random(9))


def test_shard_indices_partition():
    from diga_b200.parallel import shard_indices
    for world in (1, 2, 4, 8):
        parts = [shard_indices(2975, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(2975))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_feature_hw_matches_survey():
    from diga_b200.synthetic import feature_hw
    assert feature_hw(256, 512) == (33, 65)
    assert feature_hw(512, 896) == (65, 113)
    assert feature_hw(512, 1024) == (65, 129)
    assert feature_hw(1024, 2048) == (129, 257)


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from diga_b200.parallel import shard_indices, new_mean_accumulator, finish_mean_pass
from oracle import diga_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
C, Dm, n_img = 19, 24, 14
g = torch.Generator().manual_seed(0)
feats = [torch.randn((1, Dm, 9, 11), generator=g) for _ in range(n_img)]
outs = [3 * torch.randn((1, C, 9, 11), generator=g) for _ in range(n_img)]
for o in outs: o[:, :6] += 3
# every rank: the per-image vectors of its own shard (CPU oracle stands in for the kernels, host logic under test)
acc = new_mean_accumulator(C, Dm, "cpu")
cf = O.ClassFeaturesOracle(C, Dm)
for i in shard_indices(n_img, rank, world):
    vec, ids = cf.calculate_mean_vector(feats[i], outs[i])
    for v, c in zip(vec, ids):
        if v.sum().item() != 0:
            acc[c, :Dm] += v.reshape(-1); acc[c, Dm] += 1
vectors, num = finish_mean_pass(acc)
ref = O.centroid_pass(feats, outs, C, Dm)            # single-rank sequential reference over the union
err = (vectors - ref.objective_vectors).abs().max().item()
scale = ref.objective_vectors.abs().max().item()
assert err <= 1e-5 * scale, (err, scale)
assert torch.equal(num, ref.objective_vectors_num), (num, ref.objective_vectors_num)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


WORKER_EXACT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from diga_b200.parallel import ShardedCentroidPass, global_row_order, finish_mean_pass, new_mean_accumulator
from oracle import diga_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
C, Dm, n_img, batch = 19, 12, 23, 3
g = torch.Generator().manual_seed(1)
nb = -(-n_img // batch)
feats = [torch.randn((min(batch, n_img - k * batch), Dm, 6, 7), generator=g) for k in range(nb)]
outs = [3 * torch.randn((f.shape[0], C, 6, 7), generator=g) for f in feats]
for o in outs: o[:, :4] += 3

def rows_of(cf, feat, out):
    """what the a6 kernels deliver per image: vec [n,C,D], vecsum [n,C], valid [n,C] (CPU oracle stands in for them)"""
    n = feat.shape[0]
    vec, vs, ok = torch.zeros(n, C, Dm), torch.zeros(n, C), torch.zeros(n, C, dtype=torch.uint8)
    for i in range(n):
        vectors, ids = cf.calculate_mean_vector(feat[i:i + 1], out[i:i + 1])
        for v, t in zip(vectors, ids):
            vec[i, t], vs[i, t], ok[i, t] = v.reshape(-1), v.sum(), 1
    return vec, vs, ok

ref = O.ClassFeaturesOracle(C, Dm)                    # the single-process loop over the union, three passes, low clamp
state = O.ClassFeaturesOracle(C, Dm)
sp = ShardedCentroidPass(state, n_img, batch)
assert (sp.rank, sp.world) == (rank, world)
for _ in range(3):
    ref = O.centroid_pass(feats, outs, C, Dm, cf=ref)
    for k in sp.my_batches():
        sp.add_rows(*rows_of(state, feats[k], outs[k]))
    gvec, gsum, gvalid = sp.gather()                   # the exchange under test (all_gather_into_tensor over gloo)
    for r in global_row_order(n_img, batch, world, sp.per_shard):     # the replay order the device kernel uses
        for t in range(C):
            if gvalid[r, t]:
                state.update_objective_SingleVector(t, gvec[r, t].numpy(), "mean")
    sp.reset()
assert torch.equal(state.objective_vectors_num, ref.objective_vectors_num)
assert torch.equal(state.objective_vectors, ref.objective_vectors), "ordered replay of the gathered rows != sequential loop"

# sum mode continued from a state and across the clamp: finish_mean_pass(acc, cf) == closed form of its documented rule
class St: pass
st = St(); st.objective_vectors = torch.ones(C, Dm); st.objective_vectors_num = torch.full((C,), 2990.0)
acc = new_mean_accumulator(C, Dm, "cpu")
acc[:, :Dm] = 10.0 * (rank + 1); acc[:, Dm] = 10.0            # each rank: 10 vectors per class whose sum is 10*(rank+1)
vec, num = finish_mean_pass(acc, st)
n = 10.0 * world; mean = sum(10.0 * (r + 1) for r in range(world)) / n
m = 10.0; k = n - m
obj1 = (1.0 * 2990.0 + mean * m) / 3000.0
rho = (3000.0 / 3001.0) ** k
want = obj1 * rho + (1 - rho) * mean
assert torch.allclose(vec, torch.full((C, Dm), want), rtol=1e-6), (vec[0, 0].item(), want)
assert torch.equal(num, torch.full((C,), 3000.0)) and st.objective_vectors is vec
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_exact_mode_gather_and_order_gloo(tmp_path, world):
    """Exact mode host logic on CPU ranks: shard assignment, all-gather of the row buffers, global replay order —
    N ranks == the sequential single-process loop bit for bit over three passes; plus the documented rule of the sum mode
    when it continues from a state across the 3000 clamp."""
    script = tmp_path / "worker_exact.py"
    script.write_text(WORKER_EXACT)
    port = 29900 + (os.getpid() % 90) + world
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out


def test_mean_pass_allreduce_world2_gloo(tmp_path):
    """N ranks over disjoint image shards + one all-reduce == the sequential single-rank running mean."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29500 + (os.getpid() % 400)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out


# ------------------------------------------------------------------------------------------------ f3 writer: PNG framing
def _label_patterns():
    rng = np.random.default_rng(11)
    coarse = rng.integers(0, 19, (6, 9)).astype(np.uint8)
    coarse[rng.random(coarse.shape) < 0.15] = 255
    blocks = np.kron(coarse, np.ones((16, 16), np.uint8))[:90, :131]
    return {"blocks": blocks, "constant": np.full((40, 600), 7, np.uint8), "noise": rng.integers(0, 256, (33, 70)).astype(np.uint8),
            "one_pixel": np.array([[200]], np.uint8), "column": rng.integers(0, 19, (50, 1)).astype(np.uint8),
            "run_258_edges": np.concatenate([np.full((3, n), 5, np.uint8) for n in (257, 258, 259, 260, 261, 262)], axis=1)}


@pytest.mark.parametrize("name", ["blocks", "constant", "noise", "one_pixel", "column", "run_258_edges"])
def test_png_oracle_stream_is_valid_zlib_and_frames_to_the_reference_file(name, tmp_path):
    """The oracle's token stream (oracle/png_oracle.py) is pinned against the standard decoders: zlib inflates it to the
    Up-filtered scanlines, and the framed file opens in Pillow to the same mode, palette and pixels as the file the
    reference's `colorize_mask(...).save(...)` writes (pseudolabel_generator.py:45-49, :104-105)."""
    import zlib
    from PIL import Image
    from oracle import png_oracle as P
    from diga_b200.pseudolabel import CITYSCAPES_PALETTE, colorize_mask, frame_png
    lab = _label_patterns()[name]
    stream = P.deflate_stream(lab)
    assert zlib.decompress(stream) == P.filtered_scanlines(lab).tobytes()
    ref_path, got_path = os.path.join(tmp_path, "ref.png"), os.path.join(tmp_path, "got.png")
    colorize_mask(lab).save(ref_path)
    with open(got_path, "wb") as f:
        f.write(frame_png(stream, *lab.shape))
    Image.open(got_path).verify()
    ref, got = Image.open(ref_path), Image.open(got_path)
    assert got.mode == ref.mode == "P" and got.size == ref.size
    assert got.getpalette() == ref.getpalette() == CITYSCAPES_PALETTE
    assert np.array_equal(np.array(got), np.array(ref)) and np.array_equal(np.array(got), lab)


def test_png_write_file_native_framing_matches_python_framing(tmp_path):
    """diga_png_write_file (host-only C entry point: CRC-32 + chunk framing) writes byte for byte what frame_png builds with
    the standard library's struct / zlib.crc32, reports unwritable paths, and validates its arguments."""
    import ctypes
    from oracle import png_oracle as P
    from diga_b200 import _lib as L
    from diga_b200.pseudolabel import CITYSCAPES_PALETTE, frame_png
    pal = (ctypes.c_ubyte * 768)(*CITYSCAPES_PALETTE)
    for name, lab in _label_patterns().items():
        stream = P.deflate_stream(lab)
        buf = (ctypes.c_ubyte * len(stream)).from_buffer_copy(stream)
        path = os.path.join(tmp_path, name + ".png")
        assert L.lib.diga_png_write_file(os.fsencode(path), buf, len(stream), lab.shape[0], lab.shape[1], pal, 768) == 0
        assert open(path, "rb").read() == frame_png(stream, *lab.shape)
    assert L.lib.diga_png_write_file(os.fsencode(os.path.join(tmp_path, "no_such_dir", "x.png")), buf, len(stream), 3, 3, pal, 768) == -5
    assert "cannot open" in L.last_error()
    assert L.lib.diga_png_write_file(os.fsencode(path), buf, len(stream), 3, 3, pal, 767) == -1
    assert L.lib.diga_png_write_file(None, buf, len(stream), 3, 3, pal, 768) == -1


def test_png_oracle_random_run_structures_inflate():
    """Seeded sweep over shapes and run structures (run lengths around the 258-byte match limit, values on both sides of the
    8/9-bit literal boundary): the oracle stream always inflates to the filtered scanlines."""
    import zlib
    from oracle import png_oracle as P
    rng = np.random.default_rng(2024)
    for _ in range(60):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 700))
        n_runs = int(rng.integers(1, 12))
        row = np.repeat(rng.integers(0, 256, n_runs), rng.integers(1, 300, n_runs))[:w]
        row = np.pad(row, (0, w - row.size), mode="edge").astype(np.uint8)
        lab = np.tile(row, (h, 1))
        flip = rng.random((h, w)) < rng.choice([0.0, 0.01, 0.2])
        lab[flip] = rng.integers(0, 256, int(flip.sum()))
        assert zlib.decompress(P.deflate_stream(lab)) == P.filtered_scanlines(lab).tobytes()


def test_png_table_constants_match_the_fixture():
    """csrc/png_table.inc (what the kernels index) and tests/golden/png_table.json (what the oracle derives its codes from)
    describe the same static Huffman table: the .inc patterns are the bit-reversed canonical codes of the fixture's lengths,
    the code is complete (Kraft sum 1, which zlib requires of a literal/length code), and the header words spell the
    fixture's header bits."""
    import json
    from oracle import png_oracle as P
    src = open(os.path.join(ROOT, "diga_b200", "csrc", "png_table.inc")).read()

    def arr(name):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", src, flags=re.S).group(1)
        return [int(tok.rstrip("u"), 0) for tok in body.replace("\n", " ").split(",")]

    fix = json.load(open(os.path.join(ROOT, "tests", "golden", "png_table.json")))
    lengths = fix["lit_lengths"]
    assert abs(sum(2.0 ** -n for n in lengths) - 1.0) < 1e-12 and max(lengths[:256]) <= 12
    t = P._table()
    assert arr("kPngLitLen") == lengths[:256] and arr("kPngLenLen") == lengths[257:286]
    assert arr("kPngLitPat") == [int(v) for v in t["pat"][:256]]
    assert arr("kPngLenPat") == [int(v) for v in t["pat"][257:286]]
    assert int(re.search(r"kPngEobPat = (\d+)u", src).group(1)) == int(t["pat"][256])
    assert int(re.search(r"kPngEobLen = (\d+)", src).group(1)) == lengths[256]
    nbits = int(re.search(r"kPngHeaderBits = (\d+)", src).group(1))
    words = arr("kPngHeaderWords")
    assert nbits == len(fix["header_bits"])
    assert "".join(str((words[i // 32] >> (i % 32)) & 1) for i in range(nbits)) == fix["header_bits"]
    assert int(re.search(r"kPngMaxLitBits = (\d+)", src).group(1)) == max(lengths[:256])


def test_png_crc_combine_constants_and_operator():
    """csrc/png.cu joins the CRC-32 remainders of 256 byte ranges with x^(8 n) mod P (zlib's crc32_combine construction).  The
    kernel's constants `kCrcX2n[k] = x^(2^k) mod P` are re-derived here by squaring in GF(2)[x]/P, and the operator itself
    (restated in Python) must reproduce zlib.crc32 over a split message for every split point class."""
    import re
    import zlib
    POLY = 0xEDB88320

    def multmodp(a, b):
        p, m = 0, 1 << 31
        while True:
            if a & m:
                p ^= b
                if (a & (m - 1)) == 0:
                    break
            m >>= 1
            if m == 0:
                break
            b = (b >> 1) ^ POLY if b & 1 else b >> 1
        return p

    want = [1 << 30]
    for _ in range(31):
        want.append(multmodp(want[-1], want[-1]))
    src = open(os.path.join(ROOT, "diga_b200", "csrc", "png.cu")).read()
    body = re.search(r"kCrcX2n\[32\]\s*=\s*\{(.*?)\};", src, re.S).group(1)
    body = re.sub(r"//[^\n]*", "", body)
    got = [int(v.rstrip("u"), 16) for v in re.findall(r"0x[0-9a-fA-F]+u?", body)]
    assert got == want

    def x8n(nbytes):                      # x^(8 nbytes) mod P from the table, as crc_x8n() in the kernel
        p, k, n = 1 << 31, 3, nbytes
        while n:
            if n & 1:
                p = multmodp(want[k & 31], p)
            n >>= 1
            k += 1
        return p

    def raw(data):                        # CRC register without the pre / post conditioning
        c = 0
        for byte in data:
            c ^= byte
            for _ in range(8):
                c = (c >> 1) ^ POLY if c & 1 else c >> 1
        return c

    rng = np.random.default_rng(9)
    msg = rng.integers(0, 256, 700, dtype=np.uint8).tobytes()
    for cut in (0, 1, 4, 255, 256, 699, 700):
        a, b = msg[:cut], msg[cut:]
        joined = multmodp(x8n(len(b)), raw(a)) ^ raw(b)
        assert joined == raw(msg)
    total = raw(msg) ^ multmodp(x8n(len(msg)), 0xFFFFFFFF) ^ 0xFFFFFFFF
    assert total == zlib.crc32(msg)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (no GPU needed): one JSON line with the arm's keys, the same `config` / metric / unit the GPU
    arm prints, and — under torchrun — ranks other than 0 exit 0 without work or output."""
    import json
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "pixel-positions/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.CONFIG and line["metric"] == bench.METRIC and "l2" in line["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=120,
                           env=dict(env, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533"))
    assert other.returncode == 0 and other.stdout.strip() == ""
