#!/usr/bin/env python
"""Benchmark of the per-pixel adaptation hot path (BASELINE.json metric: pixels/sec, % of HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Step  = one pass of the hot path over one batch of synthetic input.  The headline workload is BASELINE
config 2 (configs[1]): symmetric-KD loss forward + backward on Cityscapes-shaped logits
[8,19,512,1024] (two views x batch 4) per GPU, called through the drop-in ``distillation_loss`` +
autograd.  ``value`` = pixel-positions/s over all ranks with inputs resident in HBM; ``e2e`` = the same
call with HOST (pinned) inputs, H2D copies and the D2H read of the loss inside the timed region.
``roofline`` is the dominant kernel (KD backward, 228 B/px) timed with CUDA events inside the timed region.
``stages`` carries the other §8 rows (pseudo-label, selection, ClassMix, centroids, distance) measured in the
same run; ``cpu_baseline`` is the oracle port timed on this box's host cores on a bounded sample.

Multi-GPU: images shard across ranks with no data-path collective for a1-a5 (weak scaling); the centroid
stage ends in the one real exchange, an NCCL all-reduce of the [19, D+1] buffer.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

C = 19
KD_SHAPE = (8, C, 512, 1024)          # config 2: B=4 per view
KD_BYTES_FWD = 2 * C * 4              # teacher + student read
KD_BYTES_BWD = 3 * C * 4              # teacher + student read, dstudent written
UPSTREAM = 0.25
METRIC = "pixels/sec (symmetric KD loss fwd+bwd; per-row figures for pseudo-label+selection and centroids in roofline.rows)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML, 2 ms period)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa_node(index: int):
    """Pin this process (and therefore the first touch of its pinned host buffers and its launch thread) to the CPUs of the
    NUMA node the GPU hangs off.  Best effort: returns {"node": n, "cpus": k} or {"node": None, ...} when sysfs has no answer."""
    info = {"node": None, "cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                                   # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"node": node, "cpus": len(allowed)}
    except Exception:                                             # noqa: BLE001  (no NVML / sysfs: leave the affinity alone)
        pass
    return info


def launch_pairs_ms(launch, iters, warm=3, tries=3):
    """Average duration of one launch: a CUDA event pair around every launch on the launching stream.  The stream is first
    gated with a ~3 ms spin kernel so that every pair is already queued when the GPU reaches the first one (an event pair
    otherwise also times whatever the host does between ``record`` and the launch: one scheduler hiccup of the Python thread
    inside a pair is worth ten launches).  A try whose slowest pair exceeds 1.5x its median is repeated (at most ``tries``);
    the try with the lowest mean is returned together with what was seen."""
    for i in range(warm):
        launch(i)
    best = None
    for t in range(tries):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        torch.cuda.synchronize()
        torch.cuda._sleep(6_000_000)
        for i, (a, b) in enumerate(evs):
            a.record()
            launch(i)
            b.record()
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for a, b in evs]
        cur = {"mean_ms": statistics.mean(ts), "median_ms": statistics.median(ts), "max_ms": max(ts), "launches": iters, "tries": t + 1}
        if best is None or cur["mean_ms"] < best["mean_ms"]:
            best = dict(cur)
        best["tries"] = t + 1
        if cur["max_ms"] <= 1.5 * cur["median_ms"]:
            break
    return best["mean_ms"], best


def time_loop(fn, iters, warmup, world=1, device=None, graph=False):
    """CUDA-event timing of ``iters`` calls, barrier + synchronize on both sides, max over ranks -> ms per call.
    ``graph=True`` captures one call (a sync-free public-API call) in a CUDA graph and times its replays, which removes
    the host-side launch gaps of the Python call chain; the eager figure is reported next to it by the caller."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side):
                keep = fn()                                    # noqa: F841  (outputs stay alive with the graph)
        torch.cuda.current_stream().wait_stream(side)
        call = gph.replay
        for _ in range(3):
            call()
        torch.cuda.synchronize()
    else:
        call = fn
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world, device) / iters


# --------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the reference's OWN CPU code on the host cores (baseline/_ref, staged by
# oracle/make_ref.py), the oracle port when that tree is absent
# --------------------------------------------------------------------------------------------------
CONFIG = {"workload": "config 2: symmetric KD loss fwd+bwd (distillation_loss + autograd), logits [8,19,512,1024] "
                      "(two views x batch 4) per GPU, fp32",
          "shape": list(KD_SHAPE), "pixel_positions_per_step_per_gpu": KD_SHAPE[0] * KD_SHAPE[2] * KD_SHAPE[3],
          "l2": "no flush needed: each step reads 638 MB of logits (>> 126 MB L2) from one of 3 rotating input sets"}


def cpu_functions():
    """(namespace, kind): the reference's own functions when a reference tree is reachable ($DIGA_REF, /root/reference or
    the staged baseline/_ref), else the oracle port.  CPU only: ``.cuda()`` is made a no-op for the importing process, which
    is why the CPU legs run in their own process (``--impl reference`` / ``--cpu-baseline-json``)."""
    import types
    from oracle import diga_oracle as O
    try:
        from oracle import ref_loader as R
        if R.available("G"):
            torch.Tensor.cuda = lambda self, *a, **k: self          # the reference hard-codes .cuda() (utils.py:160, calc_centroids.py:95,168)
            torch.nn.Module.cuda = lambda self, *a, **k: self
            ns = R.load("G")
            return types.SimpleNamespace(distillation_loss=ns.distillation_loss, Class_Features=ns.Class_Features,
                                         where=R.REF_ROOT), "reference"
    except Exception as e:                                          # noqa: BLE001  (any import problem -> the port)
        sys.stderr.write(f"bench: reference tree not usable ({type(e).__name__}: {e}); timing the oracle port\n")
    return types.SimpleNamespace(distillation_loss=O.distillation_loss,
                                 Class_Features=lambda numbers=19: O.ClassFeaturesOracle(numbers), where="oracle/diga_oracle.py"), "port"


def cpu_kd_step(fn, t, s, g):
    s = s.detach().requires_grad_(True)
    loss = fn(t, s, 0.5)
    (grad,) = torch.autograd.grad(loss, s, grad_outputs=g)
    return loss, grad


def cpu_kd_inputs():
    gen = torch.Generator().manual_seed(1234)
    return 3.0 * torch.randn(KD_SHAPE, generator=gen), 3.0 * torch.randn(KD_SHAPE, generator=gen), torch.tensor(UPSTREAM)


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every core this process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(n)
    return torch.get_num_threads()


def best_of(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def config1_cpu_stages(ref, threads):
    """BASELINE config 1: the reference CPU path of a3, a5 + a4 and a6 + a7 for ONE 512x256 image (H=256, W=512, features
    33x65), D in {256, 2048}, all host threads and one thread, best of 5.  a5/a6/a7 run the reference's own
    Class_Features methods; a3 and a4 are inline script blocks (pseudolabel_generator.py:77-85, self_training.py:298-304),
    timed through their line-for-line restatement in oracle/diga_oracle.py."""
    from oracle import diga_oracle as O
    gen = torch.Generator().manual_seed(1234)
    hh, ww, h, w = 256, 512, 33, 65
    lf, lds = 3.0 * torch.randn((1, C, h, w), generator=gen), 3.0 * torch.randn((1, C, 17, 33), generator=gen)
    pseudo = torch.randint(0, C, (1, hh, ww), generator=gen)
    out = 3.0 * torch.randn((1, C, h, w), generator=gen)
    stages = []

    def timed(name, px, unit, fn, rows):
        rec = {"stage": name, "units": px, "unit": unit, "rows": rows}
        for label, nthr in (("all_threads", threads), ("one_thread", 1)):
            torch.set_num_threads(nthr)
            sec = best_of(fn)
            rec[f"cpu_ms_{label}"] = sec * 1e3
            rec[f"cpu_units_per_s_{label}"] = px / sec
        torch.set_num_threads(threads)
        stages.append(rec)

    timed("a3 pseudo-label (2 up-samplings, max, softmax, numpy argmax + max, uint8)", hh * ww, "px",
          lambda: O.pseudo_label_to_uint8(O.pseudo_label_two_scale(lf, lds, (hh, ww))[0]), "pseudolabel_generator.py:77-85,92")
    for d in (256, 2048):
        feat = torch.randn((1, d, h, w), generator=gen)
        cf = ref.Class_Features(numbers=C)
        cf.objective_vectors = 0.5 * torch.randn((C, d), generator=gen)
        cf.objective_vectors_num = torch.zeros((C,))

        def rectify():
            return O.consensus_select(pseudo, cf.get_centroid_weight(feat), (hh, ww))

        def centroids():
            vectors, ids = cf.calculate_mean_vector(feat, out)
            for t in range(len(ids)):
                cf.update_objective_SingleVector(ids[t], vectors[t].detach().cpu().numpy(), "mean")

        timed(f"a5 + a4 prototype weights + consensus selection, D={d}", hh * ww, "px", rectify,
              "calc_centroids.py:166-176 + self_training.py:298-304")
        timed(f"a6 + a7 calculate_mean_vector + running mean update, D={d}", h * w, "feature-px", centroids,
              "calc_centroids.py:120-164, loop :75-78")
    return stages


def cpu_baseline(budget_s=25.0):
    """The CPU arm on THIS box's host cores: config 2 (the headline workload, full [8,19,512,1024] steps, best within a
    bounded budget) and the config-1 stages.  Runs in a process of its own (see cpu_functions)."""
    threads = use_all_host_threads()
    ref, kind = cpu_functions()
    t, s, g = cpu_kd_inputs()
    cpu_kd_step(ref.distillation_loss, t, s, g)                  # first touch + thread pool start-up
    best, reps = float("inf"), 0
    t_end = time.perf_counter() + budget_s
    while reps < 3 or (time.perf_counter() < t_end and reps < 12):
        t0 = time.perf_counter()
        cpu_kd_step(ref.distillation_loss, t, s, g)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    px = KD_SHAPE[0] * KD_SHAPE[2] * KD_SHAPE[3]
    del t, s
    what = "the reference's own util/loss.py:125-143 (imported from " + ref.where + ")" if kind == "reference" else \
        "oracle port of util/loss.py:125-143 (no reference tree on this box)"
    return {"value": px / best, "unit": "pixel-positions/s", "cores": threads, "kind": kind,
            "sample": f"{what} + autograd on the full [8,19,512,1024] step, best of {reps} steps after one warm-up, "
                      f"{threads} threads of {os.cpu_count()} logical cores",
            "ms_per_step": best * 1e3,
            "stages": config1_cpu_stages(ref, threads)}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of config 2 on all host threads, the SAME workload and
    config dict as the GPU arm (full [8,19,512,1024] per step; W warm-ups, K timed steps)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = use_all_host_threads()
    ref, kind = cpu_functions()
    t, s, g = cpu_kd_inputs()
    for _ in range(max(args.warmup, 1)):
        cpu_kd_step(ref.distillation_loss, t, s, g)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_kd_step(ref.distillation_loss, t, s, g)
    dt = time.perf_counter() - t0
    px = KD_SHAPE[0] * KD_SHAPE[2] * KD_SHAPE[3]
    value = px * args.steps / dt
    what = ("the reference's own distillation_loss (util/loss.py:125-143, imported from " + ref.where + ")") if kind == "reference" \
        else "oracle port of util/loss.py:125-143"
    line = {"impl": "reference", "metric": METRIC, "value": value,
            "unit": "pixel-positions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": "pixel-positions/s", "cores": threads, "kind": kind,
                             "sample": f"{what} + autograd, full [8,19,512,1024] per step, {threads} threads of "
                                       f"{os.cpu_count()} logical cores"},
            "e2e": {"value": value, "unit": "pixel-positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


# --------------------------------------------------------------------------------------------------
# secondary stages (the other rows of SURVEY.md §8), per rank
# --------------------------------------------------------------------------------------------------
def accum_kernel_rows(L, S, dev, peak):
    """The a6 accumulation KERNEL alone (diga_centroid_accum through the C ABI, event pair per launch) at config 4's batched
    shape [8,2048,65,129], for i.i.d. classes (worst case of the shared-memory accumulators) and 4x4-block maps."""
    g = S.gen(77 + dist_env()[0], dev)
    n, d, h, w = 8, 2048, 65, 129
    hw = h * w
    feats = [S.features((n, d, h, w), g) for _ in range(3)]
    rows = []
    for pattern in ("iid", "blocks4"):
        out = S.logits((n, C, h, w), g)
        if pattern == "blocks4":
            lab = S.block_labels(n, h, w, g, 4, C, 0.0)
            out = out + 12.0 * torch.nn.functional.one_hot(lab, C).permute(0, 3, 1, 2).float()
        cls = torch.empty((n, hw), dtype=torch.uint8, device=dev)
        clsw = torch.empty((int(L.lib.diga_centroid_clsw_bytes(n, hw)) // 4,), dtype=torch.int32, device=dev)
        counts = torch.empty((n, C), dtype=torch.int32, device=dev)
        sums = torch.empty((n, C, d), dtype=torch.float32, device=dev)
        st = L.stream()
        L.check(L.lib.diga_centroid_assign(out.data_ptr(), None, n, C, hw, cls.data_ptr(), counts.data_ptr(), clsw.data_ptr(), st))

        def launch(i):
            L.check(L.lib.diga_centroid_accum(feats[i % 3].data_ptr(), cls.data_ptr(), counts.data_ptr(), clsw.data_ptr(), n, d, C, hw,
                                              sums.data_ptr(), st))

        ms, _ = launch_pairs_ms(launch, 30)
        rows.append(("a6", f"centroid_accum_lean_kernel ({'i.i.d. classes' if pattern == 'iid' else '4x4-block class maps'})", "hbm",
                     "feature-px", n * hw, d * 4 + 1, ms, "CUDA event pair per launch, C-ABI launches queued behind a spin-kernel gate, 3 rotating 550 MB inputs",
                     "accum_dram_bytes_per_launch" if pattern == "iid" else None))
    return rows


def config1_gpu_stages(D, S, dev):
    """The GPU stages at BASELINE config 1's shapes (ONE 512x256 image: launch-latency territory), keyed like the CPU stages."""
    from diga_b200.util.loss import distillation_loss  # noqa: F401
    g = S.gen(1234, dev)
    hh, ww, h, w = 256, 512, 33, 65
    lf, lds = S.logits((1, C, h, w), g), S.logits((1, C, 17, 33), g)
    pseudo = torch.randint(0, C, (1, hh, ww), device=dev, generator=g)
    out = S.logits((1, C, h, w), g)
    res = {}

    def put(name, fn):
        res[name] = {"gpu_ms": time_loop(fn, 50, 5, graph=True), "gpu_eager_ms": time_loop(fn, 50, 5),
                     "gpu_how": "public API, CUDA-graph replay (gpu_ms) and eager call (gpu_eager_ms)"}

    put("a3 pseudo-label (2 up-samplings, max, softmax, numpy argmax + max, uint8)",
        lambda: D.pseudo_label_two_scale(lf, lds, (hh, ww), want_conf=False))
    for d in (256, 2048):
        feat = S.features((1, d, h, w), g)
        cf = D.Class_Features(C, d)
        cf.objective_vectors = S.centroids(C, d, g)
        put(f"a5 + a4 prototype weights + consensus selection, D={d}",
            lambda cf=cf, feat=feat: D.consensus_select(pseudo, cf.get_centroid_weight(feat), (hh, ww)))
        put(f"a6 + a7 calculate_mean_vector + running mean update, D={d}",
            lambda cf=cf, feat=feat: cf.update_from_features(feat, out, None, "mean"))
    return res


N_SET = 2975            # Cityscapes train: the target set of calc_centroids.py / pseudolabel_generator.py (BASELINE configs 4, 5)


def bench_stages(D, S, dev, peak, world):
    """Every other row of SURVEY.md §8 through its public API call.  Workload sizes are fixed (never tied to --steps):
    20 timed calls after 5 warm-ups per stage; the whole-set stages always run the full 2975-image set."""
    import random
    from diga_b200 import parallel as P
    g = S.gen(4321 + dist_env()[0], dev)
    iters, warm = 20, 5
    out = {}

    def add(name, px, bytes_per_px, fn, unit="px", extra=None, sync_free=True):
        """Times the public-API call ``fn``: replayed from a CUDA graph when it is sync-free (``ms``), eagerly otherwise;
        the eager figure (Python launch chain included) is always reported as ``eager_ms``."""
        eager = time_loop(fn, iters, warm, world, dev)
        ms = time_loop(fn, iters, warm, world, dev, graph=True) if sync_free else eager
        gbs = px * bytes_per_px / (ms * 1e-3) / 1e9
        out[name] = {"px_per_s": px / (ms * 1e-3) * world, "unit": f"{unit}/s (all ranks)", "ms": ms, "eager_ms": eager,
                     "algo_bytes_per_px": bytes_per_px, "gbs_per_gpu": gbs, "frac_hbm": gbs / peak}
        if extra:
            out[name].update(extra(ms) if callable(extra) else extra)

    # a3 pseudo-label at 2048x1024, one and two scales (config 5 resolution), 4 images per call
    n, hh, ww = 4, 1024, 2048
    z = S.logits((n, C, hh, ww), g)
    z2 = S.logits((n, C, hh, ww), g)
    add("pseudo_label_1scale", n * hh * ww, 4 * C + 5, lambda: D.pseudo_label(z))
    add("pseudo_label_2scale", n * hh * ww, 8 * C + 5, lambda: D.pseudo_label(z, z2))
    del z, z2
    # fused two-scale from stride-8 logits (next row f1): 5 B/px of HBM writes only, ALU-bound
    l1, l2 = S.logits((n, C, 129, 257), g), S.logits((n, C, 65, 129), g)
    add("pseudo_label_fused_upsample", n * hh * ww, 5, lambda: D.pseudo_label_two_scale(l1, l2, (hh, ww)),
        extra={"note": "replaces two bilinear up-samplings (2 x 76 B/px written + read) and the 157 B/px kernel; ALU-bound"})
    add("pseudo_label_fused_upsample_labels_only", n * hh * ww, 1,
        lambda: D.pseudo_label_two_scale(l1, l2, (hh, ww), want_conf=False),
        extra={"note": "what pseudolabel_generator.py keeps: the uint8 label map (the confidence is discarded there)"})

    # next row f5: evaluation confusion matrix (util/metrics.py:32-41) on the labels the fused kernel just produced (uint8)
    # against an int64 ground truth: 9 B/px read, nothing written but 19 x 19 counters
    from diga_b200.util.metrics import runningScore
    gt_eval = S.block_labels(n, hh, ww, g)
    pred_eval, _ = D.pseudo_label_two_scale(l1, l2, (hh, ww), want_conf=False)
    rs = runningScore(C)
    add("confusion_matrix_eval", n * hh * ww, 9, lambda: rs.update(gt_eval, pred_eval))
    del gt_eval, pred_eval

    # next row f3, reader half: decoded 1024x2048 label PNGs -> PIL-NEAREST resize to 512x1024 + id look-up -> int64
    from diga_b200.util.labels import resize_remap_labels, trainid_lut
    raw_ids = torch.randint(0, 34, (8, 1024, 2048), device=dev, dtype=torch.uint8, generator=g)
    lut_ids = trainid_lut()
    add("label_reader_resize_remap", 8 * 512 * 1024, 9, lambda: resize_remap_labels(raw_ids, (512, 1024), lut_ids),
        extra={"note": "CityLoader.py:93-95 + :113-132 after the PNG decode: 1 B gathered + 8 B written per output pixel"})
    del raw_ids

    # config 3 pieces: B=8 @512x1024, features [8,2048,65,129]
    b, hh, ww, h, w, d = 8, 512, 1024, 65, 129, 2048
    sl = S.block_labels(b, hh, ww, g)
    tl = S.perturb_labels(sl, g)
    xa, xb = S.images((b, 3, hh, ww), g), S.images((b, 3, hh, ww), g)
    rng = random.Random(1234)
    add("classmix_dacs_total", b * hh * ww, 68, lambda: D.classmix(sl, xa, xb, tl, rng=rng, return_mask=False), sync_free=False,
        extra={"note": "presence kernel (8 B/px) + D2H of 32 B/image + host random.sample + blend (60 B/px); one host sync"})
    from diga_b200.classmix import present_classes, select_classes
    classes = select_classes(present_classes(sl), rng)
    add("classmix_dacs_blend_kernel", b * hh * ww, 60,
        lambda: D.classmix(sl, xa, xb, tl, classes=classes, return_mask=False, assume_labelled=True))
    del xa, xb

    feat = S.features((b, d, h, w), g)
    cf = D.Class_Features(C, d)
    cf.objective_vectors = S.centroids(C, d, g)
    add("proto_distance_softmax_d2048", b * h * w, d * 4 + C * 4, lambda: cf.get_centroid_weight(feat), unit="feature-px")
    wl = cf.get_centroid_weight(feat)
    add("consensus_select", b * hh * ww, 24, lambda: D.consensus_select(tl, wl),
        extra={"note": "ALU-bound: 19 classes x (FMUL+FFMA+compare) per pixel on bit-exact interpolated weights"})

    # next rows f2 / f4: cross_entropy2d fwd+bwd on the student logits of config 2, EMA teacher update of a
    # ResNet-101-sized parameter set (44.5 M fp32 parameters in 314 tensors)
    del feat
    xs = S.logits((4, C, hh, ww), g)
    ce_t = S.block_labels(4, hh, ww, g)
    upc = torch.tensor(1.0, device=dev)

    def ce_step():
        x = xs.detach().requires_grad_(True)
        loss = D.cross_entropy2d(x, ce_t)
        (grad,) = torch.autograd.grad(loss, x, grad_outputs=upc)
        return loss, grad

    add("cross_entropy2d_fwd_bwd", 4 * hh * ww, (4 * C + 8) + (8 * C + 8), ce_step)
    del xs
    sizes = [64 * 3 * 49] + [256 * 64, 64 * 64 * 9, 64 * 256, 256, 256, 64] * 3 + [512 * 128 * 4, 128 * 128 * 9, 512] * 4 + \
            [1024 * 256, 256 * 256 * 9, 256 * 1024, 1024, 1024, 256] * 23 + [2048 * 512, 512 * 512 * 9, 512 * 2048, 2048] * 3 + \
            [19 * 2048 * 9] * 4
    from diga_b200.util.utils import ema_update_tensors
    tea = [torch.randn(sz, device=dev) for sz in sizes]
    stu = [torch.randn(sz, device=dev) for sz in sizes]
    n_par = sum(sizes)
    add("ema_teacher_update", n_par, 12, lambda: ema_update_tensors(tea, stu, 0.999), unit="parameters",
        extra={"tensors": len(sizes), "parameters": n_par})
    del tea, stu
    feat = S.features((b, d, h, w), g)

    # centroid accumulation + running update: i.i.d. random classes (worst case for the shared-memory accumulators)
    # and segmentation-like maps (4x4 feature-pixel blocks = 32x32 image blocks)
    logits_iid = S.logits((b, C, h, w), g)
    blocks = S.block_labels(b, h, w, g, 4, C, 0.0)
    logits_blk = S.logits((b, C, h, w), g) + 12.0 * torch.nn.functional.one_hot(blocks, C).permute(0, 3, 1, 2).float()
    per_img = lambda ms: {"images_per_s": b / (ms * 1e-3) * world}
    add("centroid_accumulate_update_d2048_iid_classes", b * h * w, d * 4 + C * 4 + 1,
        lambda: cf.update_from_features(feat, logits_iid, None, "mean"), unit="feature-px", extra=per_img)
    add("centroid_accumulate_update_d2048_blocky_classes", b * h * w, d * 4 + C * 4 + 1,
        lambda: cf.update_from_features(feat, logits_blk, None, "mean"), unit="feature-px", extra=per_img)

    # config 4 as the reference drives it: one 512x1024 image per call ([1,2048,65,129])
    f1, o1 = feat[:1].contiguous(), logits_iid[:1].contiguous()
    add("calc_centroids_per_image_call_d2048", h * w, d * 4 + C * 4 + 1,
        lambda: cf.update_from_features(f1, o1, None, "mean"), unit="feature-px",
        extra=lambda ms: {"images_per_s": 1 / (ms * 1e-3) * world, "note": "batch 1 per call like calc_centroids.py:67-78; "
                          "68.7 MB per launch is too small to fill the machine, batch the loader for throughput"})

    # config 5, one full-resolution image: labels from the two-scale logits, prototype weights at 129x257, consensus selection
    del feat
    f5 = S.features((1, d, 129, 257), g)
    l5a, l5b = S.logits((1, C, 129, 257), g), S.logits((1, C, 65, 129), g)

    def config5_image():
        lab, _ = D.pseudo_label_two_scale(l5a, l5b, (1024, 2048), want_conf=False)
        wts = cf.get_centroid_weight(f5)
        return D.consensus_select(lab, wts, want_feat_pseudo=False)

    add("config5_pseudo_label_plus_rectification_per_image", 1024 * 2048, (d * 4 + C * 4) * 129 * 257 / (1024 * 2048) + 3,
        config5_image, extra=lambda ms: {"images_per_s": 1 / (ms * 1e-3) * world,
                                         "note": "pseudo_label_two_scale + get_centroid_weight([1,2048,129,257]) + consensus_select "
                                                 "on the uint8 map; algorithmic bytes dominated by the 272 MB feature map"})
    del f5
    feat = S.features((b, d, h, w), g)

    # next row f1 for the loss consumers: the losses straight from the stride-8 logits (config 2's tensors before
    # nn.Upsample): KD fwd+bwd, cross_entropy2d fwd+bwd, and the shared-student seg + KD pair of self_training.py:348-352
    lo_t, lo_s = S.logits((8, C, h, w), g), S.logits((8, C, h, w), g)
    up_kd = torch.tensor(UPSTREAM, device=dev)

    def kd_up_step():
        x = lo_s.detach().requires_grad_(True)
        loss = D.distillation_loss_upsampled(lo_t, x, (hh, ww), 0.5)
        return loss, torch.autograd.grad(loss, x, grad_outputs=up_kd)

    def ce_up_step():
        x = lo_s[:4].detach().requires_grad_(True)
        loss = D.cross_entropy2d_upsampled(x, ce_t)
        return loss, torch.autograd.grad(loss, x, grad_outputs=upc)

    def seg_kd_up_step():
        x = lo_s.detach().requires_grad_(True)
        l_ce, l_kd = D.seg_distillation_losses_upsampled(lo_t, x, ce_t, 0.5)
        return l_ce, l_kd, torch.autograd.grad([l_ce, l_kd], x, grad_outputs=[upc, up_kd])

    lowres_bytes = 3 * C * 4 * (h * w) / (hh * ww)          # teacher + student read, gradient written, per output pixel
    add("kd_fused_upsample_fwd_bwd", 8 * hh * ww, lowres_bytes, kd_up_step,
        extra={"note": "distillation_loss_upsampled + autograd from [8,19,65,129]: replaces 2 up-samplings (76 B/px written each), "
                       "the 380 B/px KD pair and the up-sampling backward; latency-bound walk (exponentials by one multiply per row inside a source cell, MUFU only at cell crossings)"})
    add("cross_entropy2d_fused_upsample_fwd_bwd", 4 * hh * ww, 8 + 2 * C * 4 * (h * w) / (hh * ww), ce_up_step)
    ohem = D.OhemCrossEntropy(255, 0.7, 100000)

    def ohem_up_step():
        x = lo_s[:4].detach().requires_grad_(True)
        loss = ohem(x, ce_t)
        return loss, torch.autograd.grad(loss, x, grad_outputs=upc)

    add("ohem_cross_entropy_fused_upsample_fwd_bwd", 4 * hh * ww, 8 + 8 + 3 * 4 + 2 * C * 4 * (h * w) / (hh * ww), ohem_up_step,
        extra={"note": "OhemCrossEntropy (util/loss.py:65-122, Synthia tree) from [4,19,65,129]: pred/loss pass, exact radix "
                       "select of the min_kept-th probability (3 histogram passes), kept-pixel mean, CE gradient over the kept"})
    add("seg_plus_kd_fused_upsample_fwd_bwd", 8 * hh * ww, lowres_bytes + 4, seg_kd_up_step,
        extra={"note": "self_training.py:348-352 on the shared s_pred_cat_stu: CE on the source half + KD on both views, "
                       "one loss pass + one gradient pass over the stride-8 logits"})

    # config 3: the whole per-step hot path of train_DiGA_gta2city_self_training.py:259-356 (no backbone), B=8 @512x1024,
    # D=2048: ClassMix #1, prototype weights, consensus selection, ClassMix #2 (DACS), two label-gated centroid EMA
    # updates, seg losses + KD with backward to the stride-8 logits.  "dropin" = only the reference's own function names
    # swapped (losses on nn.Upsample outputs, torch interpolates), "fused" = the patched call sites of INTEGRATION.md.
    from diga_b200.calc_centroids import _labels_on_feature_grid
    import torch.nn.functional as F
    s_feat = S.features((b, d, h, w), g)
    t_pred, s_pred = S.logits((b, C, h, w), g), S.logits((b, C, h, w), g)
    rec, saug = S.images((b, 3, hh, ww), g), S.images((b, 3, hh, ww), g)
    tdata_aug, sdata = S.images((b, 3, hh, ww), g), S.images((b, 3, hh, ww), g)
    tea_cat, stu_cat = S.logits((2 * b, C, h, w), g), S.logits((2 * b, C, h, w), g)
    cross_low = S.logits((b, C, h, w), g)
    cf.objective_vectors_num = torch.full((C,), 150.0)
    rng3 = random.Random(99)
    lam = torch.tensor(0.25, device=dev)

    lazy_up = D.nn.Upsample(size=(hh, ww), mode="bilinear", align_corners=True)

    def st_step(fused, lazy=False):
        if fused:      # one presence pass over slabelv serves both ClassMix blocks.  In the script its 32 B/image host round trip
            # hides behind the backbone passes queued in between; this step has no backbone, so the statements that do not
            # depend on the class choice (a5, a4, and the two centroid updates, which read neither mix) are queued before the
            # host waits for it — same data dependencies, same results, ClassMix draws still in script order
            pres = D.present_classes_async(sl)
            wts = cf.get_centroid_weight(feat)                                                     # :301
            kept, feat_pseudo = D.consensus_select(tl, wts, (hh, ww))                              # :302-304
            cf.update_from_features(feat, t_pred, start_mean=False, labels_full=kept)              # :327-334
            cf.update_from_features(s_feat, s_pred, start_mean=False, labels_full=sl)              # :336-341
            _, mix1 = D.classmix(sl, rec, saug, rng=rng3, present=pres, return_mask=False)         # :259-275
            _, mix2, mixlabel = D.classmix(sl, tdata_aug, sdata, kept, rng=rng3, present=pres, return_mask=False)   # :306-325
        else:
            _, mix1 = D.classmix(sl, rec, saug, rng=rng3)                                          # :259-275
            wts = cf.get_centroid_weight(feat)                                                     # :301
            kept, feat_pseudo = D.consensus_select(tl, wts, (hh, ww))                              # :302-304
            _, mix2, mixlabel = D.classmix(sl, tdata_aug, sdata, kept, rng=rng3)                   # :306-325
        if not fused:  # (fused: label down-sampling — .float() + F.interpolate(nearest), :328-330, :336-337 — folded into the assign kernel above)
            cf.update_from_features(feat, t_pred, _labels_on_feature_grid(kept, (h, w)), start_mean=False)
            cf.update_from_features(s_feat, s_pred, _labels_on_feature_grid(sl, (h, w)), start_mean=False)
        stu = stu_cat.detach().requires_grad_(True)
        cpm = cross_low.detach().requires_grad_(True)
        if fused:      # loss weights (lambda_seg = 1, lambda_distil = 0.25, :102-103) known up front: one pass each
            part, l_src, l_kd = D.seg_distillation_total_upsampled(tea_cat, stu, sl, 1.0, 0.25, 0.5,   # :289,:348-352,:382
                                                                   targets_nonnegative=True)   # loader labels: trainIds or 255
            total = part + D.cross_entropy2d_upsampled(cpm, mixlabel)                              # :344,:355-356
        else:          # the reference's call sites as they are; `lazy`: its three nn.Upsample modules built from diga_b200.nn.Upsample
            up = lazy_up if lazy else (lambda x: F.interpolate(x, size=(hh, ww), mode="bilinear", align_corners=True))
            l_src = D.cross_entropy2d(up(stu[:b]), sl)                                             # :348-349
            l_kd = D.distillation_loss(up(tea_cat), up(stu), 0.5)                                  # :289,:351-352
            l_mix = D.cross_entropy2d(up(cpm), mixlabel)                                           # :344,:355
            total = (l_src + l_mix) + lam * l_kd                                                   # :356,:382
        g_stu, g_mix = torch.autograd.grad(total, [stu, cpm])
        return total, g_stu, g_mix, mix1, mix2

    step_bytes = 48 + 68 + 24 + 3 * (d * 4 + C * 4) * (h * w) / (hh * ww) + 16
    for name, fused, lazy in (("config3_self_training_step_dropin_functions", False, False),
                              ("config3_self_training_step_dropin_functions_lazy_upsample", False, True),
                              ("config3_self_training_step_fused_call_sites", True, False)):
        add(name, b * hh * ww, step_bytes, lambda fused=fused, lazy=lazy: st_step(fused, lazy), sync_free=False,
            extra={"note": "ClassMix x2 (one 32 B/image host round trip each, as the reference's random.sample needs), prototype "
                           "weights, consensus selection, 2 centroid EMA updates, CE x2 + KD with backward; eager (host syncs); "
                           "algorithmic bytes = the three [8,2048,65,129] feature reads + image/label traffic"})
    del s_feat, rec, saug, tdata_aug, sdata

    # config 4 shape of the exchange: mean pass + ONE all-reduce of [19, D+1] (NCCL over NVLink when world > 1)
    acc = P.new_mean_accumulator(C, d, dev)

    def mean_pass():
        acc.zero_()
        cf.accumulate_mean_pass(acc, feat, logits_iid)
        return P.finish_mean_pass(acc)

    add("centroid_mean_pass_allreduce_d2048", b * h * w, d * 4 + C * 4 + 1, mean_pass, unit="feature-px", sync_free=(world == 1),
        extra=lambda ms: {"images_per_s": b / (ms * 1e-3) * world, "allreduce_bytes": C * (d + 1) * 4})
    # ---- whole-set runs of BASELINE configs 4 and 5: this rank's share images[rank::world] of a synthetic 2975-image target
    # set, cycling through a pool of distinct pre-generated inputs (>> L2), timed once with CUDA events,
    # max over ranks ------------------------------------------------------------------------------------------------------
    del feat
    torch.cuda.empty_cache()
    n_set = N_SET
    mine = len(P.shard_indices(n_set, dist_env()[0], world))

    def time_once(fn):
        fn(min(mine, 16))                                                  # warm-up on a few images
        torch.cuda.synchronize()
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(mine)
        e1.record()
        torch.cuda.synchronize()
        barrier(world)
        return max_over_ranks(e0.elapsed_time(e1), world, dev)

    # config 4: calc_centroids 'mean' pass, 8 images per call, ONE all-reduce of [19, D+1] at the end
    pool4 = [(S.features((b, d, h, w), g), S.logits((b, C, h, w), g)) for _ in range(4)]       # 4 x 550 MB of features
    acc4 = P.new_mean_accumulator(C, d, dev)

    def run_config4(n_img):
        acc4.zero_()
        done, k = 0, 0
        while done < n_img:
            f, o = pool4[k % len(pool4)]
            take = min(b, n_img - done)
            cf.accumulate_mean_pass(acc4, f[:take], o[:take])
            done += take
            k += 1
        return P.finish_mean_pass(acc4)

    ms4 = time_once(run_config4)
    out["config4_calc_centroids_whole_set"] = {
        "mode": "sum (one all-reduce of [19, D+1]; equals the reference's running mean for the first pass, num + n <= 3000)",
        "images": n_set, "images_per_rank": mine, "ms": ms4, "images_per_s": n_set / (ms4 * 1e-3),
        "px_per_s": n_set * h * w / (ms4 * 1e-3), "unit": "feature-px/s (all ranks)",
        "image_px_per_s": n_set * hh * ww / (ms4 * 1e-3), "algo_bytes_per_px": d * 4 + C * 4 + 1,
        "gbs_per_gpu": mine * h * w * (d * 4 + C * 4 + 1) / (ms4 * 1e-3) / 1e9,
        "frac_hbm": mine * h * w * (d * 4 + C * 4 + 1) / (ms4 * 1e-3) / 1e9 / peak, "allreduce_bytes": C * (d + 1) * 4,
        "note": "calc_centroids.py:67-78 over the rank's share, batches of 8 x [2048,65,129], one NCCL all-reduce at the end; "
                "eager launches (the loop is the public API), events around the whole share"}
    # config 4, EXACT mode (SURVEY.md §8e option i): per-image class vectors kept, ONE all-gather of [images, 19, D] per pass
    # (463 MB at 2975 x 19 x 2048), ordered replay of the reference recurrence on every rank — bit-identical to the
    # single-process loop beyond the 3000 clamp and across the reference's five passes (calc_centroids.py:20-23)
    cf_exact = D.Class_Features(C, d)
    sp = P.ShardedCentroidPass(cf_exact, n_set, batch=b)
    gather_ms = [0.0]

    def run_config4_exact(n_img_unused):
        for k in sp.my_batches():
            f, o = pool4[k % len(pool4)]
            take = sp.batch_size_of(k)
            sp.add(f[:take], o[:take])
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        res = sp.finish()
        g1.record()
        gather_ms[0] = (g0, g1)
        return res

    ms4x = time_once(run_config4_exact)
    torch.cuda.synchronize()
    fin_ms = max_over_ranks(gather_ms[0][0].elapsed_time(gather_ms[0][1]), world, dev)
    out["config4_calc_centroids_whole_set_exact"] = {
        "mode": "exact (all-gather of the per-image vectors + ordered device replay; every rank ends with the single-process result)",
        "images": n_set, "images_per_rank": mine, "ms": ms4x, "images_per_s": n_set / (ms4x * 1e-3),
        "px_per_s": n_set * h * w / (ms4x * 1e-3), "unit": "feature-px/s (all ranks)",
        "algo_bytes_per_px": d * 4 + C * 4 + 1,
        "gbs_per_gpu": mine * h * w * (d * 4 + C * 4 + 1) / (ms4x * 1e-3) / 1e9,
        "frac_hbm": mine * h * w * (d * 4 + C * 4 + 1) / (ms4x * 1e-3) / 1e9 / peak,
        "gather_plus_replay_ms": fin_ms, "exchange": sp.exchange, "rows_bytes_per_rank": sp.per_shard * C * (d * 4 + 5),
        "note": "exchange 'multicast-stores' / 'peer-stores': the per-image vectors are stored into every rank's row buffer by the "
                "means kernel itself (NVLS multicast / NVLink peer stores on a side stream, symmetric memory), finish() = barrier + "
                "ordered replay; 'all-gather': finish() = all_gather_into_tensor (NCCL) + replay; 'local': one rank"}
    del pool4, sp, cf_exact

    # config 5: full-resolution pseudo-labels with prototype rectification.  The set is walked PER_CALL images per call (the
    # persistent distance kernel fills whole rounds and the two arg-max kernels whole waves: 100.3 / 88.2 / 83.3 / 79.6 us per image
    # at 1 / 2 / 4 / 8 images per call, tools/time_config5_batch.py); the reference's batch-1 loop is timed beside it.
    PER_CALL = 8
    pool5 = [(S.features((PER_CALL, d, 129, 257), g), S.logits((PER_CALL, C, 129, 257), g), S.logits((PER_CALL, C, 65, 129), g))
             for _ in range(3)]

    def config5_call(k, per_call):
        f, la, lb = pool5[k % len(pool5)]
        if per_call != PER_CALL:
            f, la, lb = f[:per_call], la[:per_call], lb[:per_call]
        lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
        return D.consensus_select(lab, cf.get_centroid_weight(f), want_feat_pseudo=False)

    def run_config5(n_img, per_call=PER_CALL):
        kept = None
        for k in range(-(-n_img // per_call)):
            kept = config5_call(k, min(per_call, n_img - k * per_call))
        return kept

    ms5 = time_once(run_config5)
    ms5_one = time_once(lambda n_img: run_config5(n_img, 1))
    px5 = 1024 * 2048
    out["config5_pseudo_labels_whole_set"] = {
        "images": n_set, "images_per_rank": mine, "images_per_call": PER_CALL, "ms": ms5, "images_per_s": n_set / (ms5 * 1e-3),
        "px_per_s": n_set * px5 / (ms5 * 1e-3), "unit": "px/s (all ranks)",
        "algo_bytes_per_px": (d * 4 + C * 4) * 129 * 257 / px5 + 3,
        "gbs_per_gpu": mine * ((d * 4 + C * 4) * 129 * 257 + 3 * px5) / (ms5 * 1e-3) / 1e9,
        "frac_hbm": mine * ((d * 4 + C * 4) * 129 * 257 + 3 * px5) / (ms5 * 1e-3) / 1e9 / peak,
        "ms_one_image_per_call": ms5_one,
        "note": "pseudolabel_generator.py:69-85 + the rectification of self_training.py:298-304: fused two-scale labels, "
                "prototype weights of [8,2048,129,257], consensus selection, 8 images per call (ms_one_image_per_call = the "
                "reference's batch-1 loop); no collective (image-sharded)"}
    # config 5 end to end INCLUDING the files (pseudolabel_generator.py:89-105): the kept maps go to 'P'-mode PNGs on local
    # disk through PseudoLabelWriter; wall clock from the first kernel to the last closed file.  The GPU encoder runs over
    # the rank's whole share, the Pillow encoder (what the reference does per image) over a bounded sample.
    import shutil
    import tempfile
    from diga_b200.pseudolabel import PseudoLabelWriter

    # tmpfs when there is one: the box's virtual disk throttles on dirty pages (460-790 ms for the same run on /tmp depending
    # on what was written before, tools/experiments/r01_config5_files.py), which is the VM's write-back, not this pipeline
    file_root = None
    try:
        if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) and shutil.disk_usage("/dev/shm").free > (4 << 30):
            file_root = "/dev/shm"
    except OSError:
        pass

    def run_config5_png(n_img, encoder, workers):
        out_dir = tempfile.mkdtemp(prefix="diga_pl_", dir=file_root)
        torch.cuda.synchronize()
        barrier(world)
        t0 = time.perf_counter()
        with PseudoLabelWriter(out_dir, workers=workers, slots=4, encoder=encoder, coalesce=8) as wr:
            for k in range(-(-n_img // PER_CALL)):
                m = min(PER_CALL, n_img - k * PER_CALL)
                kept, _ = config5_call(k, m)
                wr.submit(kept, [f"img_{k * PER_CALL + j}.png" for j in range(m)])
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
        names = os.listdir(out_dir)
        size = sum(os.path.getsize(os.path.join(out_dir, nm)) for nm in names) // max(len(names), 1)
        assert len(names) == n_img
        shutil.rmtree(out_dir, ignore_errors=True)
        return dt, size, wr.bytes_d2h // n_img

    host_workers = min(8, os.cpu_count() or 4)
    try:
        run_config5_png(16, "gpu", host_workers)
        runs = [run_config5_png(mine, "gpu", host_workers) for _ in range(3)]          # wall clock incl. file I/O: three runs
        ms_all = sorted(r[0] for r in runs)
        ms_g, size_g, d2h_g = ms_all[0], runs[0][1], runs[0][2]
        n_pil = min(mine, 96)
        ms_p, size_p, d2h_p = run_config5_png(n_pil, "pil", host_workers)
    except (RuntimeError, OSError) as e:      # a full or read-only scratch directory must not cost the whole bench line
        if world > 1:                         # (one rank skipping a stage with collectives would hang the others: fail loudly)
            raise
        out["config5_pseudo_labels_whole_set_to_png_files"] = {"skipped": f"{type(e).__name__}: {e}"[:200]}
        del pool5
        return out
    out["config5_pseudo_labels_whole_set_to_png_files"] = {
        "images": n_set, "images_per_rank": mine, "ms": ms_g, "images_per_s": n_set / (ms_g * 1e-3),
        "px_per_s": n_set * px5 / (ms_g * 1e-3), "unit": "px/s (all ranks)", "host_threads": host_workers,
        "host_logical_cores": os.cpu_count(), "ms_runs_sorted": ms_all, "ms_median": ms_all[1],
        "timing": "wall clock (host-bound stage with file I/O; min of 3 runs is `ms`, median beside it) — not a CUDA-event figure",
        "file_bytes_per_image": size_g, "d2h_bytes_per_image": d2h_g, "files_on": file_root or tempfile.gettempdir(),
        "pillow_encoder": {"images_per_rank": n_pil, "ms_per_image": ms_p / n_pil, "images_per_s": world * n_pil / (ms_p * 1e-3),
                           "file_bytes_per_image": size_p, "d2h_bytes_per_image": d2h_p},
        "speedup_vs_pillow_encoder": (ms_p / n_pil) / (ms_g / mine),
        "note": "config 5 per image + the PNG file: zlib stream made on the GPU (csrc/png.cu: Up filter, run tokens, fixed "
                "Huffman, IDAT CRC-32), 8 images per label call and 8 maps per encoder call, host threads only frame and write; wall clock incl. file I/O (tmpfs when available).  The synthetic "
                "label maps (arg-max of up-sampled random logits) are noise-like, the worst case for a run-length encoder; "
                "pillow_encoder = the same pipeline with the reference's Image.save on the host threads"}
    del pool5
    return out


# --------------------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------------------
_RESULT_OUT = None


def claim_stdout():
    """Rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL's version banner goes to fd 1 whenever
    NCCL_DEBUG is VERSION or above), so the real stdout is set aside for the result line and fd 1 is pointed at stderr for
    everything else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="diga_b200", choices=["diga_b200", "reference"])
    ap.add_argument("--no-stages", action="store_true", help="skip the secondary stage measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-baseline-json", action="store_true", help="(internal) run the CPU arm and print its JSON object")
    args = ap.parse_args()
    if args.cpu_baseline_json:
        emit_line(cpu_baseline())
        return
    if args.impl == "reference":
        run_reference(args)
        return

    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import diga_b200 as D
    from diga_b200 import _lib as L, synthetic as S

    peak, peak_src = load_peaks()
    K, W = args.steps, max(args.warmup, 3)
    n2, _, hh, ww = KD_SHAPE
    px_step = n2 * hh * ww
    g = S.gen(1234 + rank, dev)
    # three rotating input sets (each 638 MB, far beyond the 126 MB L2) so no step re-reads cached lines
    sets = [(S.logits(KD_SHAPE, g), S.logits(KD_SHAPE, g)) for _ in range(3)]
    up = torch.tensor(UPSTREAM, device=dev)

    def step(i, ev=None):
        t, s = sets[i % len(sets)]
        s = s.detach().requires_grad_(True)
        if ev:
            ev[0].record()
        loss = D.distillation_loss(t, s, 0.5)
        if ev:
            ev[1].record()
        (grad,) = torch.autograd.grad(loss, s, grad_outputs=up)
        if ev:
            ev[2].record()
        return loss, grad

    for i in range(W):
        step(i)
    torch.cuda.synchronize()

    # ---- timed region 1 (eager): the public-API step, and the two kernels of the step launched back to back through the C
    # ABI with a CUDA event pair around each launch (kernel time, not wrapper time) -----------------------------------------
    K_eager = min(K, 50)
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K_eager):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    eager_ms_step = max_over_ranks(e0.elapsed_time(e1), world, dev) / K_eager

    kd_ws = L.kd_workspace(dev)
    loss_buf = torch.empty((), dtype=torch.float32, device=dev)
    ds_buf = torch.empty(KD_SHAPE, dtype=torch.float32, device=dev)
    st = L.stream()
    hw_px = hh * ww

    def kd_fwd_launch(i):
        t, s = sets[i % len(sets)]
        L.check(L.lib.diga_kd_fwd(t.data_ptr(), s.data_ptr(), n2, C, hw_px, 0.5, loss_buf.data_ptr(), kd_ws.data_ptr(), st))

    def kd_bwd_launch(i):
        t, s = sets[i % len(sets)]
        L.check(L.lib.diga_kd_bwd(t.data_ptr(), s.data_ptr(), n2, C, hw_px, 0.5, up.data_ptr(), ds_buf.data_ptr(), st))

    def kd_fused_launch(i):
        t, s = sets[i % len(sets)]
        L.check(L.lib.diga_kd_fwd_bwd(t.data_ptr(), s.data_ptr(), n2, C, hw_px, 0.5, UPSTREAM, loss_buf.data_ptr(),
                                      ds_buf.data_ptr(), kd_ws.data_ptr(), st))

    (fwd_ms, _), (bwd_ms, bwd_seen), (fused_ms, _) = (launch_pairs_ms(f, K_eager) for f in (kd_fwd_launch, kd_bwd_launch, kd_fused_launch))
    del ds_buf

    # ---- timed region 2 (headline): the same API calls captured once per input set in CUDA graphs and replayed ----
    # The eager Python call chain (autograd.Function, allocator, ctypes) costs more host time per step than the two
    # kernels take on the GPU, so the eager loop is launch-bound; a graph replays exactly the same launches.
    graphs = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(len(sets)):
            step(i)                                    # warm the private pool on the capture stream
        torch.cuda.synchronize()
        for i in range(len(sets)):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=side):
                out = step(i)
            graphs.append((gph, out))
    torch.cuda.current_stream().wait_stream(side)
    for i in range(W):
        graphs[i % len(graphs)][0].replay()
    torch.cuda.synchronize()
    launches0 = L.launch_count()
    barrier(world)
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            graphs[i % len(graphs)][0].replay()
        e1.record()
        torch.cuda.synchronize()
    barrier(world)
    launches = 2 * K            # each replay launches kd_kernel<LOSS> and kd_kernel<GRAD> (captured once, not re-counted)
    ms_total = max_over_ranks(e0.elapsed_time(e1), world, dev)
    ms_step = ms_total / K
    value = world * px_step / (ms_step * 1e-3)
    # parity of the replayed result with the eager call on the same inputs
    ref_loss, ref_grad = step(0)
    assert torch.equal(ref_loss, graphs[0][1][0]) and torch.equal(ref_grad, graphs[0][1][1]), "graph replay != eager"
    bwd_gbs = px_step * KD_BYTES_BWD / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = px_step * KD_BYTES_FWD / (fwd_ms * 1e-3) / 1e9
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic_db = json.load(f)
    traffic = traffic_db.get("kd_bwd_dram_bytes_per_launch")

    # ---- end to end: host (pinned) inputs -> H2D -> distillation_loss + backward -> D2H loss -----------------
    e2e = None
    if not args.no_e2e:
        ke = max(3, min(K, 30))
        affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        numa = bind_to_gpu_numa_node(local_rank)           # pinned buffers are first-touched on the GPU's own NUMA node
        host_t = torch.empty(KD_SHAPE, dtype=torch.float32).pin_memory()
        host_s = torch.empty(KD_SHAPE, dtype=torch.float32).pin_memory()
        host_t.copy_(sets[0][0])
        host_s.copy_(sets[0][1])
        host_loss = torch.empty((ke + 3,), dtype=torch.float32).pin_memory()
        bufs = [(torch.empty(KD_SHAPE, device=dev), torch.empty(KD_SHAPE, device=dev)) for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        main_stream = torch.cuda.current_stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(nsteps):
            for f in freed:
                f.record(main_stream)
            for i in range(nsteps):
                j = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[j])               # buffer j no longer read by step i-2
                    bufs[j][0].copy_(host_t, non_blocking=True)
                    bufs[j][1].copy_(host_s, non_blocking=True)
                    ready[j].record(copy_stream)
                main_stream.wait_event(ready[j])
                s = bufs[j][1].detach().requires_grad_(True)
                loss = D.distillation_loss(bufs[j][0], s, 0.5)
                (grad,) = torch.autograd.grad(loss, s, grad_outputs=up)
                host_loss[i].copy_(loss.detach(), non_blocking=True)   # the step's result, read back to the host
                freed[j].record(main_stream)

        e2e_run(3)
        torch.cuda.synchronize()
        barrier(world)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        e2e_run(ke)
        a1.record()
        torch.cuda.synchronize()
        barrier(world)
        my_ms = a0.elapsed_time(a1) / ke
        e2e_ms = max_over_ranks(a0.elapsed_time(a1), world, dev) / ke
        h2d_bytes = 2 * host_t.numel() * 4
        per_rank = [h2d_bytes / (my_ms * 1e-3) / 1e9]
        if world > 1:
            gathered = [None] * world
            torch.distributed.all_gather_object(gathered, (per_rank[0], numa))
            per_rank, numas = [x[0] for x in gathered], [x[1] for x in gathered]
        else:
            numas = [numa]
        # the copy alone (no kernels): what the host side can feed this rank while every other rank copies too
        barrier(world)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            bufs[0][0].copy_(host_t, non_blocking=True)
            bufs[0][1].copy_(host_s, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        barrier(world)
        copy_gbs = 5 * h2d_bytes / (max_over_ranks(c0.elapsed_time(c1), world, dev) * 1e-3) / 1e9
        e2e = {"value": world * px_step / (e2e_ms * 1e-3), "unit": "pixel-positions/s",
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
               "steps": ke, "h2d_gbs_per_rank": [round(x, 2) for x in per_rank], "h2d_copy_only_gbs_slowest_rank": round(copy_gbs, 2),
               "numa": numas, "host_logical_cores": os.cpu_count(),
               "note": "pinned host logits -> H2D (double-buffered on a copy stream) -> distillation_loss + autograd.grad -> "
                       "D2H loss; gradient stays on the device for the backbone backward.  PCIe-bound: 637.5 MB per step and "
                       "rank; with N ranks the box's host memory / PCIe root complexes feed N copies at once "
                       "(h2d_copy_only_gbs_slowest_rank is the copy without any kernel)"}
        del host_t, host_s, bufs

        # Same loss through the patched call site of INTEGRATION.md (distillation_loss_upsampled on the stride-8 logits the
        # backbone emits, before nn.Upsample): host logits [8,19,65,129] in, loss AND the low-resolution gradient out.
        # Secondary figure: the headline e2e above stays the unchanged-signature call.
        lo_shape = (KD_SHAPE[0], KD_SHAPE[1], 65, 129)
        g2 = S.gen(4321 + rank)
        h_lt, h_ls = S.logits(lo_shape, g2).pin_memory(), S.logits(lo_shape, g2).pin_memory()
        h_grad = torch.empty(lo_shape, dtype=torch.float32).pin_memory()
        ke2 = 200
        h_loss2 = torch.empty((ke2,), dtype=torch.float32).pin_memory()
        d_lt, d_ls = torch.empty(lo_shape, device=dev), torch.empty(lo_shape, device=dev)

        def e2e_low(nsteps):
            for i in range(nsteps):
                d_lt.copy_(h_lt, non_blocking=True)
                d_ls.copy_(h_ls, non_blocking=True)
                x = d_ls.detach().requires_grad_(True)
                loss = D.distillation_loss_upsampled(d_lt, x, KD_SHAPE[2:], 0.5)
                (gr,) = torch.autograd.grad(loss, x, grad_outputs=up)
                h_loss2[i].copy_(loss.detach(), non_blocking=True)
                h_grad.copy_(gr, non_blocking=True)

        e2e_low(5)
        torch.cuda.synchronize()
        barrier(world)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        e2e_low(ke2)
        b1.record()
        torch.cuda.synchronize()
        barrier(world)
        low_ms = max_over_ranks(b0.elapsed_time(b1), world, dev) / ke2
        e2e["patched_call_site"] = {
            "value": world * px_step / (low_ms * 1e-3), "unit": "pixel-positions/s", "ms_per_step": low_ms, "steps": ke2,
            "h2d_bytes_per_step": 2 * h_lt.numel() * 4, "d2h_bytes_per_step": 4 + h_grad.numel() * 4,
            "note": "distillation_loss_upsampled + autograd.grad on pinned host logits [8,19,65,129] (the tensors before "
                    "nn.Upsample, self_training.py:344,351): H2D of both maps, D2H of the loss and of the gradient, one stream, "
                    "eager; includes the up-sampling the reference arm does not time"}
        del h_lt, h_ls, h_grad, d_lt, d_ls
        if affinity0 is not None:
            os.sched_setaffinity(0, affinity0)             # the CPU arm (a child process) is meant to see every core again

    stages, accum_rows, c1_gpu = None, [], None
    if not args.no_stages:
        del sets
        torch.cuda.empty_cache()
        accum_rows = accum_kernel_rows(L, S, dev, peak)
        stages = bench_stages(D, S, dev, peak, world)
        if rank == 0:
            c1_gpu = config1_gpu_stages(D, S, dev)

    # CPU arm (rank 0, N=1 only): a process of its own — it makes `.cuda()` a no-op to run the reference's code on the host
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import subprocess
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-json"], capture_output=True, text=True,
                             timeout=900)
        try:
            cpu = json.loads(res.stdout.strip().splitlines()[-1])
        except Exception:                                   # noqa: BLE001
            cpu = {"value": None, "unit": "pixel-positions/s", "cores": None, "kind": "failed", "sample": (res.stderr or res.stdout)[-300:]}
        if c1_gpu and cpu.get("stages"):
            for rec in cpu["stages"]:                       # the matching GPU stage beside each CPU stage (same shapes)
                gpu = c1_gpu.get(rec["stage"])
                if gpu:
                    rec.update(gpu)
                    rec["gpu_over_cpu_all_threads"] = rec["cpu_ms_all_threads"] / gpu["gpu_ms"]

    if rank == 0:
        def row(stage, kernel, bound, unit, units, bytes_per_unit, ms, how, traffic_key=None, extra=None, issue_key=None):
            gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
            r = {"row": stage, "kernel": kernel, "bound": bound, "unit": unit, "units_per_launch": units,
                 "algo_bytes_per_unit": bytes_per_unit, "algo_bytes_per_launch": units * bytes_per_unit, "avg_launch_ms": ms,
                 "units_per_s_per_gpu": units / (ms * 1e-3), "achieved_gbs": gbs, "frac": gbs / peak,
                 "frac_of_8000_spec": gbs / 8000.0, "traffic": traffic_db.get(traffic_key) if traffic_key else None, "how": how}
            if bound == "alu":      # the HBM fraction of an ALU / latency-bound kernel says little: its issue-slot utilisation (ncu) beside it
                r["frac_note"] = "fraction of the HBM peak on the algorithmic bytes; the kernel is bound by issue slots / latency"
                r["ncu_issue_active_pct"] = traffic_db.get(issue_key) if issue_key else None
            if extra:
                r.update(extra)
            return r

        ev_how = "CUDA event pair per launch, C-ABI launches queued behind a spin-kernel gate, rotating inputs >> L2"
        gr_how = "public-API call replayed from a CUDA graph (CUDA events around 20 replays)"
        rows = [
            row("a1", "kd_kernel<LOSS> (forward)", "hbm", "pixel-position", px_step, KD_BYTES_FWD, fwd_ms, ev_how, "kd_fwd_dram_bytes_per_launch"),
            row("a1", "kd_kernel<GRAD> (backward)", "hbm", "pixel-position", px_step, KD_BYTES_BWD, bwd_ms, ev_how, "kd_bwd_dram_bytes_per_launch"),
            row("a1", "kd_kernel<LOSS,GRAD> (single pass)", "hbm", "pixel-position", px_step, KD_BYTES_BWD, fused_ms, ev_how),
        ] + [row(*r) for r in accum_rows]
        if stages:
            def from_stage(stage, key, kernel, bound, unit, traffic_key=None, extra=None, issue_key=None):
                st_ = stages.get(key)
                if st_ and "ms" in st_:
                    px = st_["px_per_s"] / world * st_["ms"] * 1e-3
                    rows.append(row(stage, kernel, bound, unit, px, st_["algo_bytes_per_px"], st_["ms"], gr_how, traffic_key, extra, issue_key))
            from_stage("a3", "pseudo_label_1scale", "pseudo_label_kernel (one scale)", "hbm", "px", "pl1_dram_bytes_per_launch")
            from_stage("a3", "pseudo_label_2scale", "pseudo_label_kernel (two scales, max-fused)", "hbm", "px", "pl2_dram_bytes_per_launch")
            from_stage("a3+f1", "pseudo_label_fused_upsample_labels_only", "pseudo_label_upsampled_kernel (labels from stride-8 logits)", "alu", "px", issue_key="plup_issue_active_pct")
            from_stage("a4", "consensus_select", "consensus_select_kernel", "alu", "px", "select_dram_bytes_per_launch", issue_key="select_issue_active_pct")
            from_stage("a2", "classmix_dacs_blend_kernel", "classmix_blend_kernel (DACS)", "hbm", "px", "cm_dram_bytes_per_launch")
            from_stage("a5", "proto_distance_softmax_d2048", "proto_umma_kernel (tcgen05 3xTF32)", "hbm", "feature-px", "proto_dram_bytes_per_launch")
            from_stage("a6+a7", "centroid_accumulate_update_d2048_iid_classes", "update_from_features call: assign + accum + finish (i.i.d. classes)", "hbm", "feature-px")
            from_stage("a6+a7", "centroid_accumulate_update_d2048_blocky_classes", "update_from_features call: assign + accum + finish (4x4 blocks)", "hbm", "feature-px")
            from_stage("f2", "cross_entropy2d_fwd_bwd", "ce_kernel fwd + bwd", "hbm", "px")
            from_stage("f4", "ema_teacher_update", "ema_update_kernel", "hbm", "parameter")
            from_stage("f1", "kd_fused_upsample_fwd_bwd", "loss_up_kernel<KD> loss+gradient from the stride-8 logits (distillation_loss_upsampled + autograd)", "alu", "pixel-position")
            from_stage("f1+f2", "cross_entropy2d_fused_upsample_fwd_bwd", "loss_up_kernel<CE> loss+gradient from the stride-8 logits", "alu", "px", issue_key="lossup_ce_issue_active_pct")
            from_stage("f1", "seg_plus_kd_fused_upsample_fwd_bwd", "loss_up_kernel<KD+CE> loss pass + gradient pass (self_training.py:348-352)", "alu", "pixel-position", issue_key="lossup_kdce_issue_active_pct")
            from_stage("f5", "confusion_matrix_eval", "confusion_matrix_kernel (runningScore.update)", "hbm", "px")
            from_stage("f3", "label_reader_resize_remap", "label_resize_remap_kernel (CityLoader NEAREST resize + id look-up)", "hbm", "px")
            for key, label in (("config3_self_training_step_dropin_functions", "config 3: self-training step hot path with ONLY the function names swapped (losses behind torch's nn.Upsample)"),
                               ("config3_self_training_step_dropin_functions_lazy_upsample", "config 3: same unchanged call sites, the scripts' three nn.Upsample modules built from diga_b200.nn.Upsample"),
                               ("config3_self_training_step_fused_call_sites", "config 3: self-training step hot path, B=8 @512x1024, D=2048 (patched call sites)"),
                               ("config4_calc_centroids_whole_set", "config 4: calc_centroids over 2975 images, sum mode (one all-reduce)"),
                               ("config4_calc_centroids_whole_set_exact", "config 4: calc_centroids over 2975 images, exact mode (all-gather + ordered replay)"),
                               ("config5_pseudo_labels_whole_set", "config 5: 2975 full-resolution pseudo-label maps with prototype rectification")):
                st_ = stages.get(key)
                if st_:
                    rows.append({"row": key, "kernel": label, "bound": "hbm", "ms": st_["ms"], "px_per_s_all_ranks": st_["px_per_s"],
                                 "images_per_s_all_ranks": st_.get("images_per_s"), "achieved_gbs": st_["gbs_per_gpu"],
                                 "frac": st_["frac_hbm"], "how": "CUDA events around the whole stage, max over ranks"})
        line = {
            "metric": METRIC, "value": value, "unit": "pixel-positions/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "notes": {"l2": "inputs 638 MB per step exceed the 126 MB L2; 3 rotating input sets",
                      "launch": "each input set's distillation_loss + autograd.grad call captured in a CUDA graph, replayed K "
                                "times (eager_ms_per_step reported beside it)",
                      "parallelism": f"image-sharded x{world}, no data-path collective on the headline step; the centroid "
                                     "stages (roofline.rows config4_*) carry the one exchange"},
            "roofline": {"bound": "hbm", "kernel": "kd_kernel<GRAD> (backward)", "achieved": bwd_gbs, "peak": peak,
                         "unit": "GB/s", "frac": bwd_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algo_bytes_per_launch": px_step * KD_BYTES_BWD, "avg_launch_ms": bwd_ms,
                         "frac_of_8000_spec": bwd_gbs / 8000.0, "launch_timing": bwd_seen,
                         "share_of_step": bwd_ms / ms_step, "rows": rows},
            "eager_ms_per_step": eager_ms_step,
            "gpu_launches": launches, "clocks": clocks.summary(), "e2e": e2e, "stages": stages, "cpu_baseline": cpu,
        }
        emit_line(line)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
