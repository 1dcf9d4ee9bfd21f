#!/usr/bin/env python
"""Times the fused up-sampling losses (csrc/loss_up.cu) at config-2/3 sizes with CUDA events, next to the materialised
pipeline they replace (nn.Upsample -> diga kernels -> torch up-sampling backward) and the all-torch reference chain.
One JSON line per measurement, appended to gpurun_out/losses_up.jsonl."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import diga_b200 as D  # noqa: E402
from diga_b200 import _lib as L, synthetic as S  # noqa: E402
from oracle import diga_oracle as O  # noqa: E402  (timed as the eager GPU reference chain, never as product)

dev = torch.device("cuda", 0)
out_path = os.path.join(ROOT, "gpurun_out", "losses_up.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, "a")


def timeit(fn, iters=30, warm=5, graph=True):
    """ms per call; sync-free calls are replayed from a CUDA graph so the Python call chain is not what is timed."""
    for _ in range(warm):
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
                keep = fn()                          # noqa: F841
        torch.cuda.current_stream().wait_stream(side)
        fn = gph.replay
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, px, **kw):
    rec = {"name": name, "ms": round(ms, 4), "gpx_per_s": round(px / ms / 1e6, 2), **kw}
    print(json.dumps(rec), flush=True)
    fout.write(json.dumps(rec) + "\n")
    fout.flush()


def main():
    g = S.gen(5, dev)
    n2, c, lo, hi = 8, 19, (65, 129), (512, 1024)
    px = n2 * hi[0] * hi[1]
    tea, stu = S.logits((n2, c, *lo), g), S.logits((n2, c, *lo), g)
    tgt = S.block_labels(n2 // 2, hi[0], hi[1], g)
    up = torch.tensor(0.25, device=dev)
    one = torch.tensor(1.0, device=dev)
    rys = [int(v) for v in os.environ.get("RYS", "32").split(",")]
    for ry in rys:
        L.set_tunable("lossup_ry", ry)
        tag = {"ry": ry}

        def kd_fwd():
            return D.distillation_loss_upsampled(tea, stu, hi, 0.5)

        def kd_fwd_bwd():
            s = stu.detach().requires_grad_(True)
            loss = D.distillation_loss_upsampled(tea, s, hi, 0.5)
            return torch.autograd.grad(loss, s, grad_outputs=up)

        def kd_single():
            return D.distillation_loss_upsampled_and_grad(tea, stu, hi, 0.5, 0.25)

        def ce_fwd_bwd():
            s = stu[:4].detach().requires_grad_(True)
            loss = D.cross_entropy2d_upsampled(s, tgt)
            return torch.autograd.grad(loss, s, grad_outputs=one)

        def both_fwd_bwd():
            s = stu.detach().requires_grad_(True)
            l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s, tgt, 0.5)
            return torch.autograd.grad([l_ce, l_kd], s, grad_outputs=[one, up])

        report("kd_up_fwd", timeit(kd_fwd), px, **tag)
        report("kd_up_fwd_bwd", timeit(kd_fwd_bwd), px, **tag)
        report("kd_up_single_pass", timeit(kd_single), px, **tag)
        report("ce_up_fwd_bwd", timeit(ce_fwd_bwd), px // 2, **tag)
        report("seg_plus_kd_up_fwd_bwd", timeit(both_fwd_bwd), px, **tag)

    # what it replaces, (a) diga's materialised kernels behind torch's up-sampling, (b) the all-torch reference chain
    def materialised():
        s = stu.detach().requires_grad_(True)
        loss = D.distillation_loss(O.upsample_bilinear_ac(tea, hi), O.upsample_bilinear_ac(s, hi), 0.5)
        return torch.autograd.grad(loss, s, grad_outputs=up)

    def eager():
        s = stu.detach().requires_grad_(True)
        loss = O.distillation_loss_upsampled(tea, s, hi, 0.5)
        return torch.autograd.grad(loss, s, grad_outputs=up)

    def materialised_both():
        s = stu.detach().requires_grad_(True)
        ups = O.upsample_bilinear_ac(s, hi)
        l_kd = D.distillation_loss(O.upsample_bilinear_ac(tea, hi), ups, 0.5)
        l_ce = D.cross_entropy2d(O.upsample_bilinear_ac(s[:4], hi), tgt)
        return torch.autograd.grad([l_ce, l_kd], s, grad_outputs=[one, up])

    ohem_d, ohem_t = D.OhemCrossEntropy(255, 0.7, 100000), O.OhemCrossEntropyOracle(255, 0.7, 100000)

    def ohem_diga():
        s = stu[:4].detach().requires_grad_(True)
        loss = ohem_d(s, tgt)
        return torch.autograd.grad(loss, s, grad_outputs=one)

    def ohem_torch():
        s = stu[:4].detach().requires_grad_(True)
        loss = ohem_t(s, tgt)
        return torch.autograd.grad(loss, s, grad_outputs=one)

    report("ohem_up_fwd_bwd", timeit(ohem_diga), px // 2)
    report("ohem_all_torch_eager_chain", timeit(ohem_torch, iters=10, warm=2, graph=False), px // 2)
    report("kd_materialised_torch_upsample_plus_diga_kd", timeit(materialised), px)
    report("kd_all_torch_eager_chain", timeit(eager, iters=10, warm=2), px)
    report("seg_plus_kd_materialised", timeit(materialised_both), px)


if __name__ == "__main__":
    main()
