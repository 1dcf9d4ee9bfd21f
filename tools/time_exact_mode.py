#!/usr/bin/env python
"""Config 4 in exact mode under torchrun: the pass with the rows exchanged by peer stores (symmetric memory, fused into the means
kernel) against the all-gather exchange.  Prints one JSON line from rank 0.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/time_exact_mode.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import diga_b200 as D
from diga_b200 import parallel as P, synthetic as S

g = S.gen(100 + rank, dev)
C, d, h, w, b, n_set = 19, 2048, 65, 129, 8, 2975
pool = [(S.features((b, d, h, w), g), S.logits((b, C, h, w), g)) for _ in range(4)]


def mx(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"world": world}
for symmetric, mcast in ((True, "1"), (True, "0"), (False, "0")):
    os.environ["DIGA_MULTICAST"] = mcast
    cf = D.Class_Features(C, d)
    sp = P.ShardedCentroidPass(cf, n_set, batch=b, symmetric=symmetric)

    def one_pass():
        for k in sp.my_batches():
            f, o = pool[k % 4]
            take = sp.batch_size_of(k)
            sp.add(f[:take], o[:take])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sp.finish()
        e1.record()
        return e0, e1

    one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times, tails = [], []
    for _ in range(3):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        e0, e1 = one_pass()
        a1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(mx(a0.elapsed_time(a1)))
        tails.append(mx(e0.elapsed_time(e1)))
    out[sp.exchange] = {"pass_ms": min(times), "finish_ms": min(tails), "symm_error": getattr(sp, "_symm_error", None)}
    if "ref" not in out:
        out["ref"] = True
        ref = cf.objective_vectors.clone()
    else:
        out["bit_equal_" + sp.exchange] = bool(torch.equal(ref, cf.objective_vectors))
    del sp
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
