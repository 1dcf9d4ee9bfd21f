import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diga_b200 as D
from diga_b200 import _lib as L, synthetic as S
dev = torch.device("cuda", 0)
g = S.gen(17, dev)
n, d, h, w, c = 8, 2048, 65, 129, 19
feat = S.features((n, d, h, w), g)
cen = S.centroids(c, d, g)
ws = torch.empty(int(L.lib.diga_proto_workspace_bytes(c, d)), dtype=torch.uint8, device=dev)
wt = torch.empty((n, c, h, w), device=dev)
def call():
    L.check(L.lib.diga_proto_distance(feat.data_ptr(), cen.data_ptr(), n, d, c, h * w, None, wt.data_ptr(), ws.data_ptr(), L.stream()))
for dbg in [int(a) for a in sys.argv[1:]] or (0, 1, 2, 3, 4, 7, 16, 23):
    L.set_tunable("umma_debug", dbg)
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    print(f"debug={dbg}: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
# in-kernel profile of CTA 0
off = (d // 8) * 2048 + 256
names = ["loaderA", "loaderB", "mma", "conv_w0", "epi_w0"]
for extra in (0, 1, 2, 4, 7):
    L.set_tunable("umma_debug", 32 + extra)
    call(); torch.cuda.synchronize()
    prof = ws[off: off + 256 * 64 * 8].view(torch.int64).reshape(256, 8, 8).cpu()
    print(f"--- profile, debug bits {extra}")
    for cta in (0,):
        for r, nm in enumerate(names):
            print(f"cta {cta} {nm:8s} total {prof[cta, r, 7].item():8d} cyc  waits {[int(v) for v in prof[cta, r, :4]]}")
L.set_tunable("umma_debug", 0)
