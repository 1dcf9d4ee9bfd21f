import os, sys
sys.path.insert(0, "/root/repo")
import torch
import diga_b200 as D
from diga_b200 import _lib as L, synthetic as S
dev = torch.device("cuda", 0)
g = S.gen(17, dev)
for (n, d, h, w) in ((8, 2048, 65, 129), (1, 2048, 129, 257), (2, 2048, 65, 129), (8, 256, 65, 129)):
    c = 19
    feat = S.features((n, d, h, w), g); cen = S.centroids(c, d, g)
    ws = torch.empty(int(L.lib.diga_proto_workspace_bytes(c, d)), dtype=torch.uint8, device=dev)
    wt = torch.empty((n, c, h, w), device=dev)
    def call():
        L.check(L.lib.diga_proto_distance(feat.data_ptr(), cen.data_ptr(), n, d, c, h * w, None, wt.data_ptr(), ws.data_ptr(), L.stream()))
    for fit in (0, 1, 0, 1):
        L.set_tunable("umma_fit_tiles", fit)
        call(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/20
        print((n,d,h,w), "fit", fit, f"{ms*1e3:.1f} us  {feat.numel()*4/ms/1e6:.0f} GB/s", flush=True)
