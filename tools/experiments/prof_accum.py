import os, sys
sys.path.insert(0, "/root/repo")
import torch
import diga_b200 as D
from diga_b200 import _lib as L, synthetic as S
dev = torch.device("cuda", 0)
g = S.gen(3, dev)
n, d, h, w = 8, 2048, 65, 129
feat = S.features((n, d, h, w), g)
cls = torch.randint(0, 19, (n, h * w), device=dev, dtype=torch.uint8)
sums = torch.empty((n, 19, d), device=dev)
ws = L.accum_workspace(n, d, 19, h * w, dev)
for _ in range(3):
    L.check(L.lib.diga_centroid_accum_ws(feat.data_ptr(), cls.data_ptr(), n, d, 19, h * w, sums.data_ptr(), ws.data_ptr(), L.stream()))
torch.cuda.synchronize()
