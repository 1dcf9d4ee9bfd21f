"""Run the accumulation kernel alone for an ncu capture:  python tools/prof_accum.py <pattern> <variant> [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diga_b200 import _lib as L, synthetic as S
from tools.sweep_accum import class_logits
dev = torch.device("cuda", 0)
g = S.gen(11, dev)
pattern, variant = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d, h, w, c = 2048, 65, 129, 19
hw = h * w
feats = [S.features((n, d, h, w), g) for _ in range(2)]
out = class_logits(pattern, n, c, h, w, g)
cls = torch.empty((n, hw), dtype=torch.uint8, device=dev)
clsw = torch.empty((int(L.lib.diga_centroid_clsw_bytes(n, hw)) // 4,), dtype=torch.int32, device=dev)
counts = torch.empty((n, c), dtype=torch.int32, device=dev)
sums = torch.zeros((n, c, d), dtype=torch.float32, device=dev)
L.check(L.lib.diga_centroid_assign(out.data_ptr(), None, n, c, hw, cls.data_ptr(), counts.data_ptr(), clsw.data_ptr(), L.stream()))
L.set_tunable("accum_variant", variant)
for k in range(4):
    L.check(L.lib.diga_centroid_accum(feats[k % 2].data_ptr(), cls.data_ptr(), counts.data_ptr(), clsw.data_ptr(), n, d, c, hw, sums.data_ptr(), L.stream()))
torch.cuda.synchronize()
