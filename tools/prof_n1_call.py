"""The single-image a6 -> a7 chain (assign, accumulate, finish) a few times, for `ncu --metrics gpu__time_duration.sum`:
the reference loops over the target set one image per call (calc_centroids.py:67-78)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diga_b200 as D
from diga_b200 import synthetic as S
dev = torch.device("cuda", 0)
g = S.gen(17, dev)
feat = S.features((1, 2048, 65, 129), g); out = S.logits((1, 19, 65, 129), g)
cf = D.Class_Features(19, 2048)
for _ in range(6):
    cf.update_from_features(feat, out, None, "mean")
torch.cuda.synchronize()
