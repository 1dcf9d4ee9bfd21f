import sys; sys.path.insert(0, '/root/repo')
import torch, diga_b200 as D
from diga_b200 import synthetic as S
dev = torch.device('cuda', 0); g = S.gen(4321, dev)
C, d = 19, 2048
cf = D.Class_Features(C, d); cf.objective_vectors = S.centroids(C, d, g)
f5 = S.features((1, d, 129, 257), g)
l5a, l5b = S.logits((1, C, 129, 257), g), S.logits((1, C, 65, 129), g)
for _ in range(4):
    lab, _ = D.pseudo_label_two_scale(l5a, l5b, (1024, 2048), want_conf=False)
    wts = cf.get_centroid_weight(f5)
    kept = D.consensus_select(lab, wts, want_feat_pseudo=False)
torch.cuda.synchronize()
