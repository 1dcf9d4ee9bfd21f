"""One full-resolution image of config 5 (labels from the two-scale logits, prototype weights of [1,2048,129,257], consensus
selection on the uint8 map), a few times, for an ncu capture of the single-image launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diga_b200 as D
from diga_b200 import synthetic as S
dev = torch.device("cuda", 0)
g = S.gen(3, dev)
cf = D.Class_Features(19, 2048)
cf.objective_vectors = S.centroids(19, 2048, g)
pool = [(S.features((1, 2048, 129, 257), g), S.logits((1, 19, 129, 257), g), S.logits((1, 19, 65, 129), g)) for _ in range(3)]
for k in range(6):
    f, la, lb = pool[k % 3]
    lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
    kept, _ = D.consensus_select(lab, cf.get_centroid_weight(f), want_feat_pseudo=False)
torch.cuda.synchronize()
