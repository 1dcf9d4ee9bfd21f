"""Config 5 per image at 1, 2, 4 and 8 images per call (labels from the two-scale logits, prototype weights of
[n,2048,129,257], consensus selection on the uint8 maps): CUDA-event time per image over 240 images."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diga_b200 as D
from diga_b200 import synthetic as S
dev = torch.device("cuda", 0)
g = S.gen(3, dev)
cf = D.Class_Features(19, 2048)
cf.objective_vectors = S.centroids(19, 2048, g)
for n in (1, 2, 4, 8):
    pool = [(S.features((n, 2048, 129, 257), g), S.logits((n, 19, 129, 257), g), S.logits((n, 19, 65, 129), g)) for _ in range(3)]

    def run(calls):
        for k in range(calls):
            f, la, lb = pool[k % 3]
            lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
            D.consensus_select(lab, cf.get_centroid_weight(f), want_feat_pseudo=False)

    run(6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    calls = 240 // n
    e0.record(); run(calls); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"images_per_call": n, "us_per_image": round(e0.elapsed_time(e1) * 1e3 / (calls * n), 2)}))
    del pool
