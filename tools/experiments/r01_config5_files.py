"""Repeatability of the config-5-to-files stage: same loop as bench.py, 3 repeats each into /tmp and /dev/shm."""
import json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import diga_b200 as D
from diga_b200 import synthetic as S
from diga_b200.pseudolabel import PseudoLabelWriter
dev = torch.device("cuda", 0)
g = S.gen(4321, dev)
C, d = 19, 2048
cf = D.Class_Features(C, d); cf.objective_vectors = S.centroids(C, d, g)
pool5 = [(S.features((1, d, 129, 257), g), S.logits((1, C, 129, 257), g), S.logits((1, C, 65, 129), g)) for _ in range(4)]

def run(n_img, base, workers, coalesce, write=True):
    out_dir = tempfile.mkdtemp(prefix="diga_pl_", dir=base)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with PseudoLabelWriter(out_dir, workers=workers, slots=4, encoder="gpu", coalesce=coalesce) as wr:
        for k in range(n_img):
            f, la, lb = pool5[k % 4]
            lab, _ = D.pseudo_label_two_scale(la, lb, (1024, 2048), want_conf=False)
            kept, _ = D.consensus_select(lab, cf.get_centroid_weight(f), want_feat_pseudo=False)
            if write:
                wr.submit(kept, [f"img_{k}.png"])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    shutil.rmtree(out_dir, ignore_errors=True)
    return dt * 1e3

run(32, "/tmp", 8, 8)
for base in ("/tmp", "/dev/shm"):
    for workers, coalesce in ((8, 8), (4, 8), (8, 16), (12, 8)):
        ms = [round(run(2975, base, workers, coalesce), 1) for _ in range(3)]
        print(json.dumps({"dir": base, "workers": workers, "coalesce": coalesce, "ms": ms}), flush=True)
print(json.dumps({"no_write_loop_ms": [round(run(2975, "/tmp", 8, 8, write=False), 1) for _ in range(2)]}))
