"""Run one hot kernel a few times (for ncu captures):  python tools/prof_one.py proto|accum|select|plup"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diga_b200 as D
from diga_b200 import _lib as L, synthetic as S
dev = torch.device("cuda", 0)
g = S.gen(17, dev)
which = sys.argv[1]
n, d, h, w, c = 8, 2048, 65, 129, 19
if which in ("proto", "accum"):
    feat = S.features((n, d, h, w), g)
    cf = D.Class_Features(c, d); cf.objective_vectors = S.centroids(c, d, g)
    out = S.logits((n, c, h, w), g)
    for _ in range(3):
        if which == "proto":
            cf.get_centroid_weight(feat)
        else:
            cf.update_from_features(feat, out, None, "mean")
elif which == "select":
    wl = torch.softmax(S.logits((8, 19, 65, 129), g), 1)
    tl = S.block_labels(8, 512, 1024, g)
    for _ in range(3):
        D.consensus_select(tl, wl)
elif which == "pl":
    z, z2 = S.logits((4, 19, 1024, 2048), g), S.logits((4, 19, 1024, 2048), g)
    for _ in range(3):
        D.pseudo_label(z)
        D.pseudo_label(z, z2)
elif which == "cm":
    import random
    sl = S.block_labels(8, 512, 1024, g); tl = S.perturb_labels(sl, g)
    xa, xb = S.images((8, 3, 512, 1024), g), S.images((8, 3, 512, 1024), g)
    for _ in range(3):
        D.classmix(sl, xa, xb, tl, rng=random.Random(1), return_mask=False)
elif which == "plup":
    l1, l2 = S.logits((4, 19, 129, 257), g), S.logits((4, 19, 65, 129), g)
    for _ in range(3):
        D.pseudo_label_two_scale(l1, l2, (1024, 2048))
elif which == "ce":
    x = S.logits((4, 19, 512, 1024), g); t = S.block_labels(4, 512, 1024, g); one = torch.tensor(1.0, device=dev)
    for _ in range(3):
        xx = x.detach().requires_grad_(True)
        torch.autograd.grad(D.cross_entropy2d(xx, t), xx, grad_outputs=one)
elif which == "ema":
    from diga_b200.util.utils import ema_update_tensors
    sizes = [1024 * 256, 256 * 256 * 9, 256 * 1024, 1024] * 23 + [2048 * 512, 512 * 512 * 9, 512 * 2048] * 3
    tea = [torch.randn(sz, device=dev) for sz in sizes]; stu = [torch.randn(sz, device=dev) for sz in sizes]
    for _ in range(3):
        ema_update_tensors(tea, stu, 0.999)
elif which == "eval":
    from diga_b200.util.metrics import runningScore
    from diga_b200.util.labels import resize_remap_labels, trainid_lut
    gt = S.block_labels(4, 1024, 2048, g)
    pred, _ = D.pseudo_label_two_scale(S.logits((4, 19, 129, 257), g), S.logits((4, 19, 65, 129), g), (1024, 2048), want_conf=False)
    raw = torch.randint(0, 34, (8, 1024, 2048), device=dev, dtype=torch.uint8)
    rs = runningScore(19)
    for _ in range(3):
        rs.update(gt, pred)
        resize_remap_labels(raw, (512, 1024), trainid_lut())
elif which == "png":
    from diga_b200.pseudolabel import png_deflate
    lab = S.block_labels(8, 1024, 2048, g, 32).to(torch.uint8)
    for _ in range(3):
        png_deflate(lab)
elif which == "presence":
    from diga_b200.classmix import present_classes
    sl = S.block_labels(8, 512, 1024, g)
    for _ in range(3):
        present_classes(sl)
elif which == "lossup3":                     # config 3's two loss launches: seg + KD on [16,...] (8 supervised), CE on [8,...]
    tea, stu, cpm = S.logits((16, 19, 65, 129), g), S.logits((16, 19, 65, 129), g), S.logits((8, 19, 65, 129), g)
    sl, ml = S.block_labels(8, 512, 1024, g), S.block_labels(8, 512, 1024, g)
    for _ in range(3):
        s, c2 = stu.detach().requires_grad_(True), cpm.detach().requires_grad_(True)
        part, l_src, l_kd = D.seg_distillation_total_upsampled(tea, s, sl, 1.0, 0.25, 0.5)
        total = part + D.cross_entropy2d_upsampled(c2, ml)
        torch.autograd.grad(total, [s, c2])
    torch.cuda.synchronize()
    if len(sys.argv) > 2:                    # event timing of the pair (outside ncu)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            s, c2 = stu.detach().requires_grad_(True), cpm.detach().requires_grad_(True)
            part, l_src, l_kd = D.seg_distillation_total_upsampled(tea, s, sl, 1.0, 0.25, 0.5)
            total = part + D.cross_entropy2d_upsampled(c2, ml)
            torch.autograd.grad(total, [s, c2])
        e1.record(); torch.cuda.synchronize()
        print("lossup3 pair ms", e0.elapsed_time(e1) / 20, float(total))
elif which == "lossup":
    tea, stu = S.logits((8, 19, 65, 129), g), S.logits((8, 19, 65, 129), g)
    tgt = S.block_labels(4, 512, 1024, g)
    one, up = torch.tensor(1.0, device=dev), torch.tensor(0.25, device=dev)
    for _ in range(2):
        s = stu.detach().requires_grad_(True)
        l_kd = D.distillation_loss_upsampled(tea, s, (512, 1024), 0.5)                 # loss kernel (KD)
        torch.autograd.grad(l_kd, s, grad_outputs=up)                                  # grad kernel (KD) + gather
        s = stu.detach().requires_grad_(True)
        l_ce, l_kd = D.seg_distillation_losses_upsampled(tea, s, tgt, 0.5)             # loss kernel (KD+CE)
        torch.autograd.grad([l_ce, l_kd], s, grad_outputs=[one, up])                   # grad kernel (KD+CE) + gather
torch.cuda.synchronize()
