"""Config-3 step with the ClassMix classes fixed (no presence round trip), eager vs replayed from one CUDA graph: how much of
the 0.83 ms eager step is launch gaps rather than kernel time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diga_b200 as D
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import step_config3 as S3

x = S3.build()
hh, ww, h, w, b = S3.hh, S3.ww, S3.h, S3.w, S3.b
classes = [[0, 1, 2, 5, 7, 255]] * b


def step():
    cf, sl, tl, feat = x["cf"], x["sl"], x["tl"], x["feat"]
    wts = cf.get_centroid_weight(feat)
    kept, _ = D.consensus_select(tl, wts, (hh, ww))
    _, mix1 = D.classmix(sl, x["rec"], x["saug"], classes=classes, assume_labelled=True, return_mask=False)
    _, mix2, mixlabel = D.classmix(sl, x["tdata_aug"], x["sdata"], kept, classes=classes, assume_labelled=True, return_mask=False)
    cf.update_from_features(feat, x["t_pred"], start_mean=False, labels_full=kept)
    cf.update_from_features(x["s_feat"], x["s_pred"], start_mean=False, labels_full=sl)
    stu = x["stu_cat"].detach().requires_grad_(True)
    cpm = x["cross_low"].detach().requires_grad_(True)
    part, l_src, l_kd = D.seg_distillation_total_upsampled(x["tea_cat"], stu, sl, 1.0, 0.25, 0.5)
    total = part + D.cross_entropy2d_upsampled(cpm, mixlabel)
    g_stu, g_mix = torch.autograd.grad(total, [stu, cpm])
    return total, g_stu, g_mix, mix1, mix2


def timed(fn, n=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("eager, no presence round trip: %.3f ms" % timed(step))
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        keep = step()
torch.cuda.current_stream().wait_stream(side)
print("graph replay: %.3f ms" % timed(g.replay))
