import sys; sys.path.insert(0,'/root/repo')
import torch, diga_b200 as D
from diga_b200 import synthetic as S
dev=torch.device('cuda',0); g=S.gen(5,dev)
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    side=torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize(); gph=torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=side): keep=fn()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3): gph.replay()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): gph.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters
for c in (19,10):
    for n2 in (8,16):
        tea,stu=S.logits((n2,c,65,129),g),S.logits((n2,c,65,129),g)
        t1=timeit(lambda: D.distillation_loss_upsampled_and_grad(tea,stu,(512,1024),0.5,0.25))
        with torch.no_grad():
            t2=timeit(lambda: D.distillation_loss_upsampled(tea,stu,(512,1024),0.5))
        print(f"C={c} n2={n2}: single-pass {t1*1e3:.1f} us, loss-only {t2*1e3:.1f} us")
