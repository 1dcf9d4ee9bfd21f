"""Timing of the GPU PNG encoder (csrc/png.cu) and of the whole writer, gpu vs pil (results -> gpurun_out/exp_png.jsonl)."""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import diga_b200 as D  # noqa: E402
from diga_b200 import synthetic as S  # noqa: E402
from diga_b200.pseudolabel import PseudoLabelWriter, png_deflate  # noqa: E402

dev = torch.device("cuda", 0)
fout = open(os.path.join(ROOT, "gpurun_out", "exp_png.jsonl"), "a")


def emit(rec):
    print(json.dumps(rec), flush=True)
    fout.write(json.dumps(rec) + "\n")
    fout.flush()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = S.gen(5, dev)
C, d = 19, 2048
maps = {
    "blocks32": S.block_labels(8, 1024, 2048, g, 32).to(torch.uint8),
    "argmax_of_upsampled_random_logits": D.pseudo_label_two_scale(S.logits((8, C, 129, 257), g), S.logits((8, C, 65, 129), g),
                                                                  (1024, 2048), want_conf=False)[0],
}
for name, lab in maps.items():
    for n in (1, 8):
        ms = timeit(lambda: png_deflate(lab[:n]))
        _, lens = png_deflate(lab[:n])
        emit({"exp": "png_deflate", "labels": name, "batch": n, "us_per_image": round(ms * 1e3 / n, 2),
              "stream_bytes_per_image": int(lens.float().mean().item()), "gpx_per_s": round(n * 1024 * 2048 / ms / 1e6, 2)})

# whole writer: label maps already on the device, files on local disk
for name, lab in maps.items():
    for encoder, n_img, workers in (("gpu", 800, 4), ("gpu", 800, 8), ("pil", 160, 4), ("pil", 160, 16)):
        out = tempfile.mkdtemp(prefix="diga_png_")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with PseudoLabelWriter(out, workers=workers, slots=4, encoder=encoder) as wr:
            for k in range(n_img // 8):
                wr.submit(lab, [f"img_{k}_{j}.png" for j in range(8)])
        dt = time.perf_counter() - t0
        size = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out))
        emit({"exp": "writer", "labels": name, "encoder": encoder, "workers": workers, "images": n_img,
              "ms_per_image": round(dt * 1e3 / n_img, 3), "images_per_s": round(n_img / dt, 1), "file_bytes_per_image": size // n_img,
              "d2h_bytes_per_image": wr.bytes_d2h // n_img})
        shutil.rmtree(out)
