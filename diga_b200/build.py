"""Build ``libdiga_b200.so`` in-tree with nvcc for sm_100a.

``python diga_b200/build.py`` (or ``__graft_entry__.build()``; run it by path — ``python -m diga_b200.build`` imports the
package first, which refuses to load when the library is missing or stale).  One nvcc process per ``.cu`` file in
parallel, then one link step.  Objects are rebuilt only when the source (or a header) is newer.
nvcc cross-compiles without a GPU, so this runs in the build container and the resulting ``.so``
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libdiga_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("DIGA_NVCC_EXTRA", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "diga_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(spath), _headers_mtime()):
        return obj
    cmd = [NVCC, *ARCH, *FLAGS, "-I", os.path.join(ROOT, "include"), "-c", spath, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, src[:-3] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {src}")
    if verbose:
        print(f"compiled {src}")
    return obj


def build(verbose: bool = True, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            if f.endswith(".o"):
                os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
        if verbose:
            print(f"linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
