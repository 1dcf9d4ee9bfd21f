"""Stage the importable part of the reference under the git-ignored ``baseline/_ref/`` (SURVEY.md §8c "shipping the oracle").

TEST / BENCH INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container but not on the GPU box; ``gpurun`` ships
``baseline/_ref/`` (git-ignored, not gpurun-ignored) with the repo snapshot, so that on the box

* ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's OWN ``distillation_loss``,
  ``Class_Features`` ... (``cpu_baseline.kind == "reference"``) instead of the oracle port, and
* ``tests/test_oracle_golden.py`` can pin the oracle against the live reference there too.

Only the Python modules that ``calc_centroids.py`` / ``util/loss.py`` / ``util/utils.py`` / ``util/metrics.py`` import are
copied (``*.py`` under ``util/`` and ``model/`` of the GTA5 and Synthia trees, no dataset lists); nothing of it is tracked by
git or imported by ``diga_b200/``.  Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("DIGA_REF_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
TREES = ("domain_adaptation/GTA5", "domain_adaptation/Synthia")
TOP_FILES = ("calc_centroids.py",)
PACKAGES = ("util", "model")


def stage(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"make_ref: {SRC} not present, nothing staged (existing {DST} left as is)")
        return False
    n = 0
    for tree in TREES:
        for name in TOP_FILES:
            src = os.path.join(SRC, tree, name)
            if os.path.isfile(src):
                os.makedirs(os.path.join(DST, tree), exist_ok=True)
                shutil.copyfile(src, os.path.join(DST, tree, name))
                n += 1
        for pkg in PACKAGES:
            for dirpath, _, files in os.walk(os.path.join(SRC, tree, pkg)):
                for f in files:
                    if not f.endswith(".py"):
                        continue
                    rel = os.path.relpath(os.path.join(dirpath, f), SRC)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(dirpath, f), os.path.join(DST, rel))
                    n += 1
    if verbose:
        print(f"make_ref: staged {n} reference modules under {DST} (git-ignored)")
    return n > 0


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
