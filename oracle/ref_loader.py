"""Import the *real* reference functions: from ``$DIGA_REF``, ``/root/reference`` (build container) or the git-ignored
staging copy ``baseline/_ref/`` that ``oracle/make_ref.py`` writes and ``gpurun`` ships to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/diga_oracle.py``).  Reference sources are never tracked by this repo and
never imported by ``diga_b200/``.  Users:

* ``tests/golden/make_golden.py`` — generate the committed golden fixtures (build container);
* ``tests/test_oracle_golden.py`` — pin ``diga_oracle`` against the live reference (skipped when no tree is found);
* ``bench.py --impl reference`` / ``cpu_baseline`` — time the reference's own CPU implementation (``kind: "reference"``).

Shims: the reference imports ``kornia`` and ``matplotlib`` at module
top (unused by the hot path) and hard-codes ``.cuda()``; on a CUDA-less host that call is made a
no-op.  One reference tree per process (``G/`` and ``S/`` both define ``util.*``).
"""
from __future__ import annotations

import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
TREES = {
    "G": "domain_adaptation/GTA5",
    "S": "domain_adaptation/Synthia",
}




def _find_root() -> str:
    cands = [os.environ.get("DIGA_REF"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, TREES["G"], "calc_centroids.py")):
            return c
    return cands[1]


REF_ROOT = _find_root()


def available(tree: str = "G") -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, TREES[tree], "calc_centroids.py"))


def load(tree: str = "G") -> types.SimpleNamespace:
    """Returns a namespace with ``distillation_loss``, ``cross_entropy2d``, ``process_label``, ``update_teacher_params``,
    ``Class_Features``, ``runningScore``."""
    if not available(tree):
        raise FileNotFoundError(f"reference tree {tree} not found under {REF_ROOT}")
    loaded = getattr(load, "_tree", None)
    if loaded is not None and loaded != tree:
        raise RuntimeError(f"reference tree {loaded} already imported in this process; start a new one for {tree}")
    for name in ("kornia", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    path = os.path.join(REF_ROOT, TREES[tree])
    if path not in sys.path:
        sys.path.insert(0, path)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from util.loss import distillation_loss, cross_entropy2d, OhemCrossEntropy          # noqa: E402
        from util.utils import process_label, update_teacher_params       # noqa: E402
        from calc_centroids import Class_Features        # noqa: E402
        from util.metrics import runningScore            # noqa: E402
    load._tree = tree
    return types.SimpleNamespace(distillation_loss=distillation_loss, process_label=process_label,
                                 Class_Features=Class_Features, cross_entropy2d=cross_entropy2d,
                                 update_teacher_params=update_teacher_params, runningScore=runningScore, OhemCrossEntropy=OhemCrossEntropy, tree=tree)
