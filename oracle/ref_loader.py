"""Locate, stage and import the UNMODIFIED reference scripts.

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

``/root/reference`` exists only in the build container.  ``stage()`` (called by
``__graft_entry__.build()`` there, recipe: plain ``shutil.copy2`` of the five Apache-2.0 scripts
plus LICENSE) puts byte-identical copies under ``oracle/_ref/``, which is git-ignored (the
sources never enter this repository's history) but NOT gpurun-ignored, so the copies travel to
the GPU box with the snapshot.  There they give

* ``bench.py --impl reference``: the reference's own classes timed on the host cores
  (``cpu_baseline.kind = "reference"``) and on eager PyTorch-CUDA (``gpu_eager_baseline``);
* ``-m gpu`` tests that rebind the reference module's names with ``b200pinn.install`` and let the
  reference's own loops (01:948-955, 01:1413-1491, 01:1877-2010) drive the kernels.

``load("01")`` imports a script by path under a non-``__main__`` name with the two out-of-tree
shims of SURVEY.md 8c: a ``MagicMock`` matplotlib (not installed here; the import only runs the
font set-up, 01:55) and a ``StepLR`` wrapper that drops the ``verbose=`` keyword torch 2.11
removed (01:940).  No source line of the reference is edited.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"
SCRIPTS = {
    "01": "01_train_pinn_multiphysics_model.py",
    "02": "02_fault_classification_auc.py.py",
    "03": "03_unsupervised_gmm_fault_diagnosis.py.py",
    "04": "04_risk_function_early_warning_index.py.py",
    "05": "05_compare_fault_diagnosis_methods.py.py",
}
EXTRA = ["LICENSE"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(verbose: bool = False) -> bool:
    """Copy the reference scripts into ``oracle/_ref/`` (build container only).  Returns True if
    the staged copies exist afterwards."""
    if os.path.isdir(SOURCE):
        os.makedirs(STAGED, exist_ok=True)
        lines = []
        for name in list(SCRIPTS.values()) + EXTRA:
            src, dst = os.path.join(SOURCE, name), os.path.join(STAGED, name)
            if not os.path.exists(src):
                continue
            if not os.path.exists(dst) or _sha(src) != _sha(dst):
                shutil.copy2(src, dst)
            lines.append(f"{_sha(dst)}  {name}")
        with open(os.path.join(STAGED, "SHA256SUMS"), "w") as f:
            f.write("\n".join(lines) + "\n")
        if verbose:
            print(f"oracle/_ref: staged {len(lines)} files from {SOURCE}")
    return available()


def ref_dir():
    """Directory holding the reference scripts: the staged copy first (it is what travels), else the mount."""
    for d in (STAGED, SOURCE):
        if os.path.exists(os.path.join(d, SCRIPTS["01"])):
            return d
    return None


def available() -> bool:
    return ref_dir() is not None


_LOADED: dict = {}


def load(which: str = "01", device=None, fresh: bool = False):
    """Import reference script ``which`` ("01".."04") by path; ``device`` ("cpu"/"cuda") overrides the
    module-global ``device`` the script picks at import (01:21-24; it is looked up at call time)."""
    d = ref_dir()
    if d is None:
        raise FileNotFoundError("reference scripts not found: neither oracle/_ref/ (run __graft_entry__.build() in the "
                                "build container) nor /root/reference exists")
    key = (which, None if device is None else str(device))
    if fresh or key not in _LOADED:
        for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.lines", "matplotlib.patches",
                  "matplotlib.gridspec", "matplotlib.colors", "matplotlib.ticker", "matplotlib.cm"):
            sys.modules.setdefault(m, MagicMock())
        spec = importlib.util.spec_from_file_location(f"ref{which}", os.path.join(d, SCRIPTS[which]))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if which == "01":
            real = mod.StepLR

            def step_lr(opt, step_size, gamma=0.1, **kw):       # torch >= 2.4 dropped `verbose` (01:940 passes it)
                kw.pop("verbose", None)
                return real(opt, step_size=step_size, gamma=gamma, **kw)

            mod.StepLR = step_lr
        _LOADED[key] = mod
    mod = _LOADED[key]
    if device is not None and hasattr(mod, "device"):
        import torch

        mod.device = torch.device(device)
    return mod
