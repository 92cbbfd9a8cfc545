"""In-tree build of ``libb200pinn.so``: nvcc, sm_100a only, ``-lineinfo`` for ncu source pages."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["mlp_fwd_mc.cu", "mlp_tc.cu", "mlp_tc3.cu", "mlp_tc_bwd.cu", "mlp_wide_tc.cu", "mlp_wide_res.cu", "mlp_bwd.cu", "residuals.cu", "export_rf.cu", "gmm.cu", "adam_misc.cu"]
OUT = os.path.join(HERE, "libb200pinn.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200pinn.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = OUT, objdir: str | None = None) -> str:
    if not force and not _stale() and out == OUT:
        return OUT
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, *(["-Xptxas=-v"] if verbose else []), "-c", "-o", obj, os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"b200pinn: nvcc failed on {src}")
        if verbose:
            print(res.stderr)
    link = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out,
                           *[obj for _, obj, _ in results]], capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("b200pinn: link failed")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
