"""Build libnlo_b200.so in-tree with nvcc for sm_100a (and the small C test driver).

``python -m nlotrajectories_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting ``.so`` is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "_obj"
LIB = PKG / "libnlo_b200.so"
SOURCES = ["capi.cu", "sdf_simt.cu", "sdf_tc.cu", "sdf_tc256.cu", "nlp_kernels.cu", "nlp_hess.cu", "ip_solver.cu", "rrt_kernels.cu",
           "sdf_tc_hess.cu", "sdf_tc_deep.cu", "sdf_tc_deep_h128m2.cu", "sdf_tc_deep_h128m3.cu", "sdf_tc_deep_h64m2.cu", "sdf_tc_deep_h64m3.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libnlo_b200 cannot be built (there is no CPU fallback)")
    return exe


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + list((PKG.parent / "include").glob("*.h"))
    return max(h.stat().st_mtime for h in hdrs)


def _compile(src: str, verbose: bool) -> Path:
    s = CSRC / src
    o = OBJ / (Path(src).stem + ".o")
    if o.exists() and o.stat().st_mtime > max(s.stat().st_mtime, _deps_mtime()):
        return o
    cmd = [nvcc(), *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return o


def build_library(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if (not LIB.exists()) or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
