"""In-tree build of the CUDA C-ABI library (``csrc/libmatgcn.so``) with nvcc for sm_100a.

No torch headers are involved: the library's boundary is plain C (include/matgcn.h), so a
single ``nvcc -shared`` is the whole build.  The ``.so`` stays in-tree (git-ignored) so it
travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmatgcn.so")
# translation unit -> the headers it depends on (None = every .cuh in csrc/); each is compiled to its own object so
# that touching the small train_step.cu does not rebuild the 3-minute tensor-core unit
_MAIN_HDRS = ["gemm_simt.cuh", "epilogues.cuh", "gemm_tc.cuh", "res_bwd.cuh", "xside_mma.cuh", "rec_api.h"]
_REC_HDRS = ["gemm_simt.cuh", "epilogues.cuh", "gemm_tc.cuh", "rec_api.h", "rec_fwd.cuh", "rec_bwd.cuh", "dr_pass.cuh"]
SOURCES = {"matgcn.cu": _MAIN_HDRS, "rec.cu": _REC_HDRS, "train_step.cu": []}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
PUBLIC_HEADER = os.path.join(os.path.dirname(HERE), "include", "matgcn.h")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _obj(src: str) -> str:
    return os.path.join(CSRC, "_obj", os.path.splitext(src)[0] + ".o")


def _deps(src: str):
    hdrs = SOURCES[src]
    if hdrs is None:
        hdrs = [f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return [os.path.join(CSRC, src), PUBLIC_HEADER] + [os.path.join(CSRC, h) for h in hdrs]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def needs_build() -> bool:
    return any(_stale(_obj(s), _deps(s)) for s in SOURCES) or _stale(LIB, [_obj(s) for s in SOURCES])


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.join(CSRC, "_obj"), exist_ok=True)
    for src in SOURCES:
        if force or _stale(_obj(src), _deps(src)):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-c", "-o", _obj(src), os.path.join(CSRC, src)]
            if verbose:
                print("[matgcn build]", " ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    cmd = [_nvcc(), "-shared", "-o", LIB] + [_obj(s) for s in SOURCES]
    if verbose:
        print("[matgcn build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
