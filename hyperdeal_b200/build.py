"""Build libhdgpu.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The library is plain CUDA C++ with an `extern "C"` surface (include/hyperdeal_b200.h); it does
not link against torch.  `python -m hyperdeal_b200.build` or `__graft_entry__.build()`.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libhdgpu.so")
SOURCES = ["capi.cu", "kernels_generic.cu", "kernel_fast6d.cu", "kernel_vp.cu", "poisson_x.cu", "vp_diagnostics.cu", "kernel_tile_global.cu", "multi_gpu.cu"]
HEADERS = ["hd_internal.h", "basis.hpp", "rounds6d_tasks.cuh", "kernel_rounds6d.cuh", "kernel_vp_tile.cuh", os.path.join("..", "..", "include", "hyperdeal_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(LIBDIR, s.replace(".cu", ".ptxas.log"))
            with open(log, "w") as f:
                f.write(res.stderr)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed for " + s)
            if verbose:
                sys.stderr.write(res.stderr)
        objs.append(obj)
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB + ".tmp"] + objs + ["-lcuda"]
        # libcuda is only needed for cuTensorMapEncodeTiled; resolve it at run time through the
        # runtime's driver entry point instead of a link-time dependency
        cmd = cmd[:-1]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
        os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
