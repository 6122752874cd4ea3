"""Build the C++ host layer's drivers (hyperdeal_b200/cpp/*.cc) with g++ against libhdgpu.so.

The shim (cpp/hyperdeal_b200.hpp) is header-only; the two drivers re-host examples/advection and
performance/operators_advection_01 on it.  Binaries go to hyperdeal_b200/bin/ (git-ignored, shipped
to the GPU box with the snapshot); they find libhdgpu.so through an $ORIGIN-relative rpath.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CPP = os.path.join(HERE, "cpp")
BIN = os.path.join(HERE, "bin")
INCLUDE = os.path.join(HERE, "..", "include")
LIBDIR = os.path.join(HERE, "lib")
DRIVERS = ["advection", "operators_advection", "vlasov_poisson"]
HEADERS = ["hyperdeal_b200.hpp", "json_parameters.hpp"]


def build(force: bool = False) -> list[str]:
    os.makedirs(BIN, exist_ok=True)
    lib = os.path.join(LIBDIR, "libhdgpu.so")
    if not os.path.exists(lib):
        raise RuntimeError("build libhdgpu.so first (python -m hyperdeal_b200.build)")
    deps = [os.path.join(CPP, h) for h in HEADERS] + [os.path.join(INCLUDE, "hyperdeal_b200.h"), lib]
    out = []
    for d in DRIVERS:
        src, exe = os.path.join(CPP, d + ".cc"), os.path.join(BIN, d)
        if force or not os.path.exists(exe) or any(os.path.getmtime(x) > os.path.getmtime(exe) for x in [src] + deps):
            cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", INCLUDE, "-I", CPP, src, "-o", exe, "-L", LIBDIR, "-lhdgpu", "-Wl,-rpath,$ORIGIN/../lib"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("g++ failed for " + d)
            if res.stderr.strip():
                sys.stderr.write(res.stderr)
        out.append(exe)
    return out


if __name__ == "__main__":
    print("\n".join(build(force="--force" in sys.argv)))
