"""nvcc recipe for libia_b200.so (sm_100a only, built in-tree so it travels with gpurun snapshots).

    python -m instant_angelo_b200.build [--force]

Each .cu is compiled to an object with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC
(march.cu additionally with -fmad=false: its outputs are bit-exact contracts) and linked into
instant_angelo_b200/_build/libia_b200.so with a static cudart.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.environ.get("IA_BUILD_DIR") or os.path.join(HERE, "_build")     # A/B builds: IA_BUILD_DIR + IA_NVCC_EXTRA, loaded with IA_LIB_PATH
LIB = os.path.join(BUILD, "libia_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["capi.cu", "hashgrid.cu", "sh.cu", "march.cu", "composite.cu", "mlp.cu", "mlp_fp32.cu", "mlp_tc.cu", "mlp_tc2.cu", "mlp_tc3.cu", "adam.cu", "sdf_taps.cu", "linear64.cu", "weightnorm.cu", "colour_in.cu", "losses.cu"]
EXTRA = {"march.cu": ["-fmad=false"]}
if os.environ.get("IA_TC_TIMING"):     # cycle accounting of the tensor-core MLP kernels (tools/prof_mlp.py --timing); off in product builds
    EXTRA["mlp_tc.cu"] = ["-DIA_TC_TIMING=1"]
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"] + \
    os.environ.get("IA_NVCC_EXTRA", "").split()


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ia_b200.h"), __file__]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    deps = _deps()
    headers = [d for d in deps if d.endswith((".h", ".cuh", ".py"))]

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [NVCC, "-c", os.path.join(CSRC, src), "-o", obj] + COMMON + EXTRA.get(src, [])
            if verbose:
                cmd += ["-Xptxas", "-v"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
