"""Build the CUDA library (sm_100a) in-tree: ddm_b200/lib/libdddm_b200.so.

    python -m ddm_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to build/ (git-ignored); the .so stays in the
package directory so that it travels with the source snapshot and is the file the Python loader
opens (ddm_b200/_cabi.py) — there is no JIT cache and no fallback when it is missing.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "lib", "libdddm_b200.so")

SOURCES = ["api.cu", "elementwise.cu", "host_session.cu", "energy_reg_f32.cu", "energy_reg_bf16.cu",
           "energy_tile_f32.cu", "energy_tile_bf16.cu", "energy_smem_f32.cu", "energy_smem_bf16.cu", "energy_blk_f32.cu", "energy_blk_bf16.cu", "energy_wave_f32.cu", "energy_wave_bf16.cu", "energy_pipe_f32.cu", "energy_pipe_bf16.cu", "energy_tc.cu", "backbone_ops.cu", "metrics.cu", "metrics_tc.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps_mtime() -> float:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "dddm_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = _nvcc()
    hdr_time = _deps_mtime()

    def compile_one(src: str):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and os.path.exists(o) and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_time):
            return o, False
        cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return o, True

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if force or any(c for _, c in results) or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
