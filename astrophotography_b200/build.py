"""Build ``libapgpu.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m astrophotography_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the authoring container;
the built ``.so`` is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD_DIR = os.path.join(PKG_DIR, "_build")
LIB_PATH = os.path.join(PKG_DIR, "libapgpu.so")

SOURCES = ['stack_sorted_med_f32_p3.cu',
           'stack_sorted_med_u16_p3.cu',
           'stack_sorted_medmad1_f32_p3.cu',
           'stack_sorted_medmad1_u16_p3.cu',
           'stack_sorted_medunc_f32_p3.cu',
           'stack_sorted_medunc_u16_p3.cu',
           'stack_sorted_med_f32_p2.cu',
           'stack_sorted_med_u16_p2.cu',
           'stack_sorted_medmad1_f32_p2.cu',
           'stack_sorted_medmad1_u16_p2.cu',
           'stack_sorted_medunc_f32_p2.cu',
           'stack_sorted_medunc_u16_p2.cu',
           'stack_sorted_med_f32_p1.cu',
           'stack_sorted_med_u16_p1.cu',
           'stack_sorted_medmad1_f32_p1.cu',
           'stack_sorted_medmad1_u16_p1.cu',
           'stack_sorted_medunc_f32_p1.cu',
           'stack_sorted_medunc_u16_p1.cu',
           'stack_sorted_med_f32_p0.cu',
           'stack_sorted_med_u16_p0.cu',
           'stack_sorted_medmad1_f32_p0.cu',
           'stack_sorted_medmad1_u16_p0.cu',
           'stack_sorted_medunc_f32_p0.cu',
           'stack_sorted_medunc_u16_p0.cu',
           'stack_median_coop_medunc_p8.cu',
           'stack_median_coop_medunc_p4.cu',
           'stack_median_coop_medmad1_p8.cu',
           'stack_median_coop_medmad1_p4.cu',
           'stack_median_coop_med_p8.cu',
           'stack_median_coop_med_p4.cu',
           'stack_meanclip_coop_p8.cu',
           'stack_meanclip_coop_p4.cu',
           'stack_meanclip_coop_p2.cu',
           'stack_meanclip_split_p8.cu',
           'stack_meanclip_mid.cu',
           'stack_meanclip_hi.cu',
           'stack_meanclip_lo.cu',
           'stack_meanclip_mid_u16.cu',
           'stack_meanclip_hi_u16.cu',
           'stack_meanclip_lo_u16.cu',
           'stack_meanclip_smem.cu',
           'stack_generic.cu',
           'stack.cu',
           'apgpu_core.cu',
           'calibrate.cu',
           'calibrate_repair.cu',
           'imarith.cu',
           'badpix.cu',
           'stats.cu']
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",                    # parity: numpy never fuses multiply-add; fmaf() where wanted
    "-Xcompiler", "-fPIC,-O2",
    "-Xptxas", "-v",
    "-DAPGPU_BUILDING", *(["-DAPGPU_DEBUG_MEDMAD"] if os.environ.get("APGPU_DEBUG") else []),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libapgpu.so")


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(" ".join(SOURCES).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _deps():
    deps = [os.path.join(INCLUDE, "apgpu.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".inc", ".h")):
            deps.append(os.path.join(CSRC, f))
    return deps


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link ``libapgpu.so``.

    Re-uses the existing library when no source changed.  Returns its path."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    stamp = os.path.join(BUILD_DIR, "stamp.txt")
    digest = _digest(_deps())
    if (not force and os.path.exists(LIB_PATH) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        return LIB_PATH
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD_DIR, src.replace(".cu", ".ptxas.log"))
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs,
           "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
