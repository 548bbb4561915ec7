"""In-tree nvcc build of libnasr_ctc.so for sm_100a (no torch headers, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libnasr_ctc.so")
SOURCES = ["nasr_api.cu", "ctc_loss.cu", "ctc_fast.cu", "ctc_narrow.cu", "ctc_decode.cu", "ctc_beam.cu"]
HEADERS = [os.path.join(CSRC, "nasr_common.cuh"), os.path.join(ROOT, "include", "nasr_ctc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # the shared CUDA runtime: the process has one runtime (torch's libcudart.so.12 when torch is imported first)
    # and the library does not embed a private copy of every runtime entry point
    "--cudart=shared",
    "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libnasr_ctc.so")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into lib/libnasr_ctc.so (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    # NASR_TUNING=1 in the environment compiles the tuning hooks in (per-role / per-phase cycle counters behind
    # nasr_debug_profile, the ablation bits of nasr_debug_config): tools/gpu_fast_check.py roles, tools/gpu_beam_check.py
    # phases need them; production builds leave them out (cfg3 loss+grad 0.332 -> 0.285 ms without them).
    tuning = ["-DNASR_TUNING=1"] if os.environ.get("NASR_TUNING") == "1" else []
    cmd = [_nvcc()] + NVCC_FLAGS + tuning + [
        "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", LIB_PATH,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper that nvcc must not pick up as host compiler
    env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
