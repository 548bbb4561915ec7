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
SOURCES = ["nasr_api.cu", "ctc_loss.cu", "ctc_fast.cu", "ctc_narrow.cu", "ctc_decode.cu", "ctc_beam.cu", "affine.cu", "affine_tc.cu", "affine_tc_dh.cu", "affine_tc_dw.cu"]
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
    """Compile csrc/*.cu into lib/libnasr_ctc.so (cross-compiles without a GPU).  One object per source, built in
    parallel and reused while the source and the headers are older than it; then one link."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    # NASR_TUNING=1 in the environment compiles the tuning hooks in (per-role / per-phase cycle counters behind
    # nasr_debug_profile, the ablation bits of nasr_debug_config): tools/gpu_fast_check.py roles, tools/gpu_beam_check.py
    # phases need them; production builds leave them out (cfg3 loss+grad 0.332 -> 0.285 ms without them).
    tuning = ["-DNASR_TUNING=1"] if os.environ.get("NASR_TUNING") == "1" else []
    tag = "_t" if tuning else ""
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper that nvcc must not pick up as host compiler
    env.pop("CXX", None)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared", "--cudart=shared")
                     and not f.startswith("-rpath") and f != "-Xlinker"]
    hdr_t = max(os.path.getmtime(h) for h in HEADERS)
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s.replace(".cu", tag + ".o"))
        objs.append(obj)
        fresh = os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t)
        if fresh and not (force and os.environ.get("NASR_REBUILD_ALL") == "1"):
            continue
        cmd = [_nvcc()] + compile_flags + tuning + (["-Xptxas", "-v"] if verbose else []) + [
            "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        jobs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)))
    for s, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
        if verbose:
            print(out)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "--cudart=shared",
           "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-ldl", "-o", LIB_PATH] + objs
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
