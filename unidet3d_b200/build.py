"""Build libunidet3d_b200.so (sm_100a) in-tree with nvcc.  No torch headers are involved: the
library is a plain C-ABI shared object (include/unidet3d_b200.h) loaded through ctypes."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libunidet3d_b200.so")
SOURCES = ["grid.cu", "gemm.cu", "gemm_ts.cu", "encoder.cu", "attention_tc.cu", "post.cu", "criterion.cu", "train.cu", "unet_plan.cu", "augment.cu", "eval.cu", "encoder_plan.cu", "attention_bwd.cu", "criterion_grad.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "gemm_common.cuh"), os.path.join(CSRC, "boxes.cuh"), os.path.join(CSRC, "box_loss.cuh"), os.path.join(HERE, "..", "include", "unidet3d_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libunidet3d_b200.so")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + HEADERS):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("UD3D_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
