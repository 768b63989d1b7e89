"""In-tree build of the sm_100a shared libraries with nvcc.

    python robo-vln_b200/build.py [--force] [--verbose]

Two variants of the same sources are produced (csrc/h16.h):
    librobovln_b200.so       16-bit type = fp16 (default)
    librobovln_b200_bf16.so  16-bit type = bf16
They are written next to this file so that they travel to the GPU box with the repo snapshot
(git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SOURCES = ["gemm_tc.cu", "gemm_simt.cu", "elementwise.cu", "attention.cu", "attention_tc.cu", "lstm.cu", "prep.cu", "vla_block.cu", "train.cu", "engine.cu", "capi.cu"]
HEADERS = ["common.cuh", "h16.h", "rvb.h", "engine.h", os.path.join("..", "..", "include", "robovln_b200.h")]
VARIANTS = {"fp16": ("librobovln_b200.so", ["-DRVB_BF16=0"]), "bf16": ("librobovln_b200_bf16.so", ["-DRVB_BF16=1"])}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-cudart", "static",
]


def lib_path(variant: str = "fp16") -> str:
    return os.path.join(HERE, VARIANTS[variant][0])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


if os.environ.get("ROBOVLN_BUILD_STAMPS") == "1":        # phase-stamp instrumentation for tools/gemm_timeline.py (diagnostic builds only)
    NVCC_FLAGS.append("-DRVB_GEMM_STAMPS=1")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    libs = [lib_path(v) for v in VARIANTS]
    if not force and all(os.path.exists(p) for p in libs) and os.path.exists(stamp) and open(stamp).read() == dig:
        return libs
    nvcc = _nvcc()

    def compile_one(job):
        variant, src = job
        odir = os.path.join(OBJ, variant)
        os.makedirs(odir, exist_ok=True)
        obj = os.path.join(odir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *VARIANTS[variant][1], "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return variant, src, obj, r

    jobs = [(v, s) for v in VARIANTS for s in SOURCES]
    objs = {v: [] for v in VARIANTS}
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for variant, src, obj, r in ex.map(compile_one, jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src} [{variant}]")
            if verbose:
                sys.stderr.write(f"==== {src} [{variant}]\n{r.stderr}\n")
            with open(os.path.join(OBJ, variant, src + ".ptxas.log"), "w") as fh:
                fh.write(r.stderr)
            objs[variant].append(obj)
    for variant in VARIANTS:
        link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o",
                lib_path(variant), *objs[variant]]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"link failed [{variant}]")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return libs


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv, verbose="--verbose" in sys.argv):
        print(p)
