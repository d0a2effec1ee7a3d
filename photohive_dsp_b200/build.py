"""Builds photohive_dsp_b200/PhotoHive_DSP_lib/libreport_data.so with nvcc for sm_100a (in-tree, so the
binary travels with a repo snapshot).  `python -m photohive_dsp_b200.build [--force] [--verbose]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "PhotoHive_DSP_lib")   # same relative location as the reference (lib.py:21)
LIB = os.path.join(OUT_DIR, "libreport_data.so")
OBJ_DIR = os.path.join(PKG, "build")
SOURCES = ["frontend.cu", "palette_select.cu", "fft.cu", "sharpness.cu", "finalize.cu", "f64path.cu", "pipeline.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libreport_data.so cannot be built (there is no CPU build of this library)")
    return exe


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "photohive_dsp.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    # g++ from the system: the image exports CXX=/opt/gcc/bin/g++, a wrapper nvcc does not need
    common = [nvcc(), "-ccbin", "/usr/bin/g++", *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]
    common += os.environ.get("PHD_NVCC_EXTRA", "").split()  # experiments: extra -D... / -Xptxas flags

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = common + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc(), "-ccbin", "/usr/bin/g++", *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
