"""Build libbrdae.so (hand-written CUDA for sm_100a) in-tree with nvcc.

`python -m dae_scipy_b200.build` or `dae_scipy_b200.build.build()`.  Every translation unit is
compiled with `-gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false` (no implicit
FMA contraction: the fma() calls in the kernels are the arithmetic contract)
and the objects are linked into `dae_scipy_b200/libbrdae.so`, which travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbrdae.so")

SOURCES = ["brdae_abi.cu", "dispatch.cu", "k1_pendulum.cu", "k1_robertson.cu", "k1_transistor.cu", "k1_extra.cu",
           "k2_npendulum.cu", "k4_chain.cu", "fp64_peak.cu", "layout.cu", "k1_user.cu", "k2_user.cu"]
USER_TU = "k1_user.cu"       # stub in the stock library, the user's functor in a plugin build
USER_BAND_TU = "k2_user.cu"  # the same for a band problem with more than 8 unknowns (K2)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_header_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force, hdr_mtime, log):
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_mtime)):
        return obj, ""
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build_variant(tag: str, defines: dict) -> str:
    """Tuning helper: build a second library `libbrdae_<tag>.so` with extra -D macros (objects in
    build/<tag>/); select it at run time with BRDAE_LIB=<path>."""
    objdir = os.path.join(OBJDIR, tag)
    os.makedirs(objdir, exist_ok=True)
    lib = os.path.join(HERE, f"libbrdae_{tag}.so")
    dflags = [f"-D{k}={v}" for k, v in defines.items()]

    def one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        r = subprocess.run([_nvcc(), *NVCC_FLAGS, *dflags, "-c", os.path.join(CSRC, src), "-o", obj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        return obj
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, SOURCES))
    subprocess.run([_nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                   check=True)
    return lib


def build_plugin(header: str, lib: str, band: bool = False) -> str:
    """Plugin build: compile k1_user.cu (band problems: k2_user.cu) WITH the user's problem header and link it, in place of
    the stub, with the stock objects into `lib` -- a complete libbrdae whose problem BRDAE_USER is that functor."""
    build()                                                  # the stock objects (no-op when up to date)
    USER_TU = USER_BAND_TU if band else globals()["USER_TU"]
    macro = "BRDAE_USER_BAND_HEADER" if band else "BRDAE_USER_HEADER"
    obj = os.path.splitext(lib)[0] + "." + USER_TU.replace(".cu", ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, f'-D{macro}="{os.path.abspath(header)}"', "-c",
           os.path.join(CSRC, USER_TU), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for the user problem {header}:\n{r.stdout}\n{r.stderr}")
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES if s != USER_TU] + [obj]
    r = subprocess.run([_nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    hdr = _newest_header_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, hdr, verbose), SOURCES))
    objs = [o for o, _ in results]
    ptxas_log = "\n".join(log for _, log in results if log)
    if ptxas_log:
        with open(os.path.join(OBJDIR, "ptxas.log"), "w") as f:
            f.write(ptxas_log)
        if verbose:
            print(ptxas_log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
