"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo).

  libkvc_b200.so   hand-written sm_100a CUDA kernels + the C ABI (include/kvc_b200.h)
  _C.so            torch dispatcher registration (csrc/torch_bindings.cpp) -> libkvc_b200.so

nvcc cross-compiles without a GPU.  Rebuilds only what is stale.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libkvc_b200.so")
EXT = os.path.join(PKG, "_C.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU_SOURCES = ["kvc_capi.cu", "kvc_schedule.cu", "kvc_moves.cu", "kvc_cache.cu", "kvc_attention.cu", "kvc_select.cu", "kvc_blocks.cu", "kvc_prefill.cu", "kvc_alloc.cu"]
# KVC_EXTRA_NVCC_FLAGS: extra compile flags for diagnostic builds (e.g. -DKVC_ATTN_TRACE for tools/attn_trace.py)
EXTRA = os.environ.get("KVC_EXTRA_NVCC_FLAGS", "").split()
NVCC_FLAGS = EXTRA + ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I" + INC, "-I" + CSRC]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_lib(verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "kvc_common.cuh"), os.path.join(INC, "kvc_b200.h")]
    jobs = []
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if _stale(o, [s] + headers):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC] + NVCC_FLAGS + extra + ["-c", s, "-o", o])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out:
                    print(out)
    if _stale(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-cudart", "static"])
    return LIB


def build_torch_ext() -> str:
    src = os.path.join(CSRC, "torch_bindings.cpp")
    if not _stale(EXT, [src, os.path.join(INC, "kvc_b200.h"), LIB]):
        return EXT
    import torch
    from torch.utils import cpp_extension as ce
    incs = ["-I" + p for p in ce.include_paths()] + ["-I" + sysconfig.get_paths()["include"],
                                                    "-I/usr/local/cuda/include", "-I" + INC]
    tlib = ce.library_paths()[0]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    obj = os.path.join(OBJ, "torch_bindings.o")
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi,
          "-DTORCH_EXTENSION_NAME=_C", "-w"] + incs + ["-c", src, "-o", obj])
    _run(["g++", "-shared", "-o", EXT, obj, "-L" + PKG, "-lkvc_b200", "-L" + tlib, "-ltorch", "-ltorch_cpu",
          "-ltorch_cuda", "-lc10", "-lc10_cuda", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib])
    return EXT


def build_all(verbose: bool = False):
    build_lib(verbose=verbose)
    build_torch_ext()
    return LIB, EXT


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))
