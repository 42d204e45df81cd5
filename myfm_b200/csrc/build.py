"""Builds libmyfm_b200.so (the C-ABI engine) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only container too; the resulting
shared object travels to the GPU box with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmyfm_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "field_sweep.cuh", "tile_sweep.cuh", "latent_device.cuh", "eval_device.cuh", "prep_device.cuh", "host_data.hpp", "rng.hpp", "mt_device.cuh", "mt_jump.hpp", "oprobit.cuh",
           "../../include/myfm_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # element-wise results must match the CPU path bit for bit: no FMA contraction
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
    "-ldl",  # NCCL is bound with dlopen at run time (row-sharded training only)
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS + ["build.py"])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           "-o", LIB, *[os.path.join(HERE, s) for s in SOURCES]]
    env = dict(os.environ)
    env.pop("CXX", None)  # let nvcc pick the system g++ as host compiler
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmyfm_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


PYBIND_SRC = os.path.join(HERE, "pybind_binding.cpp")


def pybind_path() -> str:
    import sysconfig

    return os.path.join(os.path.dirname(HERE), "_myfm_pybind" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pybind(force: bool = False) -> str:
    """The pybind11 binding over the C ABI (myfm_b200/_myfm_pybind*.so): plain g++, links libmyfm_b200.so."""
    import sysconfig

    import pybind11

    out = pybind_path()
    lib = build()
    if not force and os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(PYBIND_SRC),
                                                                          os.path.getmtime(lib)):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", f"-I{pybind11.get_include()}",
           f"-I{sysconfig.get_paths()['include']}", PYBIND_SRC, "-o", out, f"-L{HERE}", "-lmyfm_b200",
           "-Wl,-rpath,$ORIGIN/csrc"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the pybind11 binding")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
