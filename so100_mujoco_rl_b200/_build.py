"""In-tree build of libso100_b200.so (nvcc, sm_100a only).  Called by __graft_entry__.build(); never at import time."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libso100_b200.so")
SOURCES = ["so100_b200.cu"]
HEADERS = ["so100_dyn.cuh", "so100_dyn_gen.cuh", "so100_ppo_kernels.cuh", "so100_tc.cuh", os.path.join("..", "..", "include", "so100_b200.h"),
           os.path.join("..", "..", "include", "so100_ppo.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libso100_b200.so cannot be built on this machine")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


STEP_KERNEL_SOURCES = ["so100_b200.cu", "so100_dyn.cuh", "so100_dyn_gen.cuh"]


def csrc_hash() -> str:
    """Identity of the env-step kernel's sources (compiled into the library as so100_build_id(); ncu-derived figures in
    profiles/step_kernel_profile.json are keyed by it so that bench.py can tell when they are stale)."""
    import hashlib
    h = hashlib.sha1()
    for f in STEP_KERNEL_SOURCES:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:12]


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, f'-DSO100_CSRC_HASH="{csrc_hash()}"', "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB_PATH
