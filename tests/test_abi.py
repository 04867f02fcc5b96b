"""The C-ABI library loads on a CPU-only machine and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    names = set()
    for h in ("so100_b200.h", "so100_ppo.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(so100_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_are_exported(native_lib):
    from so100_mujoco_rl_b200 import _native
    names = _header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(native_lib, n), f"libso100_b200.so does not export {n}"
    assert sorted(_native.EXPORTS) == names


def test_abi_version_and_dims(native_lib):
    assert native_lib.so100_abi_version() == 3
    assert native_lib.so100_obs_dim(1) == 15 and native_lib.so100_obs_dim(2) == 15 and native_lib.so100_obs_dim(5) == 8
    assert native_lib.so100_act_dim(5) == 6 and native_lib.so100_obs_dim(6) == 15
    assert native_lib.so100_obs_dim(3) < 0
    assert b"unknown task" in native_lib.so100_last_error()


def test_struct_layouts_match(native_lib, spec):
    """so100_create validates struct_size before touching CUDA: a layout mismatch gives ERR_ARG (-1), a correct layout
    on this GPU-less machine proceeds to the CUDA probe (-2) or succeeds on a GPU box."""
    from so100_mujoco_rl_b200.tasks import make_task_cfg
    m, c = spec.to_ctypes(), make_task_cfg(1, 8)
    h = ctypes.c_void_p()
    rc = native_lib.so100_create(ctypes.byref(m), ctypes.byref(c), 0, ctypes.byref(h))
    assert rc in (0, -2), native_lib.so100_last_error()
    if rc == 0:
        native_lib.so100_destroy(h)
    m.struct_size += 8
    assert native_lib.so100_create(ctypes.byref(m), ctypes.byref(c), 0, ctypes.byref(h)) == -1
    m.struct_size -= 8
    c.num_envs = 0
    assert native_lib.so100_create(ctypes.byref(m), ctypes.byref(c), 0, ctypes.byref(h)) == -1
    c.num_envs, c.task = 8, 3
    assert native_lib.so100_create(ctypes.byref(m), ctypes.byref(c), 0, ctypes.byref(h)) == -1


def test_null_arguments_are_rejected(native_lib):
    assert native_lib.so100_step(None, None, None, None, None, None, None, None, None, None) == -1
    assert native_lib.so100_reset(None, None, None, None) == -1
    assert native_lib.so100_get_tick(None, None) == -1


def test_oracle_and_product_share_struct_layout(spec):
    from oracle.pyoracle import lib
    from so100_mujoco_rl_b200.model import So100Model
    from so100_mujoco_rl_b200.tasks import So100TaskCfg
    assert lib().orc_sizeof_model() == ctypes.sizeof(So100Model)
    assert lib().orc_sizeof_task_cfg() == ctypes.sizeof(So100TaskCfg)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (only tests, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "so100_mujoco_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "so100_oracle" not in txt and "orc_" not in txt, f
