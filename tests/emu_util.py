"""Test-only helper: run a lowered program on the host emulator (tests/emu/emu_exec.cpp).

The emulator shares qsb_exec.cuh with the CUDA build, so these runs check the op loop, the index
math and the host compiler on a machine without a GPU.  It is NOT a product path."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu_exec.cpp")
LIB = os.path.join(HERE, "emu", "libqsb_emu.so")
HDRS = [os.path.join(ROOT, "quantum-simulator_b200", "csrc", "qsb_exec.cuh"),
        os.path.join(ROOT, "include", "qsb.h")]

_lib = None


def build():
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-shared", "-fPIC", "-pthread",
                               "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(ROOT, "quantum-simulator_b200", "csrc"),
                               "-o", LIB, SRC])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_run.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def emu_run(prog, count=1, T=2, states=None, params=None, uniforms=None, seed=0, traj_offset=0,
            init_basis=None, default_basis=0, want_branches=False, store=True, accum_probs=False, out_of_place=False):
    """Returns dict(states, snapshots, branches, probs)."""
    dim = 1 << prog.n
    flags = 0
    if states is not None:
        flags |= 1
        states = np.ascontiguousarray(states, dtype=np.complex128).reshape(count, dim).copy()
    else:
        states = np.zeros((count, dim), dtype=np.complex128)
    if store:
        flags |= 2
    if prog.normalize:
        flags |= 4
    probs = None
    if accum_probs:
        flags |= 16
        probs = np.zeros(dim, dtype=np.float64)
    ops = np.ascontiguousarray(prog.ops)
    cdata = np.ascontiguousarray(prog.cdata, dtype=np.float64)
    idata = np.ascontiguousarray(prog.idata, dtype=np.int32)
    if params is not None:
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(count, -1)
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(count, -1)
    if init_basis is not None:
        init_basis = np.ascontiguousarray(init_basis, dtype=np.int64)
    branches = np.full((count, max(prog.n_draws, 1)), -1, dtype=np.int32) if want_branches else None
    snaps = np.zeros((count, max(prog.n_snapshots, 1), dim), dtype=np.complex128) if prog.n_snapshots else None
    out_states = np.zeros_like(states) if out_of_place else None
    rc = lib().emu_run(
        ctypes.c_int(prog.n), ctypes.c_int(prog.m), ctypes.c_int(8 if T < 3 else 16), _p(ops), ctypes.c_int64(len(ops)),
        ctypes.c_int64(prog.ops_stride), _p(cdata), ctypes.c_int64(len(cdata)), _p(idata), ctypes.c_int(prog.load_perm),
        ctypes.c_int(prog.store_perm), ctypes.c_int(prog.n_snapshots), ctypes.c_int(flags), _p(states),
        ctypes.c_int64(count), _p(params), ctypes.c_int64(params.shape[1] if params is not None else 0),
        _p(uniforms), ctypes.c_int64(uniforms.shape[1] if uniforms is not None else 0),
        ctypes.c_uint64(seed), ctypes.c_int64(traj_offset), _p(init_basis), ctypes.c_int64(default_basis),
        _p(branches), ctypes.c_int64(branches.shape[1] if branches is not None else 0), _p(snaps), _p(probs),
        _p(out_states))
    assert rc == 0, rc
    return dict(states=out_states if out_of_place else states, snapshots=snaps, branches=branches, probs=probs)
