"""The drop-in boundary: libqsb.so loads, exports every entry point include/qsb.h declares, the ctypes binding
declares exactly those, and the product path fails loudly (no CPU fallback) without a device or without the library.
No compute call is made here: this file runs on the CPU-only build container."""
import ctypes as C
import os
import re

import pytest

from qsb import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qsb.h")


def declared_functions():
    text = open(HEADER, encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)            # comments may mention function names
    names = re.findall(r"^\s*(?:const\s+)?(?:int|int64_t|void\s*\*|char\s*\*|const char\s*\*)\s*\*?\s*(qsb_[a-z0-9_]+)\s*\(",
                       text, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_boundary():
    names = declared_functions()
    assert len(names) >= 40, names
    for must in ("qsb_ctx_create", "qsb_run", "qsb_program_create", "qsb_probabilities", "qsb_overlap",
                 "qsb_masked_parity", "qsb_rdm_all", "qsb_rho_accumulate", "qsb_readout_transform", "qsb_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "build it first: python __graft_entry__.py build"
    lib = C.CDLL(capi.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_binding_covers_the_header_exactly():
    assert sorted(capi.SYMBOLS) == declared_functions()
    lib = capi.load_library()
    for name in capi.SYMBOLS:
        fn = getattr(lib, name)
        assert fn.argtypes is not None, name                     # every entry point has a declared prototype


def test_no_cpu_fallback_without_a_device():
    lib = capi.load_library()
    assert isinstance(lib.qsb_version(), int) and lib.qsb_version() > 0
    if lib.qsb_device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    handle = C.c_void_p()
    rc = lib.qsb_ctx_create(0, C.byref(handle))
    assert rc < 0 and not handle.value
    msg = lib.qsb_last_error(None).decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    with pytest.raises(RuntimeError):
        capi.Context(0)
    # the engine API surfaces the same error instead of computing on the host
    from quantum_sim.engine.state_vector import StateVector
    with pytest.raises(RuntimeError):
        StateVector(3).probabilities


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", os.path.join(ROOT, "quantum-simulator_b200", "no_such_libqsb.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load_library()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/ or the emulator."""
    pkg = os.path.join(ROOT, "quantum-simulator_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|qsim_oracle|emu_exec|emu_run", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
