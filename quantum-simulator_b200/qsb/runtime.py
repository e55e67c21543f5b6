"""Glue between the engine shim (quantum_sim.engine.*) and libqsb: program cache, single-state runs.

Everything numeric goes through `capi.Context`; nothing here computes amplitudes on the host.
"""

from __future__ import annotations

import threading
from collections import OrderedDict

import numpy as np

from . import capi
from .compiler import Lowering

_CACHE_MAX = 256
_cache_lock = threading.Lock()
_programs = OrderedDict()       # (device, precision, key) -> DeviceProgram; process-wide like the contexts


def ctx(precision="c128"):
    return capi.get_context(precision=precision)


def cached_program(key, build, precision="c128"):
    """Device program for `key` (hashable) on the process-wide context of this device and precision;
    `build()` -> compiler.Program.  Evicted programs are freed when their last user drops them."""
    c = ctx(precision)
    k = (c.device, precision, key)
    with _cache_lock:
        dp = _programs.get(k)
        if dp is not None and dp.ctx is c and dp.handle is not None:
            _programs.move_to_end(k)
            return dp
    dp = c.program(build())
    with _cache_lock:
        _programs[k] = dp
        if len(_programs) > _CACHE_MAX:
            _programs.popitem(last=False)
    return dp


def matrix_key(mat):
    m = np.ascontiguousarray(mat, dtype=np.complex128)
    return (m.shape, m.tobytes())


def run_single(n, dev_state, dprog, *, uniforms=None, branches=False):
    """Apply a program in place to one device-resident state (load -> ops -> store)."""
    c = ctx()
    kw = {}
    if uniforms is not None and dprog.prog.n_draws:
        u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(1, -1)
        kw.update(uniforms=c.to_device(u), uniforms_stride=u.shape[1])
    c.run(dprog, 1, states=dev_state, load=True, store=True, **kw)
