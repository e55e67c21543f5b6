"""Glue between the engine shim (quantum_sim.engine.*) and libqsb: program cache, single-state runs.

Everything numeric goes through `capi.Context`; nothing here computes amplitudes on the host.
"""

from __future__ import annotations

import threading
from collections import OrderedDict

import numpy as np

from . import capi
from .compiler import Lowering

_tls = threading.local()
_CACHE_MAX = 256


def ctx():
    return capi.get_context()


def _cache():
    c = getattr(_tls, "programs", None)
    if c is None:
        c = _tls.programs = OrderedDict()
    return c


def cached_program(key, build):
    """Device program for `key` (hashable) on this thread's context; `build()` -> compiler.Program."""
    c = _cache()
    k = (capi.default_device(), key)
    dp = c.get(k)
    if dp is None:
        dp = ctx().program(build())
        c[k] = dp
        if len(c) > _CACHE_MAX:
            c.popitem(last=False)
    else:
        c.move_to_end(k)
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
