"""ctypes binding of libqsb.so (include/qsb.h) over NumPy buffers.

There is no CPU fallback: if the library is missing, or no CUDA device is visible, the first use
raises RuntimeError.  Errors from the library map to ValueError (QSB_E_INVAL, the reference's own
error class for bad arguments), NotImplementedError (QSB_E_UNSUPPORTED) or RuntimeError.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from .compiler import OP_DTYPE, Program

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libqsb.so")

RUN_LOAD, RUN_STORE, RUN_NORMALIZE, RUN_ASYNC, RUN_ACCUM_PROBS, RUN_LOAD_BROADCAST = 1, 2, 4, 8, 16, 32

SYMBOLS = [
    "qsb_version", "qsb_device_count", "qsb_ctx_create", "qsb_ctx_destroy", "qsb_ctx_set_stream",
    "qsb_ctx_sync", "qsb_ctx_set_precision", "qsb_ctx_trim", "qsb_buffer_upload_async",
    "qsb_event_create", "qsb_event_record", "qsb_event_wait", "qsb_event_free", "qsb_apply_dense", "qsb_last_error", "qsb_ctx_info", "qsb_timer_start", "qsb_timer_stop",
    "qsb_launch_count", "qsb_buffer_alloc", "qsb_buffer_wrap", "qsb_buffer_free", "qsb_buffer_upload",
    "qsb_buffer_download", "qsb_buffer_zero", "qsb_buffer_copy", "qsb_buffer_ptr", "qsb_buffer_bytes",
    "qsb_host_alloc", "qsb_host_free", "qsb_program_create", "qsb_program_free", "qsb_run",
    "qsb_debug_profile",
    "qsb_probabilities", "qsb_probabilities_sum", "qsb_sample_index", "qsb_overlap",
    "qsb_masked_parity", "qsb_rdm_all", "qsb_rdm_general", "qsb_mi_all_pairs", "qsb_rho_accumulate", "qsb_readout_transform",
    "qsb_stream_create", "qsb_stream_run", "qsb_stream_run_peers", "qsb_stream_run_scatter", "qsb_stream_free",
]


class RunArgs(C.Structure):
    _fields_ = [
        ("states", C.c_void_p), ("first", C.c_int64), ("count", C.c_int64),
        ("params", C.c_void_p), ("params_stride", C.c_int64),
        ("uniforms", C.c_void_p), ("uniforms_stride", C.c_int64),
        ("philox_seed", C.c_uint64), ("traj_offset", C.c_int64),
        ("init_basis", C.c_void_p), ("default_basis", C.c_int64),
        ("branches", C.c_void_p), ("branches_stride", C.c_int64),
        ("snapshots", C.c_void_p), ("probs_accum", C.c_void_p),
        ("flags", C.c_int32), ("reserved", C.c_int32),
        ("states_out", C.c_void_p), ("out_first", C.c_int64),
        ("peer_table", C.c_void_p), ("peer_shift", C.c_int32), ("reserved2", C.c_int32), ("peer_rank_or", C.c_int64),
    ]


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libqsb.so and declare the prototypes.  Raises RuntimeError when it is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  This backend has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
        P = C.POINTER
        proto = {
            "qsb_version": (C.c_int, []),
            "qsb_device_count": (C.c_int, []),
            "qsb_ctx_create": (C.c_int, [C.c_int, P(vp)]),
            "qsb_ctx_destroy": (C.c_int, [vp]),
            "qsb_ctx_set_stream": (C.c_int, [vp, vp]),
            "qsb_ctx_sync": (C.c_int, [vp]),
            "qsb_ctx_set_precision": (C.c_int, [vp, C.c_int]),
            "qsb_ctx_trim": (C.c_int, [vp]),
            "qsb_buffer_upload_async": (C.c_int, [vp, i64, vp, i64]),
            "qsb_event_create": (C.c_int, [vp, P(vp)]),
            "qsb_event_record": (C.c_int, [vp]),
            "qsb_event_wait": (C.c_int, [vp]),
            "qsb_event_free": (C.c_int, [vp]),
            "qsb_apply_dense": (C.c_int, [vp, i32, vp, i64, i64, vp, i64, i32, P(i32), vp, P(i32)]),
            "qsb_last_error": (C.c_char_p, [vp]),
            "qsb_ctx_info": (C.c_int, [vp, P(i32), P(i32), P(i32), P(i64)]),
            "qsb_timer_start": (C.c_int, [vp]),
            "qsb_timer_stop": (C.c_int, [vp, P(C.c_float)]),
            "qsb_launch_count": (i64, [vp]),
            "qsb_buffer_alloc": (C.c_int, [vp, i64, P(vp)]),
            "qsb_buffer_wrap": (C.c_int, [vp, vp, i64, P(vp)]),
            "qsb_buffer_free": (C.c_int, [vp]),
            "qsb_buffer_upload": (C.c_int, [vp, i64, vp, i64]),
            "qsb_buffer_download": (C.c_int, [vp, i64, vp, i64]),
            "qsb_buffer_zero": (C.c_int, [vp, i64, i64]),
            "qsb_buffer_copy": (C.c_int, [vp, i64, vp, i64, i64]),
            "qsb_buffer_ptr": (vp, [vp]),
            "qsb_buffer_bytes": (i64, [vp]),
            "qsb_host_alloc": (C.c_int, [i64, P(vp)]),
            "qsb_host_free": (C.c_int, [vp]),
            "qsb_program_create": (C.c_int, [vp, i32, i32, vp, i64, i64, i64, vp, i64, vp, i64, i32, i32, i32, P(vp)]),
            "qsb_program_free": (C.c_int, [vp]),
            "qsb_run": (C.c_int, [vp, P(RunArgs)]),
            "qsb_debug_profile": (C.c_int, [vp, C.c_int, vp, i64]),
            "qsb_probabilities": (C.c_int, [vp, i32, vp, i64, i64, vp, i64]),
            "qsb_probabilities_sum": (C.c_int, [vp, i32, vp, i64, i64, vp]),
            "qsb_sample_index": (C.c_int, [vp, i32, vp, i64, i64, vp, vp]),
            "qsb_overlap": (C.c_int, [vp, i32, vp, i64, vp, i64, i64, i64, vp]),
            "qsb_masked_parity": (C.c_int, [vp, i32, vp, i64, i64, vp, i32, vp]),
            "qsb_rdm_all": (C.c_int, [vp, i32, vp, i64, i64, vp, vp]),
            "qsb_rdm_general": (C.c_int, [vp, i32, vp, i64, i64, vp, i32, vp]),
            "qsb_mi_all_pairs": (C.c_int, [vp, i32, vp, i64, i64, vp, vp]),
            "qsb_rho_accumulate": (C.c_int, [vp, i32, vp, i64, i64, dbl, vp]),
            "qsb_readout_transform": (C.c_int, [vp, i32, vp, i64, dbl, dbl]),
            "qsb_stream_create": (C.c_int, [vp, i32, i32, i32, i32, P(i32), P(i32), vp, i32, vp, i64, P(vp)]),
            "qsb_stream_run": (C.c_int, [vp, vp, i64, vp, i64, i32]),
            "qsb_stream_run_peers": (C.c_int, [vp, P(vp), i32, i32, i64, vp, i64, i32]),
            "qsb_stream_run_scatter": (C.c_int, [vp, vp, i64, P(vp), i32, i32, i64, i32]),
            "qsb_stream_free": (C.c_int, [vp]),
        }
        for name, (res, args) in proto.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


_ERR = {-1: ValueError, -5: NotImplementedError}


def _check(rc, ctx_handle=None):
    if rc == 0:
        return
    msg = load_library().qsb_last_error(ctx_handle)
    msg = msg.decode() if msg else f"libqsb error {rc}"
    raise _ERR.get(rc, RuntimeError)(msg)


def _hostptr(arr):
    return arr.ctypes.data_as(C.c_void_p)


class _LockedLib:
    """libqsb entry points serialised on one lock.  A `qsb_ctx` is single-threaded by contract and ctypes drops the
    GIL during a call; the reference drives the engine from Qt worker threads (controller/simulation_controller.py:
    228-256), so the process-wide Context funnels every call of every thread through its lock."""

    def __init__(self, lib, lock):
        self._lib, self._lock = lib, lock

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        lock = self._lock

        def call(*args):
            with lock:
                return fn(*args)

        call.__name__ = name
        self.__dict__[name] = call
        return call


class Buffer:
    """Device bytes owned by (or wrapped for) a Context."""

    def __init__(self, ctx, handle, nbytes):
        self.ctx, self.handle, self.nbytes = ctx, handle, nbytes

    def upload(self, arr, offset=0):
        arr = np.ascontiguousarray(arr)
        _check(self.ctx.lib.qsb_buffer_upload(self.handle, offset, _hostptr(arr), arr.nbytes), self.ctx.handle)
        return self

    def upload_async(self, arr, offset=0):
        """No host wait; `arr` (pinned, contiguous) must stay untouched until an Event recorded afterwards is waited."""
        assert arr.flags.c_contiguous
        _check(self.ctx.lib.qsb_buffer_upload_async(self.handle, offset, _hostptr(arr), arr.nbytes), self.ctx.handle)
        return self

    def download(self, dtype, shape, offset=0, out=None):
        if out is None:
            out = np.empty(shape, dtype=dtype)
        _check(self.ctx.lib.qsb_buffer_download(self.handle, offset, _hostptr(out), out.nbytes), self.ctx.handle)
        return out

    def zero(self, offset=0, nbytes=None):
        _check(self.ctx.lib.qsb_buffer_zero(self.handle, offset, self.nbytes - offset if nbytes is None else nbytes),
               self.ctx.handle)
        return self

    def copy_from(self, src, nbytes, dst_off=0, src_off=0):
        _check(self.ctx.lib.qsb_buffer_copy(self.handle, dst_off, src.handle, src_off, nbytes), self.ctx.handle)
        return self

    @property
    def ptr(self):
        return self.ctx.lib.qsb_buffer_ptr(self.handle)

    def free(self):
        if self.handle is not None:
            if self.ctx.handle is not None:          # a closed context has already released its pool
                self.ctx.lib.qsb_buffer_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Event:
    """Marker on the context's stream (qsb_event_*)."""

    def __init__(self, ctx):
        self.ctx = ctx
        h = C.c_void_p()
        _check(ctx.lib.qsb_event_create(ctx.handle, C.byref(h)), ctx.handle)
        self.handle = h

    def record(self):
        _check(self.ctx.lib.qsb_event_record(self.handle), self.ctx.handle)
        return self

    def wait(self):
        _check(self.ctx.lib.qsb_event_wait(self.handle), self.ctx.handle)

    def __del__(self):
        try:
            if self.handle is not None:
                if self.ctx.handle is not None:
                    self.ctx.lib.qsb_event_free(self.handle)
                self.handle = None
        except Exception:
            pass


class DeviceProgram:
    def __init__(self, ctx, prog: Program):
        self.ctx, self.prog = ctx, prog
        ops = np.ascontiguousarray(prog.ops, dtype=OP_DTYPE)
        cdata = np.ascontiguousarray(prog.cdata, dtype=np.float64)
        idata = np.ascontiguousarray(prog.idata, dtype=np.int32)
        n_ops = prog.ops_stride if prog.ops_stride else len(ops)
        h = C.c_void_p()
        _check(ctx.lib.qsb_program_create(ctx.handle, prog.n, prog.m, _hostptr(ops), n_ops, prog.ops_stride,
                                          prog.n_programs, _hostptr(cdata), len(cdata), _hostptr(idata), len(idata),
                                          prog.load_perm, prog.store_perm, prog.n_snapshots, C.byref(h)), ctx.handle)
        self.handle = h

    def free(self):
        if self.handle is not None:
            if self.ctx.handle is not None:
                self.ctx.lib.qsb_program_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DevicePass:
    """One streamed pass on the device (qsb_stream_*): a `stream.StreamPass` with its sweeps uploaded."""

    def __init__(self, ctx, spass, cdata):
        from .stream import pack_blocks
        self.ctx, self.spass = ctx, spass
        arr = pack_blocks(spass)
        cd = np.ascontiguousarray(cdata, dtype=np.float64)
        n = spass.n
        pos = (C.c_int32 * n)(*[int(x) for x in spass.positions])
        pos_out = (C.c_int32 * n)(*[int(x) for x in spass.positions_out])
        h = C.c_void_p()
        _check(ctx.lib.qsb_stream_create(ctx.handle, n, spass.m, spass.l, spass.e, pos, pos_out,
                                         C.cast(arr, C.c_void_p), len(spass.blocks), _hostptr(cd), len(cd), C.byref(h)),
               ctx.handle)
        self.handle = h

    def run(self, src, dst=None, *, src_offset=0, dst_offset=0, async_=True):
        _check(self.ctx.lib.qsb_stream_run(self.handle, src.handle, src_offset, dst.handle if dst is not None else None,
                                           dst_offset, RUN_ASYNC if async_ else 0), self.ctx.handle)

    def run_peers(self, peer_ptrs, peer_shift, peer_rank_or, dst, *, dst_offset=0, async_=True):
        """Load the tiles from the peers' shards (device pointers, rank order): the qubit exchange folded into the pass."""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        _check(self.ctx.lib.qsb_stream_run_peers(self.handle, arr, len(peer_ptrs), int(peer_shift), int(peer_rank_or),
                                                 dst.handle, dst_offset, RUN_ASYNC if async_ else 0), self.ctx.handle)

    def run_scatter(self, src, peer_ptrs, peer_shift, peer_rank_or, *, src_offset=0, async_=True):
        """Store every box into the shard of the peer its destination index names: the qubit exchange folded into the
        store of this pass (peer_ptrs = the ranks' destination shard pointers, rank order)."""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        _check(self.ctx.lib.qsb_stream_run_scatter(self.handle, src.handle, src_offset, arr, len(peer_ptrs), int(peer_shift),
                                                   int(peer_rank_or), RUN_ASYNC if async_ else 0), self.ctx.handle)

    def free(self):
        if self.handle is not None:
            if self.ctx.handle is not None:
                self.ctx.lib.qsb_stream_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA device + stream + memory pool.  The C object is single-threaded by contract; every call made
    through this wrapper takes `self.lock`, so one Context can be shared by all threads of the process
    (`get_context`)."""

    def __init__(self, device=0):
        self.lock = threading.RLock()
        self.lib = _LockedLib(load_library(), self.lock)
        h = C.c_void_p()
        _check(self.lib.qsb_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = device
        sm, ma, mi, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        _check(self.lib.qsb_ctx_info(h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)), h)
        self.sm_count, self.cc, self.total_mem = sm.value, (ma.value, mi.value), mem.value
        self.precision = "c128"
        self._staging = {}

    # -- precision -------------------------------------------------------------------------
    def set_precision(self, precision):
        """"c128" (default, the reference's complex128) or "c64" (complex64 state buffers, tolerance 1e-5).
        Programs keep the mode they were created in; state buffers must be allocated / uploaded in that type."""
        if precision not in ("c128", "c64"):
            raise ValueError("precision must be 'c128' or 'c64'")
        _check(self.lib.qsb_ctx_set_precision(self.handle, 1 if precision == "c64" else 0), self.handle)
        self.precision = precision

    @property
    def amp_bytes(self):
        return 8 if self.precision == "c64" else 16

    @property
    def amp_dtype(self):
        return np.complex64 if self.precision == "c64" else np.complex128

    # -- memory -------------------------------------------------------------------------
    def alloc(self, nbytes):
        h = C.c_void_p()
        _check(self.lib.qsb_buffer_alloc(self.handle, int(nbytes), C.byref(h)), self.handle)
        return Buffer(self, h, int(nbytes))

    def wrap(self, device_ptr, nbytes):
        h = C.c_void_p()
        _check(self.lib.qsb_buffer_wrap(self.handle, C.c_void_p(device_ptr), int(nbytes), C.byref(h)), self.handle)
        return Buffer(self, h, int(nbytes))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        return self.alloc(max(arr.nbytes, 16)).upload(arr)

    def apply_dense(self, n, states, first, count, out, out_first, target_bits, matrix, out_perm=None):
        """Dense k-qubit operator (k <= 8) on states at rest, out of place (qsb_apply_dense)."""
        k = len(target_bits)
        tb = (C.c_int32 * k)(*[int(b) for b in target_bits])
        m = np.ascontiguousarray(matrix, dtype=np.complex128).reshape(2 ** k, 2 ** k)
        perm = (C.c_int32 * n)(*[int(b) for b in out_perm]) if out_perm is not None else None
        _check(self.lib.qsb_apply_dense(self.handle, n, states.handle, first, count, out.handle, out_first, k, tb,
                                        _hostptr(m), perm), self.handle)

    def trim(self):
        """Give the cached device blocks of this context's pool back to the driver."""
        _check(self.lib.qsb_ctx_trim(self.handle), self.handle)

    def event(self):
        return Event(self)

    def staging(self, tag, shape, dtype=np.float64):
        """Cached pinned host array for `tag`, grown on demand (callers fill it, upload_async it, and wait for an
        event before refilling)."""
        need = int(np.prod(shape))
        cur = self._staging.get(tag)
        if cur is None or cur.size < need or cur.dtype != np.dtype(dtype):
            cur = self.pinned((max(need, 1),), dtype)
            self._staging[tag] = cur
        return cur[:need].reshape(shape)

    def pinned(self, shape, dtype):
        """NumPy array backed by pinned host memory (kept alive by the array's base object)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _check(self.lib.qsb_host_alloc(max(n, 16), C.byref(p)))
        raw = (C.c_char * max(n, 16)).from_address(p.value)
        arr = np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape).view(PinnedArray)
        arr._owner = _Pinned(self.lib, p)      # freed when the last view dies
        return arr

    # -- execution ------------------------------------------------------------------------
    def program(self, prog: Program):
        return DeviceProgram(self, prog)

    def stream_pass(self, spass, cdata):
        return DevicePass(self, spass, cdata)

    def run(self, dprog, count, *, states=None, first=0, load=False, store=True, params=None, params_stride=0,
            uniforms=None, uniforms_stride=0, seed=0, traj_offset=0, init_basis=None, default_basis=0,
            branches=None, branches_stride=0, snapshots=None, probs_accum=None, async_=False, normalize=None,
            states_out=None, out_first=0, load_broadcast=False, peer_table=None, peer_shift=0, peer_rank_or=0):
        a = RunArgs()
        a.states = states.handle if states is not None else None
        a.first, a.count = first, count
        a.params = params.handle if params is not None else None
        a.params_stride = params_stride
        a.uniforms = uniforms.handle if uniforms is not None else None
        a.uniforms_stride = uniforms_stride
        a.philox_seed, a.traj_offset = seed & 0xFFFFFFFFFFFFFFFF, traj_offset
        a.init_basis = init_basis.handle if init_basis is not None else None
        a.default_basis = default_basis
        a.branches = branches.handle if branches is not None else None
        a.branches_stride = branches_stride
        a.snapshots = snapshots.handle if snapshots is not None else None
        a.probs_accum = probs_accum.handle if probs_accum is not None else None
        a.states_out = states_out.handle if states_out is not None else None
        a.out_first = out_first
        a.peer_table = peer_table.handle if peer_table is not None else None
        a.peer_shift, a.peer_rank_or = int(peer_shift), int(peer_rank_or)
        norm = dprog.prog.normalize if normalize is None else normalize
        a.flags = ((RUN_LOAD if load else 0) | (RUN_STORE if store and (states is not None or states_out is not None) else 0) |
                   (RUN_NORMALIZE if norm else 0) | (RUN_ASYNC if async_ else 0) |
                   (RUN_ACCUM_PROBS if probs_accum is not None else 0) |
                   (RUN_LOAD_BROADCAST if load_broadcast else 0))
        _check(self.lib.qsb_run(dprog.handle, C.byref(a)), self.handle)

    def profile(self, enable=True, read=False):
        """Developer aid: executor cycle counters (see qsb_debug_profile); returns uint64[ctas][32] when read."""
        out = np.zeros((8 * 148 * 4, 128), dtype=np.uint64) if read else None
        n = self.lib.qsb_debug_profile(self.handle, 1 if enable else 0, _hostptr(out) if read else None,
                                       out.shape[0] if read else 0)
        if n < 0:
            _check(n, self.handle)
        return out[:n] if read else None

    def sync(self):
        _check(self.lib.qsb_ctx_sync(self.handle), self.handle)

    def set_stream(self, cuda_stream_ptr):
        _check(self.lib.qsb_ctx_set_stream(self.handle, C.c_void_p(cuda_stream_ptr)), self.handle)

    def timer_start(self):
        _check(self.lib.qsb_timer_start(self.handle), self.handle)

    def timer_stop(self):
        ms = C.c_float()
        _check(self.lib.qsb_timer_stop(self.handle, C.byref(ms)), self.handle)
        return ms.value

    @property
    def launches(self):
        return int(self.lib.qsb_launch_count(self.handle))

    # -- reductions -----------------------------------------------------------------------
    def probabilities(self, n, states, first, count, out, out_first=0):
        _check(self.lib.qsb_probabilities(self.handle, n, states.handle, first, count, out.handle, out_first), self.handle)

    def probabilities_sum(self, n, states, first, count, out):
        _check(self.lib.qsb_probabilities_sum(self.handle, n, states.handle, first, count, out.handle), self.handle)

    def sample_index(self, n, states, first, count, uniforms, out):
        _check(self.lib.qsb_sample_index(self.handle, n, states.handle, first, count, uniforms.handle, out.handle),
               self.handle)

    def overlap(self, n, a, a_first, b, b_first, b_stride, count, out):
        _check(self.lib.qsb_overlap(self.handle, n, a.handle, a_first, b.handle, b_first, b_stride, count, out.handle),
               self.handle)

    def masked_parity(self, n, states, first, count, masks, out):
        mk = np.ascontiguousarray(masks, dtype=np.uint64)
        _check(self.lib.qsb_masked_parity(self.handle, n, states.handle, first, count, _hostptr(mk), len(mk),
                                          out.handle), self.handle)

    def rdm_all(self, n, states, first, count, rdm1, rdm2):
        _check(self.lib.qsb_rdm_all(self.handle, n, states.handle, first, count,
                                    rdm1.handle if rdm1 is not None else None,
                                    rdm2.handle if rdm2 is not None else None), self.handle)

    def rdm_general(self, n, states, first, count, keep_qubits, out):
        kq = np.ascontiguousarray(keep_qubits, dtype=np.int32)
        _check(self.lib.qsb_rdm_general(self.handle, n, states.handle, first, count, _hostptr(kq), len(kq), out.handle),
               self.handle)

    def mi_all_pairs(self, n, states, first, count, mi, entropy1=None):
        _check(self.lib.qsb_mi_all_pairs(self.handle, n, states.handle, first, count, mi.handle,
                                         entropy1.handle if entropy1 is not None else None), self.handle)

    def rho_accumulate(self, n, states, first, count, scale, rho):
        _check(self.lib.qsb_rho_accumulate(self.handle, n, states.handle, first, count, scale, rho.handle), self.handle)

    def readout_transform(self, n, probs, count, p01, p10):
        _check(self.lib.qsb_readout_transform(self.handle, n, probs.handle, count, p01, p10), self.handle)

    def close(self):
        """Destroy the C context (stream, events, pool).  Buffers / programs / events still alive afterwards only
        drop their Python handles (their device memory went with the pool)."""
        with self.lock:
            if self.handle is not None:
                h, self.handle = self.handle, None
                self._staging.clear()
                load_library().qsb_ctx_destroy(h)


class _Pinned:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.qsb_host_free(self.ptr)
        except Exception:
            pass


class PinnedArray(np.ndarray):
    """ndarray over cudaHostAlloc memory; views share `_owner`, which frees the block."""

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


_ctx_lock = threading.Lock()
_contexts = {}          # (device, precision) -> Context, shared by every thread of the process


def default_device():
    for key in ("QSB_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


def get_context(device=None, precision="c128"):
    """The process-wide Context of (device, precision).

    One context per device instead of one per thread: the reference's GUI starts a fresh QThread for every
    simulation (controller/simulation_controller.py:222-256), and a context per thread would leave a stream, a
    memory pool with all its cached state batches and a program cache behind for each of them.  Calls are
    serialised by the context's lock; results created on a worker thread stay valid on any other thread.
    complex64 users get a context of their own (`precision="c64"`), so the shared complex128 context never changes
    its element size under the engine classes."""
    if precision not in ("c128", "c64"):
        raise ValueError("precision must be 'c128' or 'c64'")
    dev = default_device() if device is None else device
    key = (dev, precision)
    with _ctx_lock:
        c = _contexts.get(key)
        if c is None or c.handle is None:
            c = Context(dev)
            if precision == "c64":
                c.set_precision("c64")
            _contexts[key] = c
    return c


def close_all():
    """Destroy every cached context (tests; long-running hosts that want their device memory back)."""
    with _ctx_lock:
        for c in list(_contexts.values()):
            c.close()
        _contexts.clear()
