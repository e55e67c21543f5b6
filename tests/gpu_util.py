"""Run a lowered program on the B200 through the C ABI (libqsb.so); same signature as emu_util.emu_run."""

import numpy as np

from qsb import capi


def gpu_run(prog, count=1, T=None, states=None, params=None, uniforms=None, seed=0, traj_offset=0,
            init_basis=None, default_basis=0, want_branches=False, store=True, accum_probs=False, out_of_place=False,
            precision="c128"):
    ctx = capi.get_context(precision=precision)      # one process-wide context per (device, precision)
    return _gpu_run(ctx, prog, count, states, params, uniforms, seed, traj_offset, init_basis, default_basis,
                    want_branches, store, accum_probs, out_of_place)


def _gpu_run(ctx, prog, count, states, params, uniforms, seed, traj_offset, init_basis, default_basis, want_branches,
             store, accum_probs, out_of_place):
    dim = 1 << prog.n
    AB, AD = ctx.amp_bytes, ctx.amp_dtype
    dp = ctx.program(prog)
    load = states is not None
    if load:
        host = np.ascontiguousarray(states, dtype=np.complex128).reshape(count, dim).astype(AD)
        sbuf = ctx.to_device(host)
    else:
        sbuf = ctx.alloc(count * dim * AB).zero()
    kw = {}
    if params is not None:
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(count, -1)
        kw.update(params=ctx.to_device(p), params_stride=p.shape[1])
    if uniforms is not None:
        u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(count, -1)
        kw.update(uniforms=ctx.to_device(u), uniforms_stride=u.shape[1])
    if init_basis is not None:
        kw.update(init_basis=ctx.to_device(np.ascontiguousarray(init_basis, dtype=np.int64)))
    bbuf = sn = pb = None
    if want_branches:
        bs = max(prog.n_draws, 1)
        bbuf = ctx.to_device(np.full((count, bs), -1, dtype=np.int32))
        kw.update(branches=bbuf, branches_stride=bs)
    if prog.n_snapshots:
        sn = ctx.alloc(count * prog.n_snapshots * dim * AB).zero()
        kw.update(snapshots=sn)
    if accum_probs:
        pb = ctx.alloc(dim * 8).zero()
        kw.update(probs_accum=pb)
    obuf = ctx.alloc(count * dim * AB).zero() if out_of_place else None
    ctx.run(dp, count, states=sbuf, load=load, store=store, seed=seed, traj_offset=traj_offset,
            default_basis=default_basis, states_out=obuf, **kw)
    if obuf is not None:
        sbuf = obuf
    return dict(states=sbuf.download(AD, (count, dim)).astype(np.complex128),
                snapshots=sn.download(AD, (count, prog.n_snapshots, dim)).astype(np.complex128) if sn is not None else None,
                branches=bbuf.download(np.int32, (count, max(prog.n_draws, 1))) if bbuf is not None else None,
                probs=pb.download(np.float64, (dim,)) if pb is not None else None)
