"""Executor cycle-counter probe (developer tool): where do workers / the control warp spend their cycles?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from qsb import capi
import tools.probe_traj as _pt
prog_for, ctx = _pt.prog_for, _pt.ctx
from qsb.workloads import config3_noise

KINDS = ["exit", "init", "sweep", "remap", "gflush", "rdm1", "store", "?"]


def prof(label, prog, T):
    dim = 1 << prog.n
    dp = ctx.program(prog)
    states = ctx.alloc(T * dim * 16)
    kw = {}
    if prog.n_draws:
        u = np.random.default_rng(0).random((T, prog.n_draws))
        kw.update(uniforms=ctx.to_device(u), uniforms_stride=prog.n_draws)
    ctx.profile(True)
    ctx.run(dp, T, states=states, **kw)
    p = ctx.profile(True, read=True).astype(np.float64)
    ctx.profile(False)
    C = 1 << (prog.n - prog.m)
    units = max(1, T // (len(p) // C))
    p0 = p[0] / units
    print(f"--- {label}: CTAs={len(p)} units/cluster={units}  per unit, CTA 0 (cycles):")
    print(f"   control total {p0[18]:10.0f}   ring-full wait {p0[17]:10.0f}   descriptors {p0[19]:7.1f}")
    print(f"   control: decode {p0[20]:10.0f}  fold {p0[21]:10.0f}  slow ops {p0[22]:10.0f} (n={p0[23]:6.1f}, {p0[22] / max(p0[23], 1):6.0f} each, excl. ring wait)")
    print(f"   control emit_sweep: n={p0[26]:6.1f} body {p0[24] / max(p0[26], 1):7.0f}  publish {p0[25] / max(p0[26], 1):7.0f} cycles each")
    ne = max(p0[26], 1)
    print(f"      emit body split: syncwarp {p0[27] / ne:6.0f}  pending copy {p0[28] / ne:6.0f}  group order {p0[29] / ne:6.0f}  rest {(p0[24] - p0[27] - p0[28] - p0[29]) / ne:6.0f}")
    print(f"   worker wait   {p0[0]:10.0f}")
    for k, name in enumerate(KINDS):
        if p0[9 + k]:
            print(f"   {name:7s} n={p0[9 + k]:7.1f}  busy {p0[1 + k]:10.0f}  ({p0[1 + k] / p0[9 + k]:8.0f} / desc)")
    print(f"      sweep set-up before the first load: {p0[124] / max(p0[11], 1):6.0f} cycles / sweep")
    G = ["none", "cx", "cz", "swap", "ccx", "cswap", "dense", "?"]
    for key in range(32):
        if p0[64 + key]:
            print(f"      sweep k={key // 8} gate={G[key % 8]:6s} n={p0[64 + key]:7.1f}  {p0[32 + key] / p0[64 + key]:8.0f} / desc")
    print(f"      remap: first cluster barrier {p0[120] / max(p0[12], 1):8.0f} / remap")
    nr = max(p0[12], 1)
    print(f"      remap: pull issue {p0[121] / nr:8.0f}  second barrier {p0[122] / nr:8.0f}  write {p0[123] / nr:8.0f};  rounds with 1/2/3 pairs: "
          f"{p0[125]:.0f} {p0[126]:.0f} {p0[127]:.0f}")
    print("      per-warp busy:", " ".join(f"{p0[104 + w]:9.0f}" for w in range(8)))
    print("      per-warp wait:", " ".join(f"{p0[112 + w]:9.0f}" for w in range(8)))
    for nd in range(4):
        if p0[100 + nd]:
            print(f"      sweeps with {nd} dense pending: n={p0[100 + nd]:7.1f}  {p0[96 + nd] / p0[100 + nd]:8.0f} / desc")


noise = config3_noise()
prof("16q noisy", prog_for(16, 64, noise), 150)

