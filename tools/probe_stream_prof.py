"""Streamed-pass cycle probe (developer tool): per tile of a 2^n state, where do load / sweeps / store go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from qsb.bigstate import BigState, plan_distributed
from qsb.workloads import layered_circuit
from quantum_sim.engine.gate_registry import GateRegistry
from test_bigstate import ordered

KINDS = ["exit", "init", "sweep", "remap", "gflush", "rdm1", "store", "?"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
npass = int(sys.argv[2]) if len(sys.argv) > 2 else 4
gl = ordered(n, layered_circuit(n, 20, 2026))
st = BigState(n, layout="textbook", distributed=False)
lw = st.lowering()
reg = GateRegistry.instance()
for name, targets, params in gl:
    lw.gate(name, targets, params, reg.get(name).matrix_func)
steps, _ = plan_distributed(lw, 0, None)
ctx = st.ctx
print(f"n={n} gates={len(gl)} passes={len(steps)} ops/pass={[len(s.prog.ops) for s in steps]}")
for k, s in enumerate(steps[:npass]):
    dp = ctx.program(s.prog)
    ctx.run(dp, 1, states=st._wrapped[0], load=True, store=True)          # warm
    ctx.timer_start()
    ctx.run(dp, 1, states=st._wrapped[0], load=True, store=True, async_=True)
    ms = ctx.timer_stop()
    ctx.profile(True)
    ctx.run(dp, 1, states=st._wrapped[0], load=True, store=True)
    p = ctx.profile(True, read=True).astype(np.float64)
    ctx.profile(False)
    tiles = (1 << (n - s.prog.m)) / len(p)
    p0 = p.mean(axis=0) / tiles
    print(f"pass {k}: {len(s.prog.ops)} ops, {ms:.3f} ms = {2 * 16 * 2 ** n / ms / 1e6:.0f} GB/s real; CTAs={len(p)}, tiles/CTA={tiles:.1f}; "
          f"per tile: control {p0[18]:.0f} worker-wait {p0[0]:.0f}")
    for kk, name in enumerate(KINDS):
        if p0[9 + kk]:
            print(f"     {name:7s} n={p0[9 + kk]:5.1f} busy {p0[1 + kk]:9.0f} ({p0[1 + kk] / p0[9 + kk]:7.0f} each)")
