"""Ad-hoc timing probe for the trajectory kernel (developer tool, not part of the product)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from qsb import capi
from qsb.lowering import lower_circuit
from qsb.workloads import layered_circuit, config3_noise
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.gate_registry import GateRegistry
from oracle import qsim_oracle as O

ctx = capi.get_context(0)
print("qsb_version", ctx.lib.qsb_version(), "SMs", ctx.sm_count, flush=True)
REG = GateRegistry.instance()


def prog_for(n, depth, noise, local_bits=None, seed=2026):
    gates = layered_circuit(n, depth, seed)
    qc = QuantumCircuit(n)
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    ch = (lambda name: [(k, p, None) for k, p in O.channels_for(noise, name)]) if noise else None
    prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, ch, local_bits=local_bits)
    return prog


def timeit(label, prog, T, reps=3, store=True):
    dim = 1 << prog.n
    dp = ctx.program(prog)
    states = ctx.alloc(T * dim * 16)
    kw = {}
    if prog.n_draws:
        u = np.random.default_rng(0).random((T, prog.n_draws))
        kw.update(uniforms=ctx.to_device(u), uniforms_stride=prog.n_draws)
    best = 1e9
    for r in range(reps):
        ctx.timer_start()
        ctx.run(dp, T, states=states, store=store, async_=True, **kw)
        best = min(best, ctx.timer_stop())
    kinds = np.bincount(prog.ops["kind"], minlength=60)
    print(f"{label:34s} n={prog.n} m={prog.m} T={T:5d} ops={len(prog.ops):5d} remaps={prog.n_remaps:4d} "
          f"draws={prog.n_draws:5d}  {best:9.2f} ms  {T / best * 1e3:9.1f} traj/s  {best / T * 1e3:8.1f} us/traj",
          flush=True)


noise = config3_noise()
for T in (15, 120, 960):
    timeit("16q noisy", prog_for(16, 64, noise), T)
timeit("16q noiseless", prog_for(16, 64, None), 960)
timeit("16q noisy depth 8", prog_for(16, 8, noise), 960)
timeit("13q noisy (C=1)", prog_for(13, 64, noise), 1480)
timeit("13q noiseless (C=1)", prog_for(13, 64, None), 1480)
timeit("14q noisy (C=2)", prog_for(14, 64, noise), 960)
timeit("12q noisy (C=1)", prog_for(12, 64, noise), 1480)
only_ad = {"global": [("amplitude_damping", 0.02)], "gate": {}}
only_dep = {"global": [("depolarizing", 0.01)], "gate": {}}
timeit("16q only AD", prog_for(16, 64, only_ad), 960)
timeit("16q only depol", prog_for(16, 64, only_dep), 960)
