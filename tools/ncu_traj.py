"""Short 16-qubit noisy trajectory batch for an ncu capture of the executor kernel (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from qsb import capi
from qsb.lowering import lower_circuit
from qsb.workloads import layered_circuit, config3_noise
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.gate_registry import GateRegistry
from oracle import qsim_oracle as O

T = int(sys.argv[1]) if len(sys.argv) > 1 else 60
ctx = capi.get_context(0)
n, noise = 16, config3_noise()
gates = layered_circuit(n, 64, 2026)
qc = QuantumCircuit(n)
for g in gates:
    qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
prog, _ = lower_circuit(n, qc.get_ordered_gates(), GateRegistry.instance(),
                        lambda name: [(k, p, None) for k, p in O.channels_for(noise, name)])
dp = ctx.program(prog)
states = ctx.alloc(T * (1 << n) * 16)
u = ctx.to_device(np.random.default_rng(0).random((T, prog.n_draws)))
for r in range(3):
    ctx.timer_start()
    ctx.run(dp, T, states=states, uniforms=u, uniforms_stride=prog.n_draws, async_=True)
    print("ms", ctx.timer_stop(), "traj/s", T / ctx.timer_stop() * 1e3 if False else "")
