"""Where does Simulator.run_with_noise spend host time?  (GPU box; prints per-phase wall times for a few repeats)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "quantum-simulator_b200"))
from qsb import capi
from qsb.workloads import layered_circuit, to_gate_instances
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
from quantum_sim.engine.simulator import Simulator

n, T = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
qc = QuantumCircuit(n)
for g in to_gate_instances(layered_circuit(n, 64, 2026), GateInstance):
    qc.add_gate(g)
nm = NoiseModel(); nm.add_global_noise(DepolarizingNoise(0.01)); nm.add_global_noise(AmplitudeDampingNoise(0.02)); nm.set_seed(1)
sim = Simulator(nm)
c = capi.get_context()
dp, _ = sim._program(qc)
d = dp.prog.n_draws
sim.run_with_noise(qc, shots=64, seed=1)
for rep in range(4):
    t = [time.perf_counter()]
    u = nm._rng.random(T * d).reshape(T, d); t.append(time.perf_counter())
    states = c.alloc(T * 16 << n); t.append(time.perf_counter())
    ud = c.to_device(u); t.append(time.perf_counter())
    c.run(dp, T, states=states, uniforms=ud, uniforms_stride=d); c.sync(); t.append(time.perf_counter())
    del states, ud; t.append(time.perf_counter())
    names = ["draws", "alloc", "h2d", "kernel", "free"]
    print(rep, {k: round((b - a) * 1e3, 2) for k, a, b in zip(names, t, t[1:])}, flush=True)
    t0 = time.perf_counter(); r = sim.run_with_noise(qc, shots=T, seed=5 + rep); t1 = time.perf_counter()
    print(rep, "run_with_noise ms", round((t1 - t0) * 1e3, 2), "traj/s", round(T / (t1 - t0), 1), flush=True)
