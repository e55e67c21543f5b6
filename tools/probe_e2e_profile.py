"""cProfile of Simulator.run_with_noise on the headline workload (developer tool): host time outside the kernel.
   python tools/probe_e2e_profile.py [shots=2048] [calls=3]"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200")]
import numpy as np
import bench

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
from quantum_sim.engine.simulator import Simulator
qc, nm = bench.headline_circuit_and_noise() if hasattr(bench, "headline_circuit_and_noise") else (None, None)
if qc is None:
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
    from qsb.workloads import layered_circuit
    qc = QuantumCircuit(16)
    for g in layered_circuit(16, 64, 2026):
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    nm = NoiseModel()
    nm.add_global_noise(DepolarizingNoise(0.01))
    nm.add_global_noise(AmplitudeDampingNoise(0.02))
nm.set_seed(7)
sim = Simulator(nm)
sim.run_with_noise(qc, shots=256, seed=1)                   # warm-up: program, pools
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for k in range(calls):
    sim.run_with_noise(qc, shots=shots, seed=k)
pr.disable()
dt = time.perf_counter() - t0
print(f"{calls} calls x {shots} shots: {dt * 1e3 / calls:.1f} ms per call = {shots * calls / dt:.0f} trajectories/s")
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(18)
print(out.getvalue()[:5000])
