"""Freeze the real reference's concurrence, EntanglementEventDetector, ConvergenceAnalysis and BenchmarkAnalysis outputs
(analysis.py:194-219, :255-413, :420-621) into tests/golden/golden_analysis2.json.
Build container only:  python tests/golden/make_golden_analysis2.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "quantum-simulator_b200"))     # qsb.workloads only

import numpy as np                                                                     # noqa: E402
from quantum_sim.engine.analysis import (StateAnalysis, EntanglementEventDetector, ConvergenceAnalysis,  # noqa: E402
                                         BenchmarkAnalysis)
from quantum_sim.engine.simulator import Simulator                                      # noqa: E402
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance                      # noqa: E402
from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise                       # noqa: E402
from quantum_sim.engine.measurement import MeasurementEngine                             # noqa: E402
from qsb.workloads import layered_circuit                                                # noqa: E402

assert "/root/reference" in sys.modules["quantum_sim"].__file__
J = {}


def circ(n, gates):
    qc = QuantumCircuit(n)
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    return qc


def ev_rows(evs):
    return [[e.step, list(e.qubit_pair), e.event_type.value, e.magnitude, e.entropy_before, e.entropy_after] for e in evs]


n = 5
gates = [list(g) for g in layered_circuit(n, 6, 11)]
res = Simulator().run(circ(n, gates), shots=0, record_steps=True)
J["conc"] = {"n": n, "gates": gates,
             "values": [[a, b, StateAnalysis.concurrence(res.final_state, a, b)] for a in range(n) for b in range(n) if a != b]}
bell = circ(2, [("H", [0], [], 0), ("CNOT", [0, 1], [], 1)])
J["conc_bell"] = StateAnalysis.concurrence(Simulator().run(bell, shots=0).final_state, 0, 1)
dets = []
for kw in ({}, {"epsilon": 0.05, "persistence": 2}, {"epsilon": 0.02, "epsilon_on": 0.2, "epsilon_off": 0.05}):
    det = EntanglementEventDetector(**kw)
    per_step = [ev_rows(det.process_step(st, i)) for i, st in enumerate(res.step_states)]
    hist = det.get_pair_history(3, 1)
    det.reset()
    after = ev_rows(det.process_step(res.step_states[-1], 99))
    dets.append({"kw": kw, "per_step": per_step, "history_3_1": hist, "after_reset": after,
                 "n_timeline": len(det.get_timeline())})
J["detector"] = dets
sv = res.final_state
J["shot_conv"] = ConvergenceAnalysis.shot_convergence(sv, [10, 100, 1000, 5000], seed=5)
counts = MeasurementEngine.sample(sv, 300, rng=np.random.default_rng(8))
J["tvd"] = ConvergenceAnalysis.tvd(sv.probabilities, counts, 300)
J["kl"] = ConvergenceAnalysis.kl_divergence(sv.probabilities, counts, 300)
J["kl_eps"] = ConvergenceAnalysis.kl_divergence(sv.probabilities, counts, 300, epsilon=1e-3)
J["counts300"] = counts
J["qv_ideal"] = BenchmarkAnalysis.quantum_volume(max_qubits=4, num_trials=5, seed=3)
nm = NoiseModel()
nm.add_global_noise(DepolarizingNoise(0.2))
nm.set_seed(17)
J["qv_noisy"] = BenchmarkAnalysis.quantum_volume(max_qubits=3, num_trials=6, noise_model=nm, seed=4)
with open(os.path.join(HERE, "golden_analysis2.json"), "w") as f:
    json.dump(J, f, indent=1)
print("wrote golden_analysis2.json")
