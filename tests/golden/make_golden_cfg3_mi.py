"""Golden values for BASELINE config 3's per-trajectory, per-layer mutual information: the REAL reference
(imported from /root/reference, this container only) runs layered_circuit(12, 16, 2026) with depolarizing(0.01) +
amplitude_damping(0.02) for two noise seeds with record_steps, and StateAnalysis.mutual_information
(analysis.py:183-191 -> partial_trace :120-166 -> eigvalsh :99-104) is evaluated for all 66 pairs of two layers of each.

    python tests/golden/make_golden_cfg3_mi.py          (about 2 minutes: every pair builds a 4096 x 4096 outer product)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(HERE)), "quantum-simulator_b200"))
import quantum_sim
assert quantum_sim.__file__.startswith("/root/reference"), quantum_sim.__file__
from quantum_sim.engine.analysis import StateAnalysis
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
from quantum_sim.engine.simulator import Simulator
from qsb.workloads import layered_circuit          # tuples only: no engine import

N, DEPTH, SEED = 12, 16, 2026
LAYERS = (5, 15)
NOISE_SEEDS = (101, 202)


def main():
    qc = QuantumCircuit(N)
    for name, targets, params, col in layered_circuit(N, DEPTH, SEED):
        qc.add_gate(GateInstance(name, list(targets), list(params), col))
    out = {"n": N, "depth": DEPTH, "circuit_seed": SEED, "layers": list(LAYERS), "noise_seeds": list(NOISE_SEEDS), "mi": {}}
    t0 = time.time()
    for s in NOISE_SEEDS:
        nm = NoiseModel()
        nm.add_global_noise(DepolarizingNoise(0.01))
        nm.add_global_noise(AmplitudeDampingNoise(0.02))
        nm.set_seed(s)
        res = Simulator(nm).run(qc, shots=0, record_steps=True, seed=s)
        assert len(res.step_states) == DEPTH
        for layer in LAYERS:
            st = res.step_states[layer]
            vals = [StateAnalysis.mutual_information(st, i, j) for i in range(N) for j in range(i + 1, N)]
            out["mi"][f"{s}:{layer}"] = vals
            print(s, layer, f"{time.time() - t0:.0f}s", max(vals), flush=True)
    with open(os.path.join(HERE, "golden_cfg3_mi.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
