"""Synthetic workloads of BASELINE.json's configs.

Gates are plain tuples ``(name, targets, params, column)`` -- the same four
fields as the reference's ``GateInstance`` (circuit.py:8-14) -- so they feed
the oracle directly and convert 1:1 into ``GateInstance`` objects.
"""

from __future__ import annotations

import math

import numpy as np

LAYER_NAMES = ["H", "Rx", "Ry", "Rz", "U3", "CNOT", "CZ", "Toffoli"]
_ARITY = {"CNOT": 2, "CZ": 2, "Toffoli": 3}
_NPAR = {"Rx": 1, "Ry": 1, "Rz": 1, "U3": 3}


def layered_circuit(n, depth, seed):
    """Random layered circuit: every column is a random permutation of the
    qubits cut into H/Rx/Ry/Rz/U3/CNOT/CZ/Toffoli blocks (tail filled with H).

    (16, 64, 2026) -> 683 gates / 1024 target slots / 473 parameters."""
    rng = np.random.default_rng(seed)
    gates = []
    for col in range(depth):
        perm = rng.permutation(n).tolist()
        i = 0
        while i < n:
            g = LAYER_NAMES[int(rng.integers(0, 8))]
            k = _ARITY.get(g, 1)
            if i + k > n:
                g, k = "H", 1
            params = [float(x) for x in rng.uniform(-math.pi, math.pi, _NPAR.get(g, 0))]
            gates.append((g, perm[i:i + k], params, col))
            i += k
    return gates


def ghz(n):
    return [("H", [0], [], 0)] + [("CNOT", [0, q], [], q) for q in range(1, n)]


def config3_noise():
    """Config 3 / north-star noise: depolarizing 0.01 then amplitude damping 0.02
    as global channels, readout error (0.02, 0.05)."""
    return {"global": [("depolarizing", 0.01), ("amplitude_damping", 0.02)],
            "gate": {}, "readout": (0.02, 0.05)}


def target_slots(gates):
    return sum(len(g[1]) for g in gates)


def to_gate_instances(gates, GateInstance):
    return [GateInstance(g[0], list(g[1]), list(g[2]), g[3]) for g in gates]
