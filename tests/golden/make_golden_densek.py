"""Freeze the real reference's `StateVector.apply_gate` for dense k-qubit operators with k > 3 (state_vector.py:41-74)
into tests/golden/golden_densek.npz.  Build container only:  python tests/golden/make_golden_densek.py

Inputs are regenerated from the seed by the test (`densek_inputs`), only the reference's outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(6, [0, 1, 2, 3]), (6, [5, 2, 0, 3]), (7, [6, 0, 3, 1, 4]), (8, [2, 7, 1, 5, 0, 3]), (8, [7, 6, 5, 4, 3, 2, 1]),
         (9, [8, 1, 6, 3, 0, 5, 2, 7]), (4, [3, 1, 0, 2]), (10, [9, 4, 0, 7])]
SEED = 4242


def densek_inputs():
    """[(n, targets, matrix, psi)]: dense, non-unitary, not a Kronecker product."""
    rng = np.random.default_rng(SEED)
    out = []
    for n, targets in CASES:
        k = len(targets)
        m = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
        v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        out.append((n, targets, m, v / np.linalg.norm(v)))
    return out


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference")
    from quantum_sim.engine.state_vector import StateVector
    assert "/root/reference" in sys.modules["quantum_sim"].__file__
    arrs = {}
    for ci, (n, targets, m, psi) in enumerate(densek_inputs()):
        sv = StateVector(n)
        sv.data = psi
        sv.apply_gate(m, targets)
        arrs[f"c{ci}_out"] = sv.data.copy()
        sv.apply_gate(m.conj().T, targets[::-1])            # a second call on the scrambled state
        arrs[f"c{ci}_out2"] = sv.data.copy()
    np.savez_compressed(os.path.join(HERE, "golden_densek.npz"), **arrs)
    print("wrote golden_densek.npz:", len(arrs), "arrays")
