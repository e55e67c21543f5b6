"""Measurement and sampling -- the reference's measurement.py API.

Device work: basis rotation (H / S-dagger then H on every qubit, measurement.py:91-98, through the
executor so the reference's axis scramble carries over), probabilities, the readout confusion transform.
Host work: `rng.multinomial` / `rng.random`, because they must consume the caller's NumPy generator
(the number of draws multinomial makes depends on the data).
"""

from __future__ import annotations

from enum import Enum

import numpy as np

from qsb import runtime
from qsb.compiler import Lowering
from .gates import H_MATRIX, S_DAG_MATRIX
from .state_vector import StateVector


class MeasurementBasis(Enum):
    Z = "Z"
    X = "X"
    Y = "Y"


def lower_basis_rotation(lw: Lowering, basis) -> None:
    """Append the X/Y basis change for every qubit to a lowering (no-op for Z)."""
    name = basis.value if isinstance(basis, MeasurementBasis) else str(basis)
    if name == "Z":
        return
    for q in range(lw.n):
        if name == "Y":
            lw.matrix(S_DAG_MATRIX, [q])
        lw.matrix(H_MATRIX, [q])


def _counts_from_array(counts_array, n) -> dict:
    nz = np.nonzero(counts_array)[0]
    return {format(int(i), f"0{n}b"): int(counts_array[i]) for i in nz}


def _normalised(probs):
    total = probs.sum()
    if total > 1e-15:
        return probs / total
    return np.ones_like(probs) / len(probs)


def readout_shots(counts: dict, readout_error, n: int, rng) -> dict:
    """Shot-mode readout corruption (measurement.py:121-127 + noise.py:128-139), vectorised: shots are
    visited in the counts dict's order, one uniform per bit, left to right -- `rng.random(k)` returns the
    same stream as k scalar calls."""
    if not counts:
        return counts
    from .noise import ReadoutError
    if type(readout_error).apply_to_bitstring is not ReadoutError.apply_to_bitstring:
        # a subclass with its own per-shot corruption: call it shot by shot exactly like measurement.py:121-127
        out: dict = {}
        for bitstring, count in counts.items():
            for _ in range(count):
                noisy = readout_error.apply_to_bitstring(bitstring, rng)
                out[noisy] = out.get(noisy, 0) + 1
        return out
    keys = list(counts.keys())
    reps = np.fromiter((counts[k] for k in keys), dtype=np.int64, count=len(keys))
    bits = np.array([[ch == "1" for ch in k] for k in keys], dtype=bool)
    true_bits = np.repeat(bits, reps, axis=0)
    r = rng.random(true_bits.size).reshape(true_bits.shape)
    flip = np.where(true_bits, r < readout_error.p10, r < readout_error.p01)
    measured = true_bits ^ flip
    weights = 1 << np.arange(n - 1, -1, -1, dtype=np.int64)
    idx = measured.astype(np.int64) @ weights
    out: dict = {}
    for i in idx.tolist():                      # insertion order = first occurrence, as in the reference
        key = format(i, f"0{n}b")
        out[key] = out.get(key, 0) + 1
    return out


class MeasurementEngine:
    """Sampling helpers; same static API as the reference."""

    @staticmethod
    def measure_qubit(state: StateVector, qubit: int, rng=None):
        collapsed = state.copy()
        outcome = collapsed.measure_qubit(qubit, rng)
        return outcome, collapsed

    @staticmethod
    def measure_all(state: StateVector, rng=None):
        collapsed = state.copy()
        bitstring = collapsed.measure_all(rng)
        return bitstring, collapsed

    @staticmethod
    def sample(state: StateVector, shots: int, rng=None) -> dict:
        rng = rng or np.random.default_rng()
        probs = _normalised(state.probabilities)
        return _counts_from_array(rng.multinomial(shots, probs), state.num_qubits)

    @staticmethod
    def sample_with_basis(state: StateVector, shots: int, basis: MeasurementBasis = MeasurementBasis.Z,
                          readout_error=None, readout_mode: str = "shot", rng=None) -> dict:
        rng = rng or np.random.default_rng()
        n = state.num_qubits
        if basis != MeasurementBasis.Z:
            layout = state.layout

            def build():
                lw = Lowering(n, layout=layout)
                lower_basis_rotation(lw, basis)
                return lw.finish()

            rotated = state.copy()
            dp = runtime.cached_program(("basis", n, layout, basis.value), build)
            runtime.run_single(n, rotated._device(), dp)
            rotated._touched_on_device()
        else:
            rotated = state

        if readout_error is not None and readout_mode == "distribution":
            probs = rotated.probabilities
            total = probs.sum()
            if total > 1e-15:
                probs = probs / total
            noisy = readout_error.apply_to_distribution(probs, n)
            return _counts_from_array(rng.multinomial(shots, noisy), n)

        counts = MeasurementEngine.sample(rotated, shots, rng=rng)
        if readout_error is not None and readout_mode == "shot":
            counts = readout_shots(counts, readout_error, n, rng)
        return counts
