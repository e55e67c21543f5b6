"""State analysis -- the hot-path subset of the reference's analysis.py on device reductions.

On the device: overlaps (`state_fidelity`, `expectation_value`), reduced density matrices
(`partial_trace` for 1 and 2 kept qubits in one pass over the state instead of the reference's full
2^n x 2^n outer product, analysis.py:136), all-pairs mutual information.  On the host: eigenvalues of
the resulting 2x2 / 4x4 matrices (`np.linalg.eigvalsh`, the reference's own call, analysis.py:102).
"""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

import numpy as np

from qsb import runtime
from qsb.compiler import Lowering
from .gates import X_MATRIX, Y_MATRIX, Z_MATRIX
from .state_vector import StateVector

_PAULI = {"X": X_MATRIX, "Y": Y_MATRIX, "Z": Z_MATRIX}


def _entropy_bits_batch(mats: np.ndarray) -> np.ndarray:
    """-sum lambda log2 lambda over eigenvalues > 1e-15 (analysis.py:102-104) for a stack of matrices."""
    w = np.linalg.eigvalsh(mats)
    keep = w > 1e-15
    safe = np.where(keep, w, 1.0)
    return -np.sum(np.where(keep, w * np.log2(safe), 0.0), axis=-1)


def device_rdms(state: StateVector):
    """(rdm1[n,2,2], rdm2[npairs,4,4]) of one state; pairs (i<j) row-major, first kept qubit = MSB."""
    n = state.num_qubits
    c = runtime.ctx()
    npairs = n * (n - 1) // 2
    r1 = c.alloc(n * 4 * 16)
    r2 = c.alloc(max(npairs, 1) * 16 * 16)
    c.rdm_all(n, state._device(), 0, 1, r1, r2 if npairs else None)
    return (r1.download(np.complex128, (n, 2, 2)),
            r2.download(np.complex128, (npairs, 4, 4)) if npairs else np.zeros((0, 4, 4), dtype=np.complex128))


def pair_index(n, a, b):
    i, j = (a, b) if a < b else (b, a)
    return i * (n - 1) - i * (i - 1) // 2 + (j - i - 1)


def all_pairs_mutual_information_device(n, states, first, count) -> np.ndarray:
    """float64[count][n(n-1)/2]: I(i:j) in bits for every i < j of `count` device-resident states, entirely on the
    device (RDM reductions + Jacobi eigenvalues + entropies; analysis.py:99-104, :183-191)."""
    c = runtime.ctx()
    npairs = n * (n - 1) // 2
    out = c.alloc(max(count * npairs, 1) * 8)
    c.mi_all_pairs(n, states, first, count, out)
    return out.download(np.float64, (count, npairs))


def all_pairs_mutual_information(state: StateVector) -> np.ndarray:
    """I(i:j) for every i < j in one device pass (order of analysis.py:331-333); eigenvalues on the host with the
    reference's own np.linalg.eigvalsh (see all_pairs_mutual_information_device for the all-device batch path)."""
    n = state.num_qubits
    r1, r2 = device_rdms(state)
    s1 = _entropy_bits_batch(r1)
    s2 = _entropy_bits_batch(r2) if len(r2) else np.zeros(0)
    out = np.empty(len(r2))
    k = 0
    for i in range(n):
        for j in range(i + 1, n):
            out[k] = max(0.0, s1[i] + s1[j] - s2[k])
            k += 1
    return out


class StateAnalysis:
    """Static metrics on states and density matrices (same signatures as the reference)."""

    # ---- fidelity ---------------------------------------------------------------------------
    @staticmethod
    def state_fidelity(psi: np.ndarray, phi: np.ndarray) -> float:
        """|<psi|phi>|^2."""
        a = np.ascontiguousarray(psi, dtype=np.complex128).reshape(-1)
        b = np.ascontiguousarray(phi, dtype=np.complex128).reshape(-1)
        if a.shape != b.shape or a.shape[0] & (a.shape[0] - 1):
            raise ValueError("state vectors must have the same power-of-two length")
        n = max(a.shape[0].bit_length() - 1, 1) if a.shape[0] > 1 else 0
        if n == 0:
            return float(np.abs(np.vdot(a, b)) ** 2)
        c = runtime.ctx()
        out = c.alloc(16)
        c.overlap(n, c.to_device(a), 0, c.to_device(b), 0, 1, 1, out)
        return float(np.abs(out.download(np.complex128, (1,))[0]) ** 2)

    @staticmethod
    def _sanitize_density_matrix(rho: np.ndarray) -> np.ndarray:
        """Hermitian symmetrisation + trace normalisation (analysis.py:66-77); works on stacks of matrices."""
        rho = np.asarray(rho)
        rho = (rho + np.conj(np.swapaxes(rho, -1, -2))) / 2
        tr = np.real(np.trace(rho, axis1=-2, axis2=-1))
        scale = np.where(tr > 1e-15, tr, 1.0)
        return rho / np.asarray(scale)[..., None, None]

    @staticmethod
    def _matrix_sqrt(mat: np.ndarray) -> np.ndarray:
        w, v = np.linalg.eigh(mat)
        w = np.maximum(w, 0.0)
        return (v * np.sqrt(w)[..., None, :]) @ np.conj(np.swapaxes(v, -1, -2))

    @staticmethod
    def density_fidelity(rho: np.ndarray, sigma: np.ndarray):
        """Uhlmann fidelity (Tr sqrt(sqrt(rho) sigma sqrt(rho)))^2 (analysis.py:47-64).  Small-matrix host linear
        algebra, like the reference; accepts stacks [..., d, d] and then returns an array."""
        rho = StateAnalysis._sanitize_density_matrix(rho)
        sigma = StateAnalysis._sanitize_density_matrix(sigma)
        sq = StateAnalysis._matrix_sqrt(rho)
        w = np.maximum(np.linalg.eigvalsh(sq @ sigma @ sq), 0.0)
        fid = np.minimum(np.sum(np.sqrt(w), axis=-1) ** 2, 1.0)
        return float(fid) if np.ndim(fid) == 0 else fid

    @staticmethod
    def process_fidelity(ideal: StateVector, actual: StateVector) -> float:
        c = runtime.ctx()
        out = c.alloc(16)
        c.overlap(ideal.num_qubits, ideal._device(), 0, actual._device(), 0, 1, 1, out)
        return float(np.abs(out.download(np.complex128, (1,))[0]) ** 2)

    # ---- entropy ------------------------------------------------------------------------------
    @staticmethod
    def von_neumann_entropy(state: StateVector) -> float:
        return StateAnalysis.von_neumann_entropy_dm(state.get_density_matrix())

    @staticmethod
    def von_neumann_entropy_dm(rho: np.ndarray) -> float:
        return float(_entropy_bits_batch(np.asarray(rho)[None])[0])

    @staticmethod
    def entanglement_entropy(state: StateVector, subsystem_qubits: list) -> float:
        return StateAnalysis.von_neumann_entropy_dm(StateAnalysis.partial_trace(state, subsystem_qubits))

    # ---- partial trace -------------------------------------------------------------------------
    @staticmethod
    def partial_trace(state: StateVector, keep_qubits: list) -> np.ndarray:
        """Reduced density matrix of the kept qubits (sorted; first kept = MSB of the output index)."""
        n = state.num_qubits
        keep = sorted(int(q) for q in keep_qubits)
        if any(q < 0 or q >= n for q in keep) or len(set(keep)) != len(keep):
            raise ValueError(f"bad keep_qubits {keep_qubits} for {n} qubits")
        k = len(keep)
        if k == n:
            return state.get_density_matrix()
        r1, r2 = device_rdms(state) if k <= 2 else (None, None)
        if k == 1:
            return r1[keep[0]].copy()
        if k == 2:
            return r2[pair_index(n, keep[0], keep[1])].copy()
        if k <= 6:
            c = runtime.ctx()
            out = c.alloc(16 << (2 * k))
            c.rdm_general(n, state._device(), 0, 1, keep, out)
            return out.download(np.complex128, (2 ** k, 2 ** k))
        raise NotImplementedError("device partial_trace keeps at most 6 qubits (or all of them)")

    # ---- purity ---------------------------------------------------------------------------------
    @staticmethod
    def purity(state: StateVector) -> float:
        return StateAnalysis.purity_dm(state.get_density_matrix())

    @staticmethod
    def purity_dm(rho: np.ndarray) -> float:
        """Tr(rho^2) = sum_ij rho_ij rho_ji, without forming the product."""
        r = np.asarray(rho)
        return float(np.real(np.sum(r * r.T)))

    # ---- entanglement ------------------------------------------------------------------------------
    @staticmethod
    def mutual_information(state: StateVector, qubit_a: int, qubit_b: int) -> float:
        n = state.num_qubits
        r1, r2 = device_rdms(state)
        s = _entropy_bits_batch(np.stack([r1[qubit_a], r1[qubit_b]]))
        sab = _entropy_bits_batch(r2[pair_index(n, qubit_a, qubit_b)][None])[0]
        return float(max(0.0, s[0] + s[1] - sab))

    # ---- expectation values -----------------------------------------------------------------------------
    @staticmethod
    def expectation_value(state: StateVector, observable: np.ndarray, target_qubits: list) -> complex:
        """<psi| O psi> with O applied through apply_gate (so the reference's axis scramble is included)."""
        temp = state.copy()
        temp.apply_gate(observable, target_qubits)
        c = runtime.ctx()
        out = c.alloc(16)
        c.overlap(state.num_qubits, state._device(), 0, temp._device(), 0, 1, 1, out)
        return complex(out.download(np.complex128, (1,))[0])

    @staticmethod
    def pauli_expectation(state: StateVector, pauli: str, qubit: int) -> float:
        if pauli.upper() not in _PAULI:
            raise ValueError(f"Unknown Pauli: {pauli}. Use 'X', 'Y', or 'Z'.")
        return float(np.real(StateAnalysis.expectation_value(state, _PAULI[pauli.upper()], [qubit])))
