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
        if a.shape != b.shape:
            raise ValueError("cannot take the overlap of vectors of different lengths")   # np.vdot raises likewise
        length = a.shape[0]
        n = max((length - 1).bit_length(), 1)
        if length != 1 << n:
            # the reference takes any pair of equal-length vectors (analysis.py:37-40); zero padding up to the next
            # power of two leaves vdot unchanged and keeps the arithmetic on the device
            a = np.concatenate([a, np.zeros((1 << n) - length, dtype=np.complex128)])
            b = np.concatenate([b, np.zeros((1 << n) - length, dtype=np.complex128)])
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

    @staticmethod
    def concurrence(state: StateVector, qubit_a: int, qubit_b: int) -> float:
        """Wootters concurrence of the pair (analysis.py:194-219): the 4x4 RDM comes from the device, the 4x4 algebra
        (spin flip, eigenvalues of rho * rho~) stays on the host like the reference's."""
        rho = StateAnalysis.partial_trace(state, [qubit_a, qubit_b])
        flip = np.kron(Y_MATRIX, Y_MATRIX)
        ev = np.real(np.linalg.eigvals(rho @ (flip @ rho.conj() @ flip)))
        lam = np.sort(np.sqrt(np.maximum(ev, 0.0)))[::-1]
        return float(max(0.0, lam[0] - np.sum(lam[1:])))

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



# =================================================================================================
# Entanglement events (analysis.py:255-413): pairwise I(A:B) per step, hysteresis, persistence
# =================================================================================================
class EntanglementEventType(Enum):
    CREATION = "creation"
    DISENTANGLEMENT = "disentanglement"
    INCREASE = "increase"
    DECREASE = "decrease"


@dataclass
class EntanglementEvent:
    step: int
    qubit_pair: tuple
    event_type: EntanglementEventType
    magnitude: float
    entropy_before: float
    entropy_after: float


class EntanglementEventDetector:
    """Same state machine as the reference's detector; the n(n-1)/2 mutual informations of a step come from ONE pass
    over the state on the device (`all_pairs_mutual_information`) instead of three `partial_trace` calls per pair."""

    def __init__(self, epsilon: float = 0.01, epsilon_on: float | None = None, epsilon_off: float | None = None,
                 persistence: int = 1):
        self.epsilon_on = epsilon if epsilon_on is None else epsilon_on
        self.epsilon_off = epsilon * 0.5 if epsilon_off is None else epsilon_off
        self.epsilon = epsilon
        self.persistence = max(1, persistence)
        self._prev_mi: dict = {}
        self._entangled: dict = {}
        self._pending: dict = {}
        self._pending_type: dict = {}
        self._events: list = []
        self._pair_history: dict = {}

    def _classify(self, pair, mi, delta):
        linked = self._entangled.get(pair, False)
        if not linked and mi >= self.epsilon_on:
            return EntanglementEventType.CREATION
        if linked and mi < self.epsilon_off:
            return EntanglementEventType.DISENTANGLEMENT
        if abs(delta) > self.epsilon:
            return EntanglementEventType.INCREASE if delta > 0 else EntanglementEventType.DECREASE
        return None

    def process_step(self, state: StateVector, step_index: int) -> list:
        n = state.num_qubits
        mis = all_pairs_mutual_information(state)
        fired = []
        k = 0
        for i in range(n):
            for j in range(i + 1, n):
                pair, mi = (i, j), float(mis[k])
                k += 1
                self._pair_history.setdefault(pair, []).append((step_index, mi))
                before = self._prev_mi.get(pair, 0.0)
                delta = mi - before
                kind = self._classify(pair, mi, delta)
                if kind is None:
                    self._pending.pop(pair, None)
                    self._pending_type.pop(pair, None)
                else:
                    if self._pending_type.get(pair) == kind:
                        self._pending[pair] = self._pending.get(pair, 0) + 1
                    else:
                        self._pending[pair] = 1
                        self._pending_type[pair] = kind
                    if self._pending.get(pair, 0) >= self.persistence:
                        if kind is EntanglementEventType.CREATION:
                            self._entangled[pair] = True
                        elif kind is EntanglementEventType.DISENTANGLEMENT:
                            self._entangled[pair] = False
                        ev = EntanglementEvent(step=step_index, qubit_pair=pair, event_type=kind, magnitude=abs(delta),
                                               entropy_before=before, entropy_after=mi)
                        fired.append(ev)
                        self._events.append(ev)
                        self._pending[pair] = 0
                        self._pending_type.pop(pair, None)
                self._prev_mi[pair] = mi
        return fired

    def get_timeline(self) -> list:
        return list(self._events)

    def get_pair_history(self, qa: int, qb: int) -> list:
        return list(self._pair_history.get((min(qa, qb), max(qa, qb)), []))

    def get_all_pair_histories(self) -> dict:
        return dict(self._pair_history)

    def reset(self) -> None:
        # like the reference (analysis.py:409-413): history and last values only; the entangled flags and the
        # persistence counters survive a reset
        self._prev_mi.clear()
        self._events.clear()
        self._pair_history.clear()


# =================================================================================================
# Shot-count convergence (analysis.py:420-497) and benchmarks (analysis.py:500-621)
# =================================================================================================
class ConvergenceAnalysis:
    @staticmethod
    def _empirical(ideal_probs, counts, total_shots):
        dim = len(ideal_probs)
        width = int(np.log2(dim))
        emp = np.zeros(dim)
        for key, c in counts.items():
            if len(key) == width:
                try:
                    emp[int(key, 2)] = c / total_shots
                except ValueError:
                    pass
        return emp

    @staticmethod
    def tvd(ideal_probs: np.ndarray, empirical_counts: dict, total_shots: int) -> float:
        """0.5 * sum_i |p_i - count_i / shots|, accumulated in index order like the reference's loop."""
        emp = ConvergenceAnalysis._empirical(ideal_probs, empirical_counts, total_shots)
        acc = 0.0
        for d in np.abs(np.asarray(ideal_probs, dtype=np.float64) - emp).tolist():
            acc += d
        return float(0.5 * acc)

    @staticmethod
    def kl_divergence(ideal_probs: np.ndarray, empirical_counts: dict, total_shots: int, epsilon: float = 1e-10) -> float:
        """D_KL(ideal || empirical) in bits with epsilon smoothing of the empirical side."""
        p = np.asarray(ideal_probs, dtype=np.float64)
        q = ConvergenceAnalysis._empirical(p, empirical_counts, total_shots) + epsilon
        acc = 0.0
        for pi, qi in zip(p.tolist(), q.tolist()):
            if pi >= epsilon:
                acc += pi * np.log2(pi / qi)
        return float(max(0.0, acc))

    @staticmethod
    def shot_convergence(state: StateVector, shot_counts: list, seed: int | None = None) -> list:
        from .measurement import MeasurementEngine
        ideal = state.probabilities
        rng = np.random.default_rng(seed)
        rows = []
        for shots in shot_counts:
            child = np.random.default_rng(rng.integers(0, 2 ** 63))
            counts = MeasurementEngine.sample(state, shots, rng=child)
            rows.append({"shots": shots, "tvd": ConvergenceAnalysis.tvd(ideal, counts, shots),
                         "kl_divergence": ConvergenceAnalysis.kl_divergence(ideal, counts, shots)})
        return rows


class BenchmarkAnalysis:
    @staticmethod
    def gate_timing(num_qubits_range, gate_matrix: np.ndarray, target_qubits_func, repetitions: int = 20) -> list:
        """Wall time of `apply_gate` on a fresh |0..0> per width (includes the device launch and synchronisation)."""
        import time
        rows = []
        for nq in num_qubits_range:
            targets = target_qubits_func(nq)
            ms = []
            for _ in range(repetitions):
                sv = StateVector(nq)
                t0 = time.perf_counter()
                sv.apply_gate(gate_matrix, targets)
                runtime.ctx().sync()
                ms.append((time.perf_counter() - t0) * 1000)
            rows.append({"num_qubits": nq, "mean_time_ms": float(np.mean(ms)), "std_time_ms": float(np.std(ms))})
        return rows

    @staticmethod
    def quantum_volume(max_qubits: int = 8, num_trials: int = 100, noise_model: object | None = None,
                       seed: int | None = None) -> dict:
        """Heavy-output test on the reference's random Rz-Ry-Rz layer circuits (widths 2..min(max_qubits, 8))."""
        from .circuit import GateInstance, QuantumCircuit
        from .simulator import Simulator
        rng = np.random.default_rng(seed)
        per_width, best = [], 1
        for m in range(2, min(max_qubits + 1, 9)):
            heavy = 0
            for _ in range(num_trials):
                qc = QuantumCircuit(num_qubits=m)
                for col in range(m):
                    for q in range(m):
                        a, b, c = rng.uniform(0, 2 * np.pi, 3)
                        for off, (name, ang) in enumerate((("Rz", a), ("Ry", b), ("Rz", c))):
                            qc.add_gate(GateInstance(name, [q], [ang], col * 3 + off))
                ideal = Simulator().run(qc, shots=0).final_state.probabilities
                actual = ideal if noise_model is None else \
                    Simulator(noise_model=noise_model).run(qc, shots=0).final_state.probabilities
                if float(np.sum(actual[ideal > float(np.median(ideal))])) > 2.0 / 3.0:
                    heavy += 1
            rate = heavy / num_trials
            ok = rate > 2.0 / 3.0
            per_width.append({"width": m, "success_rate": rate, "passed": ok})
            if ok:
                best = m
        return {"quantum_volume": 2 ** best, "log2_qv": best, "results_per_width": per_width}
