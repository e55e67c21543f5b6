"""Quantum error correction (the reference's qec.py) with the cycle batched on the device.

Same public names as the reference (qec.py:26-701): QECResult, ThresholdPoint, QECCode, BitFlipCode,
PhaseFlipCode, SteaneCode, QECSimulator, AVAILABLE_CODES.  The reference spends 92 % of a Steane cycle in
two pure-Python loops over 2^13 indices (`_compute_z_parity`, `logical_z_expectation`);
here every per-state piece is a device reduction, and `threshold_sweep` / `projection_logical_error` run
ALL trials of a sweep point as one batch:

    codeword --(per-trial Pauli errors, decided by the host from the trial's own generator)--> noisy
    noisy --(shared H rotation)--> temp ;  masked-parity reductions on noisy / temp --> syndromes
    host: decode table --> per-trial corrections --> corrected ;  overlaps + signed parity --> metrics

Which Paulis fire depends only on the random stream, never on the state (qec.py:679-693), so the host
decides them up front with the reference's generators (`default_rng(trial_seed).random()` per data qubit).
Every Pauli goes through `apply_gate`, so the reference's axis scramble (state_vector.py:66-73) differs
from trial to trial; each trial therefore has its own tiny op list and its own store permutation.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np

from qsb import runtime
from qsb.compiler import OP_DTYPE, Program, PX, PY, PZ, NOP, SNAPSHOT, U1, sigma
from .analysis import StateAnalysis
from .circuit import QuantumCircuit, GateInstance
from .gates import H_MATRIX, X_MATRIX, Y_MATRIX, Z_MATRIX
from .noise import NoiseModel, BitFlipNoise, PhaseFlipNoise, DepolarizingNoise   # noqa: F401 (reference imports)
from .simulator import Simulator
from .state_vector import StateVector

_BATCH = 8192          # trials per device batch


@dataclass
class QECResult:
    """Result of a single QEC cycle."""
    encoded_state: StateVector
    noisy_state: StateVector
    syndrome: list
    corrected_state: StateVector
    fidelity_before: float
    fidelity_after: float
    correction_applied: list
    logical_z_expectation: float = 0.0
    logical_error_detected: bool = False


@dataclass
class ThresholdPoint:
    """Result at one physical error rate in a threshold sweep."""
    physical_rate: float
    logical_rate: float
    success_rate: float
    avg_fidelity: float
    logical_z_fidelity: float = 0.0
    decoder_success_rate: float = 0.0
    projection_logical_rate: float = 0.0


def _mask(n, qubits):
    m = 0
    for q in qubits:
        m |= 1 << (n - 1 - q)
    return m


def _parity_weights(n, states, first, count, masks):
    """float64[count][len(masks)][2] = (p_even, p_odd) per mask (qec.py:466-484), up to 8 masks per launch."""
    c = runtime.ctx()
    out = c.alloc(count * len(masks) * 16)
    c.masked_parity(n, states, first, count, masks, out)
    return out.download(np.float64, (count, len(masks), 2))


def _compute_z_parity(state: StateVector, qubits: list) -> int:
    """Most likely parity of Z measurements on `qubits`: 0 if p_even >= p_odd else 1 (qec.py:466-486)."""
    n = state.num_qubits
    w = _parity_weights(n, state._device(), 0, 1, [_mask(n, qubits)])[0, 0]
    return 0 if w[0] >= w[1] else 1


def _extract_parity_syndrome(state, parity_checks, ancilla_start, rng):
    return [_compute_z_parity(state, [qa, qb]) for qa, qb in parity_checks]


class QECCode(ABC):
    """Abstract base for quantum error correcting codes (qec.py:53-151)."""

    @property
    @abstractmethod
    def name(self) -> str: ...

    @property
    @abstractmethod
    def data_qubits(self) -> int: ...

    @property
    @abstractmethod
    def ancilla_qubits(self) -> int: ...

    @property
    def total_qubits(self) -> int:
        return self.data_qubits + self.ancilla_qubits

    @property
    @abstractmethod
    def code_distance(self) -> int: ...

    @abstractmethod
    def encode(self, logical_state: int) -> StateVector: ...

    @abstractmethod
    def extract_syndrome(self, state: StateVector, rng: np.random.Generator) -> list: ...

    @abstractmethod
    def decode_syndrome(self, syndrome: list) -> list: ...

    def apply_correction(self, state: StateVector, corrections: list) -> None:
        gate_map = {"X": X_MATRIX, "Z": Z_MATRIX}
        for gate_name, qubit in corrections:
            if gate_name in gate_map and qubit < state.num_qubits:
                state.apply_gate(gate_map[gate_name], [qubit])

    @abstractmethod
    def logical_fidelity(self, state: StateVector, logical_state: int) -> float: ...

    @abstractmethod
    def logical_z_operators(self) -> list: ...

    def logical_z_expectation(self, state: StateVector) -> float:
        """<psi|Z_L|psi> = p_even - p_odd of the Z_L qubits (qec.py:131-151)."""
        n = state.num_qubits
        w = _parity_weights(n, state._device(), 0, 1, [_mask(n, self.logical_z_operators())])[0, 0]
        return float(w[0] - w[1])

    # ---- description of the cycle for the batched path (None = use the per-state methods) ----------------
    def _batch_plan(self):
        return None


class _RepetitionCode(QECCode):
    @property
    def data_qubits(self) -> int:
        return 3

    @property
    def ancilla_qubits(self) -> int:
        return 2

    @property
    def code_distance(self) -> int:
        return 1

    def logical_fidelity(self, state: StateVector, logical_state: int) -> float:
        ideal = self.encode(logical_state)
        return StateAnalysis.state_fidelity(ideal.data, state.data)

    def logical_z_operators(self) -> list:
        return [0, 1, 2]

    _corr = "X"

    def decode_syndrome(self, syndrome: list) -> list:
        s0, s1 = syndrome[0], syndrome[1]
        if s0 == 1 and s1 == 0:
            return [(self._corr, 0)]
        if s0 == 1 and s1 == 1:
            return [(self._corr, 1)]
        if s0 == 0 and s1 == 1:
            return [(self._corr, 2)]
        return []


class BitFlipCode(_RepetitionCode):
    """3-qubit bit-flip repetition code, 3 data + 2 ancilla qubits (qec.py:156-225)."""

    @property
    def name(self) -> str:
        return "Bit-Flip [3,1,1]"

    def encode(self, logical_state: int) -> StateVector:
        qc = QuantumCircuit(5)
        if logical_state == 1:
            qc.add_gate(GateInstance("X", [0], [], 0))
        qc.add_gate(GateInstance("CNOT", [0, 1], [], 1))
        qc.add_gate(GateInstance("CNOT", [0, 2], [], 2))
        return Simulator().run(qc, shots=0).final_state

    def extract_syndrome(self, state: StateVector, rng: np.random.Generator) -> list:
        return _extract_parity_syndrome(state, [(0, 1), (1, 2)], 3, rng)

    def _batch_plan(self):
        return {"syndrome": [(None, [[0, 1], [1, 2]])], "z_rot": None}


class PhaseFlipCode(_RepetitionCode):
    """3-qubit phase-flip repetition code (qec.py:230-318)."""
    _corr = "Z"

    @property
    def name(self) -> str:
        return "Phase-Flip [3,1,1]"

    def encode(self, logical_state: int) -> StateVector:
        qc = QuantumCircuit(5)
        if logical_state == 1:
            qc.add_gate(GateInstance("X", [0], [], 0))
        qc.add_gate(GateInstance("CNOT", [0, 1], [], 1))
        qc.add_gate(GateInstance("CNOT", [0, 2], [], 2))
        for q in range(3):
            qc.add_gate(GateInstance("H", [q], [], 3))
        return Simulator().run(qc, shots=0).final_state

    def extract_syndrome(self, state: StateVector, rng: np.random.Generator) -> list:
        temp = state.copy()
        for q in range(3):
            temp.apply_gate(H_MATRIX, [q])
        return _extract_parity_syndrome(temp, [(0, 1), (1, 2)], 3, rng)

    def logical_z_expectation(self, state: StateVector) -> float:
        """Logical Z of the phase-flip code is X0 X1 X2: rotate to the X basis, then the Z parity."""
        temp = state.copy()
        for q in range(3):
            temp.apply_gate(H_MATRIX, [q])
        return super().logical_z_expectation(temp)

    def _batch_plan(self):
        return {"syndrome": [([0, 1, 2], [[0, 1], [1, 2]])], "z_rot": [0, 1, 2]}


class SteaneCode(QECCode):
    """Steane [[7,1,3]] CSS code: 7 data + 6 ancilla qubits (qec.py:323-447)."""

    _HX = [[0, 2, 4, 6], [1, 2, 5, 6], [3, 4, 5, 6]]
    _HZ = [[0, 2, 4, 6], [1, 2, 5, 6], [3, 4, 5, 6]]

    @property
    def name(self) -> str:
        return "Steane [[7,1,3]]"

    @property
    def data_qubits(self) -> int:
        return 7

    @property
    def ancilla_qubits(self) -> int:
        return 6

    @property
    def code_distance(self) -> int:
        return 3

    def encode(self, logical_state: int) -> StateVector:
        """|0>_L / |1>_L = uniform superposition of the even / odd weight [7,4,3] Hamming codewords, ancillas |0>."""
        gen = np.array([[1, 0, 0, 0, 1, 1, 0], [0, 1, 0, 0, 1, 0, 1], [0, 0, 1, 0, 0, 1, 1], [0, 0, 0, 1, 1, 1, 1]], dtype=int)
        n_total = 13
        data = np.zeros(2 ** n_total, dtype=np.complex128)
        words = []
        for i in range(16):
            cw = np.array([(i >> b) & 1 for b in range(4)]) @ gen % 2
            if int(cw.sum()) % 2 == (logical_state & 1):
                words.append(cw)
        amp = 1.0 / np.sqrt(len(words))
        for cw in words:
            idx = 0
            for qi, bit in enumerate(cw.tolist()):
                if bit:
                    idx |= 1 << (n_total - 1 - qi)
            data[idx] = amp
        sv = StateVector(n_total)
        sv.data = data
        return sv

    def extract_syndrome(self, state: StateVector, rng: np.random.Generator) -> list:
        n = state.num_qubits
        w = _parity_weights(n, state._device(), 0, 1, [_mask(n, c) for c in self._HX])[0]
        x_syndrome = [0 if w[k, 0] >= w[k, 1] else 1 for k in range(3)]
        temp = state.copy()
        for q in range(7):
            temp.apply_gate(H_MATRIX, [q])
        w = _parity_weights(n, temp._device(), 0, 1, [_mask(n, c) for c in self._HZ])[0]
        return x_syndrome + [0 if w[k, 0] >= w[k, 1] else 1 for k in range(3)]

    def decode_syndrome(self, syndrome: list) -> list:
        corrections = []
        z_syn = syndrome[3:6]
        z_idx = z_syn[0] + 2 * z_syn[1] + 4 * z_syn[2]
        if 0 < z_idx <= 7:
            corrections.append(("X", z_idx - 1))
        x_syn = syndrome[0:3]
        x_idx = x_syn[0] + 2 * x_syn[1] + 4 * x_syn[2]
        if 0 < x_idx <= 7:
            corrections.append(("Z", x_idx - 1))
        return corrections

    def logical_fidelity(self, state: StateVector, logical_state: int) -> float:
        ideal = self.encode(logical_state)
        return StateAnalysis.state_fidelity(ideal.data, state.data)

    def logical_z_operators(self) -> list:
        return [0, 1, 2, 3, 4, 5, 6]

    def _batch_plan(self):
        return {"syndrome": [(None, self._HX), (list(range(7)), self._HZ)], "z_rot": None}


# ---- per-trial Pauli programs -------------------------------------------------------------------------------
_PAULI_KIND = {"X": PX, "Y": PY, "Z": PZ}
_SIGMA1 = {}


def _sigma1(n, q):
    key = (n, q)
    if key not in _SIGMA1:
        _SIGMA1[key] = sigma(n, [q])
    return _SIGMA1[key]


def _pauli_programs(n, codes, qubits, layout):
    """One tiny program per trial: its Paulis (through apply_gate, i.e. with the axis scramble), then a store of
    the result with the trial's own bit permutation.  Vectorised over the trials.

    codes, qubits: int arrays [T][L]; event l of trial t applies Pauli code (1 X, 2 Y, 3 Z; 0 = no event) to qubit
    qubits[t][l], in order l = 0..L-1.  Returns a Program with ops_stride = L + 1."""
    codes = np.asarray(codes, dtype=np.int64).reshape(len(codes), -1)
    qubits = np.asarray(qubits, dtype=np.int64).reshape(codes.shape)
    T, L = codes.shape
    stride = L + 1
    ops = np.zeros((T, stride), dtype=OP_DTYPE)
    ops["data"], ops["param"], ops["draw"] = -1, -1, -1
    boa = np.tile(np.arange(n - 1, -1, -1, dtype=np.int64), (T, 1))       # physical bit playing reference axis j
    pos = np.zeros(T, dtype=np.int64)                                      # next free op slot of each trial
    rows = np.arange(T)
    kind_of = np.array([NOP, PX, PY, PZ], dtype=np.int32)
    for l in range(L):
        fired = codes[:, l] != 0
        if not fired.any():
            continue
        r = rows[fired]
        q = qubits[fired, l]
        ops["kind"][r, pos[r]] = kind_of[codes[fired, l]]
        ops["b0"][r, pos[r]] = boa[r, q]
        pos[r] += 1
        if layout == "reference":
            for qv in np.unique(q):
                rr = r[q == qv]
                boa[rr] = boa[rr][:, np.array(_sigma1(n, int(qv)))]
    perm = np.empty((T, n), dtype=np.int64)
    np.put_along_axis(perm, boa, np.tile(np.arange(n - 1, -1, -1, dtype=np.int64), (T, 1)), axis=1)
    # distinct store permutations: rows as base-n numbers (n <= 30 digits of < 5 bits fit 63 bits only for small n, so
    # fall back to the row-wise unique when they would not); a 1-D unique is ~20x faster than the axis-0 one
    if n * max(1, int(n - 1).bit_length()) <= 62:
        shift = max(1, int(n - 1).bit_length())
        pkey = (perm << (shift * np.arange(n, dtype=np.int64))).sum(axis=1)
        _, first, inv = np.unique(pkey, return_index=True, return_inverse=True)
        uniq = perm[first]
    else:
        uniq, inv = np.unique(perm, axis=0, return_inverse=True)
    ident = list(range(n))
    idata = np.concatenate([np.array(ident + ident, dtype=np.int64), uniq.reshape(-1)])
    ops["kind"][rows, pos] = SNAPSHOT
    ops["b0"][rows, pos] = 0
    ops["aux"][rows, pos] = 2 * n + np.asarray(inv).reshape(-1) * n
    return Program(n=n, m=n, ops=ops.reshape(-1), cdata=np.zeros(2), idata=idata.astype(np.int32), load_perm=0,
                   store_perm=n, n_snapshots=1, n_draws=0, n_params=0, normalize=False, ops_stride=stride, n_programs=T)


def _noise_codes(noise_type, prob, uniforms):
    """QECSimulator._apply_noise (qec.py:669-693) on a [T][n_data] array of the per-qubit draws: Pauli code per data
    qubit (0 none, 1 X, 2 Y, 3 Z).  An unknown noise type fires nothing."""
    u = np.asarray(uniforms)
    codes = np.zeros(u.shape, dtype=np.int64)
    if noise_type == "bit_flip":
        codes[u < prob] = 1
    elif noise_type == "phase_flip":
        codes[u < prob] = 3
    elif noise_type == "depolarizing":
        codes[u < prob] = 3
        codes[u < 2 * prob / 3] = 2
        codes[u < prob / 3] = 1
    return codes


class QECSimulator:
    """Run QEC cycles with noise injection and threshold analysis (qec.py:491-693)."""

    def __init__(self, code: QECCode):
        self._code = code

    # ---- one cycle, per-state methods (any QECCode subclass) --------------------------------------------------
    def run_cycle(self, logical_state: int = 0, noise_type: str = "bit_flip", noise_prob: float = 0.1,
                  seed=None) -> QECResult:
        rng = np.random.default_rng(seed)
        ideal = self._code.encode(logical_state)
        noisy = ideal.copy()
        self._apply_noise(noisy, noise_type, noise_prob, rng)
        syndrome = self._code.extract_syndrome(noisy, rng)
        corrections = self._code.decode_syndrome(syndrome)
        corrected = noisy.copy()
        self._code.apply_correction(corrected, corrections)
        fid_before = StateAnalysis.process_fidelity(ideal, noisy)
        fid_after = StateAnalysis.process_fidelity(ideal, corrected)
        z_exp = self._code.logical_z_expectation(corrected)
        expected_sign = 1.0 if logical_state == 0 else -1.0
        return QECResult(encoded_state=ideal, noisy_state=noisy, syndrome=syndrome, corrected_state=corrected,
                         fidelity_before=fid_before, fidelity_after=fid_after, correction_applied=corrections,
                         logical_z_expectation=z_exp, logical_error_detected=(z_exp * expected_sign) < 0)

    def _apply_noise(self, state: StateVector, noise_type: str, prob: float, rng: np.random.Generator) -> None:
        mats = {"X": X_MATRIX, "Y": Y_MATRIX, "Z": Z_MATRIX}
        for q in range(self._code.data_qubits):
            # one draw per data qubit, consumed in order (the vectorised helper draws the same stream)
            if noise_type in ("bit_flip", "phase_flip"):
                if rng.random() < prob:
                    state.apply_gate(mats["X" if noise_type == "bit_flip" else "Z"], [q])
            elif noise_type == "depolarizing":
                r = rng.random()
                if r < prob / 3:
                    state.apply_gate(X_MATRIX, [q])
                elif r < 2 * prob / 3:
                    state.apply_gate(Y_MATRIX, [q])
                elif r < prob:
                    state.apply_gate(Z_MATRIX, [q])

    # ---- many cycles as device batches ------------------------------------------------------------------------
    def _rotation_program(self, n, qubits):
        """apply_gate(H, [q]) for q in qubits, in order (shared by every trial)."""
        from qsb.compiler import Lowering
        key = ("qec-h", StateVector.layout, n, tuple(qubits))

        def build():
            lw = Lowering(n, layout=StateVector.layout)
            for q in qubits:
                lw.matrix(H_MATRIX, [q])
            return lw.finish()

        return runtime.cached_program(key, build)

    def run_cycles(self, logical_states, noise_type, noise_prob, seeds, uniforms=None, lean=False):
        """Batched `run_cycle`: arrays of the per-trial scalars
        {syndrome[B][k], fidelity_before[B], fidelity_after[B], z_exp[B], logical_error[B], corrections[B]}.
        `uniforms` (float64[B][data_qubits], optional) replaces the per-trial `default_rng(seed)` draws -- the
        throughput mode of BASELINE config 4, where one vectorised generator feeds a whole sweep point.
        `lean`: skip the per-trial Python lists (`syndrome` stays an int array, `corrections` is None) -- the sweep
        estimators only read the three float / bool arrays, and the lists were most of the host time of a sweep."""
        code = self._code
        plan = code._batch_plan()
        lg = np.asarray(logical_states, dtype=np.int64).reshape(-1)
        logical_states = lg if lean else [int(x) for x in lg.tolist()]
        B = len(lg)
        if plan is None:                           # custom code: the per-state path
            rs = [self.run_cycle(l, noise_type, noise_prob, s) for l, s in zip(logical_states, seeds)]
            return {"syndrome": np.array([r.syndrome for r in rs]), "fidelity_before": np.array([r.fidelity_before for r in rs]),
                    "fidelity_after": np.array([r.fidelity_after for r in rs]),
                    "z_exp": np.array([r.logical_z_expectation for r in rs]),
                    "logical_error": np.array([r.logical_error_detected for r in rs]),
                    "corrections": [r.correction_applied for r in rs]}
        n = code.total_qubits
        dim = 1 << n
        c = runtime.ctx()
        ideal = c.alloc(2 * dim * 16)
        for l in (0, 1):
            ideal.copy_from(code.encode(l)._device(), dim * 16, dst_off=l * dim * 16)
        out = {"syndrome": None if lean else [None] * B, "fidelity_before": np.empty(B), "fidelity_after": np.empty(B),
               "z_exp": np.empty(B), "logical_error": np.empty(B, dtype=bool), "corrections": None if lean else [None] * B}
        nd = code.data_qubits
        if uniforms is None:
            # the reference's streams: trial t draws from default_rng(seed_t), one double per data qubit -- and none at
            # all for a noise type it does not know (qec.py:679-693)
            known = noise_type in ("bit_flip", "phase_flip", "depolarizing")
            uniforms = np.ones((B, nd))
            if known:
                for t, sd in enumerate(seeds):
                    uniforms[t] = np.random.default_rng(sd).random(nd)
        codes = _noise_codes(noise_type, noise_prob, np.asarray(uniforms, dtype=np.float64).reshape(B, nd))
        for logical in (0, 1):
            idx_all = np.nonzero(lg == logical)[0]
            for lo in range(0, len(idx_all), _BATCH):
                idx = idx_all[lo:lo + _BATCH]
                self._cycle_batch(c, code, plan, n, ideal, logical, idx, codes[idx], out, lean)
        if not lean:
            out["syndrome"] = np.array(out["syndrome"])
        return out

    def _cycle_batch(self, c, code, plan, n, ideal, logical, idx, codes, out, lean=False):
        dim = 1 << n
        cnt = len(idx)
        layout = StateVector.layout
        # noisy = Paulis(ideal codeword); data qubit q is hit (if at all) in order q = 0, 1, ...
        noisy = c.alloc(cnt * dim * 16)
        qs = np.tile(np.arange(codes.shape[1]), (cnt, 1))
        c.run(c.program(_pauli_programs(n, codes, qs, layout)), cnt, states=ideal, first=logical, load=True, store=False,
              load_broadcast=True, snapshots=noisy)
        # syndromes: parity reductions on noisy, or on an H-rotated copy
        temp = None
        bits = []
        for rot, checks in plan["syndrome"]:
            src = noisy
            if rot is not None:
                if temp is None:
                    temp = c.alloc(cnt * dim * 16)
                c.run(self._rotation_program(n, rot), cnt, states=noisy, load=True, store=True, states_out=temp)
                src = temp
            w = _parity_weights(n, src, 0, cnt, [_mask(n, ch) for ch in checks])
            bits.append(np.where(w[:, :, 0] >= w[:, :, 1], 0, 1))
        syndrome = np.concatenate(bits, axis=1)
        # decode once per distinct syndrome (at most 64), then spread over the trials with one gather
        skey = syndrome @ (1 << np.arange(syndrome.shape[1], dtype=np.int64))
        _, first, inv = np.unique(skey, return_index=True, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        decoded = [code.decode_syndrome(syndrome[f].tolist()) for f in first.tolist()]
        width = max((len(v) for v in decoded), default=0)
        tab_c = np.zeros((len(decoded), max(width, 1)), dtype=np.int64)
        tab_q = np.zeros((len(decoded), max(width, 1)), dtype=np.int64)
        gate_code = {"X": 1, "Z": 3}
        for k, cs in enumerate(decoded):
            for l, (g, q) in enumerate(cs):
                if g in gate_code and q < n:                                   # qec.py:113-116
                    tab_c[k, l], tab_q[k, l] = gate_code[g], q
        ccode, cq = tab_c[inv], tab_q[inv]
        corrected = c.alloc(cnt * dim * 16)
        c.run(c.program(_pauli_programs(n, ccode, cq, layout)), cnt, states=noisy, load=True, store=False, snapshots=corrected)
        # fidelities |<ideal|.>|^2 and <Z_L>
        ov = c.alloc(cnt * 16)
        c.overlap(n, noisy, 0, ideal, logical, 0, cnt, ov)
        fb = np.abs(ov.download(np.complex128, (cnt,))) ** 2
        c.overlap(n, corrected, 0, ideal, logical, 0, cnt, ov)
        fa = np.abs(ov.download(np.complex128, (cnt,))) ** 2
        zsrc = corrected
        if plan["z_rot"] is not None:
            if temp is None:
                temp = c.alloc(cnt * dim * 16)
            c.run(self._rotation_program(n, plan["z_rot"]), cnt, states=corrected, load=True, store=True, states_out=temp)
            zsrc = temp
        w = _parity_weights(n, zsrc, 0, cnt, [_mask(n, code.logical_z_operators())])[:, 0]
        z = w[:, 0] - w[:, 1]
        sign = 1.0 if logical == 0 else -1.0
        if not lean:
            for k, t in enumerate(np.asarray(idx).tolist()):
                out["syndrome"][t] = syndrome[k].tolist()
                out["corrections"][t] = decoded[inv[k]]
        out["fidelity_before"][idx] = fb
        out["fidelity_after"][idx] = fa
        out["z_exp"][idx] = z
        out["logical_error"][idx] = (z * sign) < 0

    def threshold_sweep(self, noise_probs: list, n_trials: int = 100, noise_type: str = "bit_flip", seed=None) -> list:
        """Sweep the physical error rate (qec.py:551-622): ONE generator spans all points and trials."""
        rng = np.random.default_rng(seed)
        results = []
        for p in noise_probs:
            seeds = [int(rng.integers(0, 2 ** 63)) for _ in range(n_trials)]
            r = self.run_cycles([t % 2 for t in range(n_trials)], noise_type, p, seeds)
            results.append(self._point(p, r, n_trials))
        return results

    def threshold_sweep_sharded(self, noise_probs: list, n_trials: int = 100, noise_type: str = "bit_flip", seed=None,
                                _run_cycles=None) -> list:
        """`threshold_sweep` with the trials of every point split over the ranks of the default process group (one
        process per GPU, launched by torchrun).  The seed chain is sequential (qec.py:574-586), so every rank walks the
        whole chain and keeps its contiguous slice; the per-trial scalars meet in one gather per point and every rank
        sums them in trial order -- the result equals the single-process sweep bit for bit.  No data-path collective.
        `_run_cycles` lets the CPU gloo test stand in for the device (tests only)."""
        from qsb import distributed as D
        world, rank = D.world_info()
        run = _run_cycles or self.run_cycles
        rng = np.random.default_rng(seed)
        results = []
        for p in noise_probs:
            seeds = [int(rng.integers(0, 2 ** 63)) for _ in range(n_trials)]
            lo, hi = D.shard_bounds(n_trials, world, rank)
            r = run([t % 2 for t in range(lo, hi)], noise_type, p, seeds[lo:hi]) if hi > lo else \
                {"fidelity_after": np.empty(0), "z_exp": np.empty(0), "logical_error": np.empty(0, dtype=bool)}
            packed = np.stack([np.asarray(r["fidelity_after"], dtype=np.float64), np.asarray(r["z_exp"], dtype=np.float64),
                               np.asarray(r["logical_error"], dtype=np.float64)], axis=1).reshape(-1)
            allp = D.gather_concat(packed).reshape(n_trials, 3)
            merged = {"fidelity_after": allp[:, 0], "z_exp": allp[:, 1], "logical_error": allp[:, 2] != 0.0}
            results.append(self._point(p, merged, n_trials))
        return results

    def threshold_sweep_philox(self, noise_probs: list, n_trials: int = 100000, noise_type: str = "depolarizing",
                               seed: int = 0, batch: int = 16384, _run_cycles=None) -> list:
        """Throughput mode of the sweep (BASELINE config 4 with 10^5 trials per point): the per-qubit error draws come
        from a COUNTER-BASED generator -- NumPy's Philox keyed by (seed, point, batch) -- instead of one
        `default_rng(trial_seed)` per trial (qec.py:585-586), so no host generator is created
        per trial and any rank can produce exactly its own slice.  The trials of every point are split over the ranks
        of the default process group; the only collective is ONE all-reduce of the per-point sums at the very end.
        Same estimators as `threshold_sweep` (qec.py:607-620), statistically equivalent, not stream-compatible."""
        from qsb import distributed as D
        world, rank = D.world_info()
        run = _run_cycles or self.run_cycles
        lean_kw = {} if _run_cycles is not None else {"lean": True}        # the test stand-in keeps the old signature
        nd = self._code.data_qubits
        sums = np.zeros((len(noise_probs), 4))                      # successes, fidelity, |<Z_L>|, no-logical-error
        n_batches = (n_trials + batch - 1) // batch
        for k, p in enumerate(noise_probs):
            # batch j of point k has its own Philox key: any rank can produce exactly its batches, and the result
            # does not depend on how many ranks share the work
            lo, hi = D.shard_bounds(n_batches, world, rank)
            for j in range(lo, hi):
                b0, b1 = j * batch, min(n_trials, (j + 1) * batch)
                gen = np.random.Generator(np.random.Philox(key=[int(seed) & (2 ** 64 - 1), (k << 32) | j]))
                u = gen.random((b1 - b0, nd))
                r = run(np.arange(b0, b1, dtype=np.int64) % 2, noise_type, p, None, uniforms=u, **lean_kw)
                fa = np.asarray(r["fidelity_after"], dtype=np.float64)
                sums[k] += (np.count_nonzero(fa > 0.5), fa.sum(), np.abs(np.asarray(r["z_exp"], dtype=np.float64)).sum(),
                            np.count_nonzero(~np.asarray(r["logical_error"], dtype=bool)))
        sums = D.allreduce_sum_numpy(sums)
        out = []
        for k, p in enumerate(noise_probs):
            succ, fid, zf, zok = sums[k]
            out.append(ThresholdPoint(physical_rate=p, logical_rate=1.0 - succ / n_trials, success_rate=succ / n_trials,
                                      avg_fidelity=fid / n_trials, logical_z_fidelity=zf / n_trials,
                                      decoder_success_rate=zok / n_trials, projection_logical_rate=1.0 - fid / n_trials))
        return out

    @staticmethod
    def _point(p, r, n_trials):
        # the reference accumulates trial by trial in Python floats; do the same so the sums round identically
        successes, total_fid, z_fid_sum, z_ok = 0, 0.0, 0.0, 0
        for t in range(n_trials):
            fa = float(r["fidelity_after"][t])
            if fa > 0.5:
                successes += 1
            total_fid += fa
            z_fid_sum += abs(float(r["z_exp"][t]))
            if not bool(r["logical_error"][t]):
                z_ok += 1
        return ThresholdPoint(physical_rate=p, logical_rate=1.0 - successes / n_trials, success_rate=successes / n_trials,
                              avg_fidelity=total_fid / n_trials, logical_z_fidelity=z_fid_sum / n_trials,
                              decoder_success_rate=z_ok / n_trials, projection_logical_rate=1.0 - total_fid / n_trials)

    def projection_logical_error(self, logical_state: int, noise_type: str, noise_prob: float, n_trials: int = 100,
                                 seed=None) -> dict:
        rng = np.random.default_rng(seed)
        seeds = [int(rng.integers(0, 2 ** 63)) for _ in range(n_trials)]
        r = self.run_cycles([logical_state] * n_trials, noise_type, noise_prob, seeds)
        fid_sum = 0.0
        for t in range(n_trials):
            fid_sum += float(r["fidelity_after"][t])
        mean_fid = fid_sum / n_trials
        return {"mean_fidelity": mean_fid, "logical_error_rate": 1.0 - mean_fid,
                "z_sign_error_rate": int(np.sum(r["logical_error"])) / n_trials, "n_trials": n_trials}


AVAILABLE_CODES = {
    "Bit-Flip [3,1,1]": BitFlipCode,
    "Phase-Flip [3,1,1]": PhaseFlipCode,
    "Steane [[7,1,3]]": SteaneCode,
}
