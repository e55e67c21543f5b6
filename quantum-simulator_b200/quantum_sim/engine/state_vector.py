"""StateVector -- the reference's n-qubit state API (state_vector.py:8-193) on a device-resident array.

The amplitudes live in HBM; a host NumPy mirror is materialised lazily when `data` / `_data` is read
(and then becomes authoritative, because callers mutate it in place: noise.py:260, qec.py:391-396).
Gate application runs the executor of libqsb.so on the device copy; the reference's axis scramble
(state_vector.py:66-73) is reproduced by the host compiler (qsb/compiler.py), default
``layout="reference"``.
"""

from __future__ import annotations

import numpy as np

from qsb import runtime
from qsb.compiler import Lowering, sigma, _kron_factors

MAX_QUBITS = 16


class StateVector:
    """n-qubit pure state, complex128[2^n], qubit 0 = most significant bit of the index."""

    layout = "reference"      # "textbook" disables the reference's axis scramble

    def __init__(self, num_qubits: int):
        if num_qubits < 1 or num_qubits > MAX_QUBITS:
            raise ValueError(f"num_qubits must be 1-16, got {num_qubits}")
        self._num_qubits = num_qubits
        host = np.zeros(2 ** num_qubits, dtype=np.complex128)
        host[0] = 1.0
        self._host = host
        self._dev = None
        self._dev_valid = False

    # ---- host / device coherence --------------------------------------------------------
    @classmethod
    def _blank(cls, n):
        sv = cls.__new__(cls)
        sv._num_qubits = n
        sv._host = None
        sv._dev = None
        sv._dev_valid = False
        return sv

    @classmethod
    def _from_device(cls, n, buf):
        """Adopt a device buffer holding complex128[2^n] (no copy)."""
        sv = cls._blank(n)
        sv._dev = buf
        sv._dev_valid = True
        return sv

    @classmethod
    def _from_host(cls, n, arr):
        sv = cls._blank(n)
        sv._host = arr
        return sv

    def _device(self):
        """Device buffer with the current amplitudes (uploads the host mirror if it is newer)."""
        if not self._dev_valid:
            c = runtime.ctx()
            nbytes = 16 << self._num_qubits
            if self._dev is None or self._dev.nbytes != nbytes or self._dev.ctx is not c:
                self._dev = c.alloc(nbytes)
            self._dev.upload(np.ascontiguousarray(self._host, dtype=np.complex128))
            self._dev_valid = True
        return self._dev

    def _touched_on_device(self):
        self._host = None

    @property
    def _data(self) -> np.ndarray:
        if self._host is None:
            self._host = self._dev.download(np.complex128, (2 ** self._num_qubits,))
        self._dev_valid = False          # the caller may write through the returned array
        return self._host

    @_data.setter
    def _data(self, value):
        self._host = value
        self._dev_valid = False

    # ---- reference API ------------------------------------------------------------------
    @property
    def num_qubits(self) -> int:
        return self._num_qubits

    @property
    def data(self) -> np.ndarray:
        return self._data

    @data.setter
    def data(self, value: np.ndarray):
        if value.shape != (2 ** self._num_qubits,):
            raise ValueError(f"Expected shape ({2**self._num_qubits},), got {value.shape}")
        self._data = value.astype(np.complex128)

    @property
    def probabilities(self) -> np.ndarray:
        n = self._num_qubits
        c = runtime.ctx()
        out = c.alloc(8 << n)
        c.probabilities(n, self._device(), 0, 1, out)
        return out.download(np.float64, (2 ** n,))

    def apply_gate(self, gate_matrix: np.ndarray, target_qubits: list):
        n = self._num_qubits
        targets = [int(q) for q in target_qubits]
        for q in targets:
            if q < 0 or q >= n:
                raise ValueError(f"Qubit index {q} out of range [0, {n-1}]")
        layout = self.layout
        k = len(targets)
        if k > 3 and len(set(targets)) == k:
            mat = np.asarray(gate_matrix, dtype=np.complex128)
            if mat.size == 4 ** k and _kron_factors(mat.reshape(2 ** k, 2 ** k), k) is None:
                return self._apply_dense_big(mat.reshape(2 ** k, 2 ** k), targets)

        def build():
            lw = Lowering(n, layout=layout)
            lw.matrix(gate_matrix, targets)
            return lw.finish()

        dp = runtime.cached_program(("apply", n, layout, tuple(targets), runtime.matrix_key(gate_matrix)), build)
        runtime.run_single(n, self._device(), dp)
        self._touched_on_device()

    def _apply_dense_big(self, mat, targets):
        """k > 3 operator that is not a Kronecker product: one out-of-place pass over the state at rest
        (`qsb_apply_dense`, k <= 8); the reference's axis scramble rides on the store."""
        n = self._num_qubits
        if len(targets) > 8:
            raise NotImplementedError("dense operators on more than 8 qubits are not supported by the device path")
        c = runtime.ctx()
        perm = None
        if self.layout == "reference":
            sg = sigma(n, targets)                     # afterwards array axis i holds textbook qubit sg[i]
            perm = [0] * n
            for i in range(n):
                perm[n - 1 - sg[i]] = n - 1 - i
        src = self._device()
        out = c.alloc(16 << n)
        c.apply_dense(n, src, 0, 1, out, 0, [n - 1 - q for q in targets], mat, perm)
        self._dev = out
        self._dev_valid = True
        self._touched_on_device()

    def measure_qubit(self, qubit: int, rng: np.random.Generator | None = None) -> int:
        if qubit < 0 or qubit >= self._num_qubits:
            raise ValueError(f"Qubit {qubit} out of range")
        rng = rng or np.random.default_rng()
        n = self._num_qubits
        c = runtime.ctx()
        out = c.alloc(16)
        c.masked_parity(n, self._device(), 0, 1, [1 << (n - 1 - qubit)], out)
        p0 = float(out.download(np.float64, (2,))[0])
        outcome = 0 if rng.random() < p0 else 1
        # collapse = projector on the outcome, then renormalise (state_vector.py:95-103)
        proj = np.zeros((2, 2), dtype=np.complex128)
        proj[outcome, outcome] = 1.0
        layout = self.layout

        def build():
            lw = Lowering(n, layout="textbook")
            lw.matrix(proj, [qubit])
            lw.normalize = True
            return lw.finish()

        dp = runtime.cached_program(("collapse", n, qubit, outcome), build)
        runtime.run_single(n, self._device(), dp)
        self._touched_on_device()
        return outcome

    def measure_all(self, rng: np.random.Generator | None = None) -> str:
        rng = rng or np.random.default_rng()
        n = self._num_qubits
        c = runtime.ctx()
        u = c.to_device(np.array([rng.random()], dtype=np.float64))    # Generator.choice draws one uniform
        out = c.alloc(8)
        c.sample_index(n, self._device(), 0, 1, u, out)
        idx = int(out.download(np.int64, (1,))[0])
        host = np.zeros(2 ** n, dtype=np.complex128)
        host[idx] = 1.0
        self._data = host
        return format(idx, f"0{n}b")

    def get_reduced_density_matrix(self, qubit: int) -> np.ndarray:
        n = self._num_qubits
        if qubit < 0 or qubit >= n:
            raise ValueError(f"Qubit {qubit} out of range")
        c = runtime.ctx()
        r1 = c.alloc(n * 4 * 16)
        c.rdm_all(n, self._device(), 0, 1, r1, None)
        return r1.download(np.complex128, (n, 2, 2))[qubit].copy()

    def get_bloch_coordinates(self, qubit: int):
        rho = self.get_reduced_density_matrix(qubit)
        return (float(2.0 * np.real(rho[0, 1])), float(2.0 * np.imag(rho[1, 0])),
                float(np.real(rho[0, 0] - rho[1, 1])))

    def get_density_matrix(self) -> np.ndarray:
        n = self._num_qubits
        c = runtime.ctx()
        rho = c.alloc(16 << (2 * n)).zero()
        c.rho_accumulate(n, self._device(), 0, 1, 1.0, rho)
        return rho.download(np.complex128, (2 ** n, 2 ** n))

    def copy(self) -> "StateVector":
        n = self._num_qubits
        sv = StateVector._blank(n)
        if self._dev_valid:
            c = runtime.ctx()
            sv._dev = c.alloc(16 << n).copy_from(self._dev, 16 << n)
            sv._dev_valid = True
        else:
            sv._host = self._host.copy()
        return sv

    @staticmethod
    def _basis_index(initial_states, n=None) -> int:
        n = len(initial_states) if n is None else n
        return sum(1 << (n - 1 - i) for i, bit in enumerate(initial_states) if bit)

    @classmethod
    def from_initial_states(cls, initial_states: list) -> "StateVector":
        sv = cls(len(initial_states))
        host = np.zeros(2 ** len(initial_states), dtype=np.complex128)
        host[cls._basis_index(initial_states)] = 1.0
        sv._data = host
        return sv

    def reset(self, initial_states: list | None = None):
        host = np.zeros(2 ** self._num_qubits, dtype=np.complex128)
        if initial_states and any(s != 0 for s in initial_states):
            host[self._basis_index(initial_states, self._num_qubits)] = 1.0
        else:
            host[0] = 1.0
        self._data = host

    def __repr__(self) -> str:
        return f"StateVector(num_qubits={self._num_qubits})"
