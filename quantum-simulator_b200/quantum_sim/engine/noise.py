"""Noise channels, readout error and NoiseModel -- the reference's noise.py API on the device executor.

Kraus sets are kernel inputs; the stochastic selection of noise.py:224-260 (p_i = ||K_i psi||^2, one
`Generator.choice` draw per (channel, target qubit), renormalise) runs inside the trajectory kernel.
The draws still come from the model's own NumPy generator (`_rng`, noise.py:192) in the reference's
order, so seeded runs take the same branches.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from qsb import runtime
from qsb.compiler import Lowering
from .gates import I_MATRIX, X_MATRIX, Y_MATRIX, Z_MATRIX


class NoiseChannel(ABC):
    """A single-qubit channel given by its Kraus operators."""

    kind = "generic"          # device op family (qsb/compiler.py); subclasses below override

    @abstractmethod
    def get_kraus_operators(self) -> list:
        ...

    @property
    @abstractmethod
    def probability(self) -> float:
        ...


class _ProbabilityChannel(NoiseChannel):
    _label = "Probability"

    def __init__(self, p: float):
        if not 0 <= p <= 1:
            raise ValueError(f"{self._label} must be in [0, 1], got {p}")
        self._p = p

    @property
    def probability(self) -> float:
        return self._p


class BitFlipNoise(_ProbabilityChannel):
    """X with probability p."""
    kind = "bit_flip"

    def get_kraus_operators(self):
        return [np.sqrt(1 - self._p) * I_MATRIX, np.sqrt(self._p) * X_MATRIX]


class PhaseFlipNoise(_ProbabilityChannel):
    """Z with probability p."""
    kind = "phase_flip"

    def get_kraus_operators(self):
        return [np.sqrt(1 - self._p) * I_MATRIX, np.sqrt(self._p) * Z_MATRIX]


class DepolarizingNoise(_ProbabilityChannel):
    """X, Y, Z with probability p/3 each."""
    kind = "depolarizing"

    def get_kraus_operators(self):
        w = np.sqrt(self._p / 3)
        return [np.sqrt(1 - self._p) * I_MATRIX, w * X_MATRIX, w * Y_MATRIX, w * Z_MATRIX]


class AmplitudeDampingNoise(_ProbabilityChannel):
    """Energy relaxation with rate gamma."""
    kind = "amplitude_damping"
    _label = "Gamma"

    @property
    def _gamma(self):
        return self._p

    def get_kraus_operators(self):
        g = self._p
        return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=np.complex128),
                np.array([[0, np.sqrt(g)], [0, 0]], dtype=np.complex128)]


def channel_spec(channel):
    """(kind, p, kraus_ops | None) for the lowering pass; unknown subclasses run as generic Kraus sets."""
    kind = getattr(type(channel), "kind", "generic")
    if kind in ("bit_flip", "phase_flip", "depolarizing", "amplitude_damping") and \
            type(channel) in (BitFlipNoise, PhaseFlipNoise, DepolarizingNoise, AmplitudeDampingNoise):
        return (kind, float(channel.probability), None)
    return ("generic", float(channel.probability), channel.get_kraus_operators())


class ReadoutError:
    """Classical readout confusion: p01 = P(read 1 | true 0), p10 = P(read 0 | true 1)."""

    def __init__(self, p01: float = 0.0, p10: float = 0.0):
        if not (0 <= p01 <= 1 and 0 <= p10 <= 1):
            raise ValueError("Readout error probabilities must be in [0, 1]")
        self.p01 = p01
        self.p10 = p10

    @property
    def confusion_matrix(self) -> np.ndarray:
        """C[measured][true]."""
        return np.array([[1 - self.p01, self.p10], [self.p01, 1 - self.p10]])

    def apply_to_bitstring(self, bitstring: str, rng: np.random.Generator) -> str:
        """Per-shot corruption: one rng.random() per character, in order.  Stays on the host because it
        must consume the caller's generator draw by draw (noise.py:128-139)."""
        out = []
        for ch in bitstring:
            r = rng.random()
            if ch == "0":
                out.append("1" if r < self.p01 else "0")
            else:
                out.append("0" if r < self.p10 else "1")
        return "".join(out)

    def apply_to_distribution(self, probs: np.ndarray, num_qubits: int) -> np.ndarray:
        """Per-axis confusion transform of a 2^n probability vector, then renormalise (device kernel)."""
        c = runtime.ctx()
        p = np.ascontiguousarray(probs, dtype=np.float64).reshape(-1)
        if p.shape[0] != 2 ** num_qubits:
            raise ValueError(f"cannot reshape array of size {p.shape[0]} into shape {tuple([2] * num_qubits)}")
        buf = c.to_device(p)
        c.readout_transform(num_qubits, buf, 1, float(self.p01), float(self.p10))
        return buf.download(np.float64, (2 ** num_qubits,))

    def to_dict(self) -> dict:
        return {"p01": self.p01, "p10": self.p10}

    @classmethod
    def from_dict(cls, data: dict) -> "ReadoutError":
        return cls(p01=data.get("p01", 0.0), p10=data.get("p10", 0.0))


_CHANNEL_TYPES = {c.__name__: c for c in (BitFlipNoise, PhaseFlipNoise, DepolarizingNoise, AmplitudeDampingNoise)}


class NoiseModel:
    """Which channels follow which gates (global first, then gate-specific), plus readout error."""

    def __init__(self):
        self._global_noise: list = []
        self._gate_noise: dict = {}
        self._readout_error = None
        self._rng = np.random.default_rng()

    @property
    def readout_error(self):
        return self._readout_error

    def set_readout_error(self, error: ReadoutError) -> None:
        self._readout_error = error

    def add_global_noise(self, channel: NoiseChannel):
        self._global_noise.append(channel)

    def add_gate_noise(self, gate_name: str, channel: NoiseChannel):
        self._gate_noise.setdefault(gate_name, []).append(channel)

    def set_seed(self, seed: int):
        self._rng = np.random.default_rng(seed)

    def channels_for(self, gate_name: str) -> list:
        return list(self._global_noise) + list(self._gate_noise.get(gate_name, []))

    def _channel_specs(self, gate_name: str) -> list:
        return [channel_spec(ch) for ch in self.channels_for(gate_name)]

    def _signature(self):
        """Hashable description of the model's channels (program cache key)."""
        def sig(ch):
            kind, p, ops = channel_spec(ch)
            return (kind, p, None if ops is None else tuple(runtime.matrix_key(k) for k in ops))
        return (tuple(sig(c) for c in self._global_noise),
                tuple((name, tuple(sig(c) for c in chans)) for name, chans in self._gate_noise.items()))

    def apply(self, state, gate):
        """Kraus steps that follow one gate, in place on `state` (one device launch)."""
        n = state.num_qubits
        specs = self._channel_specs(gate.gate_name)
        targets = [int(q) for q in gate.target_qubits]
        if not specs or not targets:
            return
        layout = state.layout

        def build():
            lw = Lowering(n, layout=layout)
            for kind, p, ops in specs:
                for q in targets:
                    lw.kraus(kind, p, q, ops)
            return lw.finish()

        key = ("noise", n, layout, tuple(targets), tuple((k, p, None if o is None else tuple(
            runtime.matrix_key(x) for x in o)) for k, p, o in specs))
        dp = runtime.cached_program(key, build)
        if dp.prog.n_draws == 0:
            return
        draws = self._rng.random(dp.prog.n_draws)      # choice() consumes one double per draw
        runtime.run_single(n, state._device(), dp, uniforms=draws)
        state._touched_on_device()

    def to_dict(self) -> dict:
        def enc(ch):
            return {"type": type(ch).__name__, "probability": ch.probability}
        out = {"global": [enc(c) for c in self._global_noise],
               "gate_specific": {name: [enc(c) for c in chans] for name, chans in self._gate_noise.items()}}
        if self._readout_error is not None:
            out["readout_error"] = self._readout_error.to_dict()
        return out

    @classmethod
    def from_dict(cls, data: dict) -> "NoiseModel":
        model = cls()
        for ch in data.get("global", []):
            model.add_global_noise(_CHANNEL_TYPES[ch["type"]](ch["probability"]))
        for name, chans in data.get("gate_specific", {}).items():
            for ch in chans:
                model.add_gate_noise(name, _CHANNEL_TYPES[ch["type"]](ch["probability"]))
        if "readout_error" in data:
            model.set_readout_error(ReadoutError.from_dict(data["readout_error"]))
        return model
