"""Singleton gate registry (mirror of the reference's gate_registry.py API).

Table-driven: the 21 built-in names of gate_registry.py:34-148 plus user-registered custom
unitaries (`register`), looked up by the simulator for every executed gate.
"""

from __future__ import annotations

from . import gates as G
from .gates import GateDefinition, GateType

_S, _C, _M = GateType.SINGLE, GateType.CONTROLLED, GateType.MULTI

# name, display name, type, qubits, param names, matrix factory, symbol, colour, controls, targets
_BUILTINS = [
    ("I", "Identity", _S, 1, (), G._const(G.I_MATRIX), "I", "#888888", 0, 1),
    ("H", "Hadamard", _S, 1, (), G._const(G.H_MATRIX), "H", "#4A90D9", 0, 1),
    ("X", "Pauli-X", _S, 1, (), G._const(G.X_MATRIX), "X", "#E74C3C", 0, 1),
    ("Y", "Pauli-Y", _S, 1, (), G._const(G.Y_MATRIX), "Y", "#2ECC71", 0, 1),
    ("Z", "Pauli-Z", _S, 1, (), G._const(G.Z_MATRIX), "Z", "#3498DB", 0, 1),
    ("S", "S Gate", _S, 1, (), G._const(G.S_MATRIX), "S", "#9B59B6", 0, 1),
    ("S_DAG", "S† Gate", _S, 1, (), G._const(G.S_DAG_MATRIX), "S†", "#8E44AD", 0, 1),
    ("T", "T Gate", _S, 1, (), G._const(G.T_MATRIX), "T", "#E67E22", 0, 1),
    ("T_DAG", "T† Gate", _S, 1, (), G._const(G.T_DAG_MATRIX), "T†", "#D35400", 0, 1),
    ("Rx", "Rotation-X", _S, 1, ("θ",), G.rx_matrix, "Rx", "#E91E63", 0, 1),
    ("Ry", "Rotation-Y", _S, 1, ("θ",), G.ry_matrix, "Ry", "#00BCD4", 0, 1),
    ("Rz", "Rotation-Z", _S, 1, ("θ",), G.rz_matrix, "Rz", "#FF9800", 0, 1),
    ("Phase", "Phase Gate", _S, 1, ("φ",), G.phase_matrix, "P", "#795548", 0, 1),
    ("U3", "Universal U3", _S, 1, ("θ", "φ", "λ"), G.u3_matrix, "U3", "#607D8B", 0, 1),
    ("CNOT", "Controlled-NOT", _C, 2, (), G._const(G.CNOT_MATRIX), "CX", "#FF5722", 1, 1),
    ("CZ", "Controlled-Z", _C, 2, (), G._const(G.CZ_MATRIX), "CZ", "#673AB7", 1, 1),
    ("SWAP", "SWAP", _M, 2, (), G._const(G.SWAP_MATRIX), "SW", "#009688", 0, 2),
    ("Toffoli", "Toffoli (CCX)", _C, 3, (), G._const(G.TOFFOLI_MATRIX), "CCX", "#F44336", 2, 1),
    ("Fredkin", "Fredkin (CSWAP)", _C, 3, (), G._const(G.FREDKIN_MATRIX), "CSW", "#4CAF50", 1, 2),
    ("Measure", "Measurement", GateType.MEASUREMENT, 1, (), G._const(G.I_MATRIX), "M", "#FFC107", 0, 1),
    ("Barrier", "Barrier", GateType.BARRIER, 1, (), G._const(G.I_MATRIX), "||", "#BDBDBD", 0, 1),
]


class GateRegistry:
    """Maps gate names to GateDefinition objects; one shared instance per process."""

    _instance = None

    def __init__(self):
        self._gates: dict[str, GateDefinition] = {}
        self._builtin_defs: dict[str, GateDefinition] = {}
        self.generation = 0          # bumped by every register() after the built-ins: keys the device-program caches

    @classmethod
    def instance(cls) -> "GateRegistry":
        if cls._instance is None:
            reg = cls()
            reg._register_builtins()
            cls._instance = reg
        return cls._instance

    @classmethod
    def reset(cls):
        cls._instance = None

    def _register_builtins(self):
        for name, disp, typ, nq, pnames, fn, sym, col, nc, nt in _BUILTINS:
            self.register(GateDefinition(name=name, display_name=disp, gate_type=typ, num_qubits=nq,
                                         num_params=len(pnames), param_names=pnames, matrix_func=fn,
                                         symbol=sym, color=col, num_controls=nc, num_targets=nt))
        self._builtin_defs = dict(self._gates)
        self.generation = 0

    def register(self, gate_def: GateDefinition):
        self._gates[gate_def.name] = gate_def
        self.generation += 1

    def is_builtin(self, name: str) -> bool:
        """True while `name` still maps to the definition this module registered.  The device compiler only takes its
        structured kernels (X/Y/Z/CNOT/CZ/SWAP/Toffoli/Fredkin/I, in-kernel Rx/Ry/Rz/Phase/U3) for those; a custom
        or re-registered gate under a built-in name runs its own `matrix_func(*params)` like in the reference
        (simulator.py:110-114)."""
        gd = self._gates.get(name)
        return gd is not None and gd is self._builtin_defs.get(name)

    def get(self, name: str) -> GateDefinition:
        try:
            return self._gates[name]
        except KeyError:
            raise KeyError(f"Gate '{name}' not found in registry") from None

    def all_gates(self):
        return list(self._gates.values())

    def single_qubit_gates(self):
        return [g for g in self._gates.values() if g.gate_type == GateType.SINGLE]

    def multi_qubit_gates(self):
        return [g for g in self._gates.values() if g.gate_type in (GateType.CONTROLLED, GateType.MULTI)]

    def parameterized_gates(self):
        return [g for g in self._gates.values() if g.num_params > 0]

    def gate_names(self):
        return list(self._gates.keys())
