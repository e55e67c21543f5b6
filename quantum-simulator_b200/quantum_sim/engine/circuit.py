"""Circuit data model (mirror of the reference's circuit.py API).

Consumed by the simulator shim, which lowers `get_ordered_gates()` into one device program
(qsb/compiler.py).  Same fields and JSON schema as circuit.py:8-173 of the reference so `.qsim`
files and `core/serialization.py` round-trip unchanged.
"""

from __future__ import annotations

from dataclasses import dataclass, field

MAX_QUBITS = 16


@dataclass
class GateInstance:
    """One placed gate: registry name, target qubits (order matters), parameters, column."""
    gate_name: str
    target_qubits: list
    params: list = field(default_factory=list)
    column: int = 0

    def to_dict(self) -> dict:
        return {"name": self.gate_name, "targets": self.target_qubits,
                "params": self.params, "column": self.column}

    @classmethod
    def from_dict(cls, data: dict) -> "GateInstance":
        return cls(gate_name=data["name"], target_qubits=data["targets"],
                   params=data.get("params", []), column=data.get("column", 0))


@dataclass
class QuantumCircuit:
    """Gate instances on `num_qubits` wires plus per-qubit initial basis states."""
    num_qubits: int = 4
    gates: list = field(default_factory=list)
    initial_states: list = field(default_factory=list)

    def __post_init__(self):
        self._fit_initial_states(self.num_qubits)

    def _fit_initial_states(self, n):
        cur = list(self.initial_states) if self.initial_states else []
        self.initial_states = (cur + [0] * n)[:n]

    # -- editing ----------------------------------------------------------------------
    def add_gate(self, gate: GateInstance):
        self.gates.append(gate)

    def remove_gate(self, gate: GateInstance):
        if gate in self.gates:
            self.gates.remove(gate)

    def move_gate(self, gate: GateInstance, new_col: int, new_targets: list):
        if gate in self.gates:
            gate.column = new_col
            gate.target_qubits = new_targets

    def clear(self):
        self.gates.clear()

    def set_num_qubits(self, n: int):
        if n < 1 or n > MAX_QUBITS:
            raise ValueError(f"num_qubits must be 1-16, got {n}")
        self.gates = [g for g in self.gates if all(q < n for q in g.target_qubits)]
        self.num_qubits = n
        self._fit_initial_states(n)

    def toggle_qubit_initial_state(self, qubit: int) -> None:
        if 0 <= qubit < self.num_qubits:
            self.initial_states[qubit] = 1 - self.initial_states[qubit]

    def set_qubit_initial_state(self, qubit: int, state: int) -> None:
        if 0 <= qubit < self.num_qubits and state in (0, 1):
            self.initial_states[qubit] = state

    # -- queries ----------------------------------------------------------------------
    def get_column_count(self) -> int:
        return max((g.column for g in self.gates), default=-1) + 1

    def get_gates_at_column(self, col: int) -> list:
        return [g for g in self.gates if g.column == col]

    def get_ordered_gates(self) -> list:
        """Non-empty columns in ascending order; inside a column a stable sort on the first target
        (ties keep insertion order) -- the execution order of the simulator."""
        by_col: dict = {}
        for g in self.gates:
            by_col.setdefault(g.column, []).append(g)
        return [sorted(by_col[c], key=lambda g: g.target_qubits[0])
                for c in sorted(by_col) if c >= 0]

    def compute_layers(self) -> list:
        """Gate indices grouped by column, columns ascending."""
        by_col: dict = {}
        for gi, g in enumerate(self.gates):
            by_col.setdefault(g.column, []).append(gi)
        return [by_col[c] for c in sorted(by_col)]

    def gate_to_layer_map(self) -> list:
        mapping = [0] * len(self.gates)
        for layer, idxs in enumerate(self.compute_layers()):
            for gi in idxs:
                mapping[gi] = layer
        return mapping

    def circuit_hash(self) -> int:
        parts: list = [self.num_qubits, tuple(self.initial_states)]
        parts += [(g.gate_name, tuple(g.target_qubits), tuple(g.params), g.column) for g in self.gates]
        return hash(tuple(parts))

    def gate_count(self) -> int:
        return len(self.gates)

    # -- (de)serialisation ---------------------------------------------------------------
    def to_dict(self) -> dict:
        d = {"version": "1.0", "num_qubits": self.num_qubits, "gates": [g.to_dict() for g in self.gates]}
        if any(self.initial_states):
            d["initial_states"] = self.initial_states
        return d

    @classmethod
    def from_dict(cls, data: dict) -> "QuantumCircuit":
        qc = cls(num_qubits=data["num_qubits"], initial_states=data.get("initial_states", []))
        for g in data["gates"]:
            qc.add_gate(GateInstance.from_dict(g))
        return qc
