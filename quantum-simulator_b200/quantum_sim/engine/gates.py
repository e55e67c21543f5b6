"""Gate matrices and the GateDefinition record (mirror of the reference's gates.py API).

The matrices are kernel *inputs*: the device executor receives them as 2^k x 2^k complex128
blocks (or recognises the structured ones by name, see qsb/compiler.py).  Same names and values
as gates.py:11-134 of the reference so `from quantum_sim.engine.gates import H_MATRIX` keeps working.
"""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Callable

import numpy as np


class GateType(Enum):
    SINGLE = "single"
    CONTROLLED = "controlled"
    MULTI = "multi"
    MEASUREMENT = "measurement"
    BARRIER = "barrier"


@dataclass(frozen=True)
class GateDefinition:
    """Immutable description of one registered gate (gates.py:19-33 of the reference)."""
    name: str
    display_name: str
    gate_type: GateType
    num_qubits: int
    num_params: int
    param_names: tuple
    matrix_func: Callable[..., np.ndarray]
    symbol: str
    color: str
    num_controls: int = 0
    num_targets: int = 1


def _c(rows):
    return np.array(rows, dtype=np.complex128)


def _permutation(dim, a, b):
    m = np.eye(dim, dtype=np.complex128)
    m[[a, b]] = m[[b, a]]
    return m


I_MATRIX = _c([[1, 0], [0, 1]])
X_MATRIX = _c([[0, 1], [1, 0]])
Y_MATRIX = _c([[0, -1j], [1j, 0]])
Z_MATRIX = _c([[1, 0], [0, -1]])
H_MATRIX = _c([[1, 1], [1, -1]]) / np.sqrt(2)
S_MATRIX = _c([[1, 0], [0, 1j]])
S_DAG_MATRIX = _c([[1, 0], [0, -1j]])
T_MATRIX = _c([[1, 0], [0, np.exp(1j * np.pi / 4)]])
T_DAG_MATRIX = _c([[1, 0], [0, np.exp(-1j * np.pi / 4)]])

CNOT_MATRIX = _permutation(4, 2, 3)
CZ_MATRIX = np.diag([1, 1, 1, -1]).astype(np.complex128)
SWAP_MATRIX = _permutation(4, 1, 2)
TOFFOLI_MATRIX = _permutation(8, 6, 7)
FREDKIN_MATRIX = _permutation(8, 5, 6)


def rx_matrix(theta: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return _c([[c, -1j * s], [-1j * s, c]])


def ry_matrix(theta: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return _c([[c, -s], [s, c]])


def rz_matrix(theta: float) -> np.ndarray:
    return _c([[np.exp(-1j * theta / 2), 0], [0, np.exp(1j * theta / 2)]])


def phase_matrix(phi: float) -> np.ndarray:
    return _c([[1, 0], [0, np.exp(1j * phi)]])


def u3_matrix(theta: float, phi: float, lam: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return _c([[c, -np.exp(1j * lam) * s],
               [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]])


def _const(matrix: np.ndarray) -> Callable[[], np.ndarray]:
    return lambda: matrix
