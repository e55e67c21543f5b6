"""Parameterised-circuit optimisation (the reference's optimizer.py) on batched device launches.

Same public names and call signatures as the reference (optimizer.py:28-559): ParameterBinding,
ParameterizedCircuitConfig, CostFunction, GradientEstimator, CircuitOptimizer, OptimizationResult,
BarrenPlateauAnalysis.  What changes is the execution shape: the reference binds every shifted
parameter vector into a deep copy of the circuit and re-simulates it gate by gate
(optimizer.py:66-72, :194-229 -- 2P full runs per gradient).  Here the circuit is lowered ONCE with its
tunable angles left symbolic (QSB_OP_RX/RY/RZ/PHASE/U3 read them from a per-state parameter row), all
2P (or n_samples * 2P) parameter vectors go to the device as one float64[B][P] matrix, and ONE launch of
the tile executor produces the B final states; cost functions built by `CostFunction` are evaluated on the
whole batch with the reference's semantics (observables applied through `apply_gate`, so the axis
scramble of state_vector.py:66-73 is part of <psi|O psi>, analysis.py:222-237).
Adam bookkeeping stays on the host -- it is a handful of length-P vector operations.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

from qsb import runtime
from qsb.compiler import Lowering, _PARAM_COUNT
from qsb.lowering import lower_circuit
from .analysis import StateAnalysis
from .circuit import QuantumCircuit, GateInstance
from .gate_registry import GateRegistry
from .gates import X_MATRIX, Y_MATRIX, Z_MATRIX, I_MATRIX
from .simulator import Simulator, _circuit_key
from .state_vector import StateVector

_CHUNK_BYTES = 8 << 30


# ---- parameter binding ------------------------------------------------------------------------
@dataclass
class ParameterBinding:
    """Maps an optimisation variable to a gate parameter slot."""
    gate_index: int
    param_index: int
    name: str = ""


class ParameterizedCircuitConfig:
    """A circuit with identified tunable parameters (optimizer.py:36-88)."""

    def __init__(self, circuit: QuantumCircuit, bindings: list):
        self._circuit = circuit
        self._bindings = bindings

    @property
    def circuit(self) -> QuantumCircuit:
        return self._circuit

    @property
    def bindings(self) -> list:
        return self._bindings

    @property
    def num_params(self) -> int:
        return len(self._bindings)

    def get_values(self) -> np.ndarray:
        vals = np.zeros(self.num_params)
        for i, b in enumerate(self._bindings):
            vals[i] = self._circuit.gates[b.gate_index].params[b.param_index]
        return vals

    def bind_values(self, values: np.ndarray) -> QuantumCircuit:
        from copy import deepcopy
        qc = deepcopy(self._circuit)
        for i, b in enumerate(self._bindings):
            qc.gates[b.gate_index].params[b.param_index] = float(values[i])
        return qc

    @classmethod
    def auto_detect(cls, circuit: QuantumCircuit) -> "ParameterizedCircuitConfig":
        registry = GateRegistry.instance()
        bindings = []
        for gi, gate in enumerate(circuit.gates):
            try:
                gate_def = registry.get(gate.gate_name)
            except KeyError:
                continue
            if gate_def.num_params > 0:
                for pi in range(gate_def.num_params):
                    bindings.append(ParameterBinding(gi, pi, f"{gate.gate_name}[{gi}].p{pi}"))
        return cls(circuit, bindings)

    # ---- batched execution ---------------------------------------------------------------------
    def _device_program(self):
        """(device program, row layout): the circuit lowered once, bound gates reading angles from a
        parameter row.  A gate is symbolic only when ALL of its parameters are bound and it is one of
        Rx/Ry/Rz/Phase/U3; otherwise `run_batch` falls back to one lowering per parameter vector."""
        qc = self._circuit
        by_gate = {}
        for col, b in enumerate(self._bindings):
            by_gate.setdefault(b.gate_index, {})[b.param_index] = col
        offsets, row_cols = {}, []
        for gi, slots in by_gate.items():
            g = qc.gates[gi]
            k = _PARAM_COUNT.get(g.gate_name)
            if (k is None or sorted(slots) != list(range(k)) or len(g.target_qubits) != 1
                    or not GateRegistry.instance().is_builtin(g.gate_name)):
                return None, None
            offsets[id(g)] = len(row_cols)
            row_cols += [slots[p] for p in range(k)]
        key = ("param-circuit", StateVector.layout, _circuit_key(qc), tuple(sorted((gi, tuple(sorted(s.items())))
                                                                               for gi, s in by_gate.items())))

        def build():
            prog, _ = lower_circuit(qc.num_qubits, qc.get_ordered_gates(), GateRegistry.instance(),
                                    param_offsets=offsets, layout=StateVector.layout)
            return prog

        return runtime.cached_program(key, build), np.array(row_cols, dtype=np.int64)

    def run_batch(self, values: np.ndarray):
        """Final states of the circuit for every row of `values` (float64[B][P]) -> (ctx, device buffer with
        complex128[B][2^n])."""
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, max(self.num_params, 1))[:, :self.num_params]
        B = values.shape[0]
        qc = self._circuit
        n = qc.num_qubits
        c = runtime.ctx()
        states = c.alloc(max(B, 1) * (16 << n))
        dp, row_cols = self._device_program()
        basis = StateVector._basis_index(qc.initial_states)
        if dp is not None:
            rows = np.ascontiguousarray(values[:, row_cols]) if len(row_cols) else np.zeros((B, 1))
            stride = rows.shape[1]
            c.run(dp, B, states=states, params=c.to_device(rows) if dp.prog.n_params else None,
                  params_stride=stride if dp.prog.n_params else 0, default_basis=basis)
            return c, states
        sim = Simulator()                     # unusual bindings: one lowering per parameter vector
        for t in range(B):
            st = sim.run(self.bind_values(values[t]), shots=0).final_state
            states.copy_from(st._device(), 16 << n, dst_off=t * (16 << n))
        return c, states


# ---- cost functions -----------------------------------------------------------------------------
class _Cost:
    """Callable (state) -> float, as the reference's closures; `terms` lets a whole batch of states be
    evaluated with a few launches: cost = offset + sum_k coeff_k * Re <psi| O_k psi>."""

    def __init__(self, terms=None, offset=0.0, target=None, single=None):
        self.terms = terms          # [(coeff, observable matrix, target qubits)]
        self.offset = offset
        self.target = target        # state_fidelity: cost = 1 - |<target|psi>|^2
        self._single = single

    def __call__(self, state: StateVector) -> float:
        if self._single is not None:
            return self._single(state)
        return float(self.batch(state.num_qubits, state._device(), 1)[0])

    def batch(self, n, states, count) -> np.ndarray:
        c = runtime.ctx()
        dim = 1 << n
        if self.target is not None:
            out = c.alloc(count * 16)
            tgt = c.to_device(np.ascontiguousarray(self.target, dtype=np.complex128).reshape(-1))
            c.overlap(n, states, 0, tgt, 0, 0, count, out)     # stride 0: one target for every state
            return 1.0 - np.abs(out.download(np.complex128, (count,))) ** 2
        total = np.full(count, float(self.offset))
        temp = c.alloc(count * dim * 16)
        out = c.alloc(count * 16)
        for coeff, obs, targets in self.terms:
            key = ("observable", StateVector.layout, n, runtime.matrix_key(obs), tuple(targets))

            def build(obs=obs, targets=targets):
                lw = Lowering(n, layout=StateVector.layout)
                lw.matrix(obs, targets)
                return lw.finish()

            dp = runtime.cached_program(key, build)
            c.run(dp, count, states=states, load=True, store=True, states_out=temp)
            c.overlap(n, states, 0, temp, 0, 1, count, out)
            total += coeff * out.download(np.complex128, (count,)).real
        return total


class CostFunction:
    """Static factory methods for cost functions; each returns a callable (state: StateVector) -> float
    (optimizer.py:93-186)."""

    @staticmethod
    def expectation_value(observable: np.ndarray, target_qubits: list) -> Callable:
        return _Cost([(1.0, np.asarray(observable, dtype=np.complex128), list(target_qubits))])

    @staticmethod
    def state_fidelity(target_state: np.ndarray) -> Callable:
        return _Cost(target=np.asarray(target_state, dtype=np.complex128))

    @staticmethod
    def z_expectation(qubit: int) -> Callable:
        return _Cost([(1.0, Z_MATRIX, [qubit])])

    @staticmethod
    def vqe_hamiltonian(terms: list) -> Callable:
        pauli_map = {"I": I_MATRIX, "X": X_MATRIX, "Y": Y_MATRIX, "Z": Z_MATRIX}
        out = []
        for coeff, pauli_str, qubits in terms:
            if len(pauli_str) == 1 and len(qubits) == 1:
                if pauli_str.upper() not in ("X", "Y", "Z"):
                    raise ValueError(f"Unknown Pauli: {pauli_str}. Use 'X', 'Y', or 'Z'.")
                out.append((float(coeff), pauli_map[pauli_str.upper()], list(qubits)))
            else:
                obs = pauli_map[pauli_str[0]]
                for p in pauli_str[1:]:
                    obs = np.kron(obs, pauli_map[p])
                out.append((float(coeff), obs, list(qubits)))
        return _Cost(out)

    @staticmethod
    def qaoa_maxcut(edges: list) -> Callable:
        zz = np.kron(Z_MATRIX, Z_MATRIX)
        return _Cost([(-0.5, zz, [i, j]) for i, j in edges], offset=0.5 * len(edges))


def batch_costs(config: ParameterizedCircuitConfig, cost_fn: Callable, values: np.ndarray) -> np.ndarray:
    """cost_fn(final state) for every row of `values`, batched on the device."""
    values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, max(config.num_params, 1))
    n = config.circuit.num_qubits
    per = max(1, _CHUNK_BYTES // (3 * (16 << n)))
    out = np.empty(values.shape[0])
    for lo in range(0, values.shape[0], per):
        rows = values[lo:lo + per]
        c, states = config.run_batch(rows)
        if isinstance(cost_fn, _Cost) and cost_fn._single is None:
            out[lo:lo + len(rows)] = cost_fn.batch(n, states, len(rows))
        else:                                   # user closure: hand it one StateVector view per state
            for t in range(len(rows)):
                buf = c.alloc(16 << n)
                buf.copy_from(states, 16 << n, src_off=t * (16 << n))
                out[lo + t] = cost_fn(StateVector._from_device(n, buf))
    return out


def batch_costs_sharded(config: ParameterizedCircuitConfig, cost_fn: Callable, values: np.ndarray,
                        _batch_costs=None) -> np.ndarray:
    """`batch_costs` with the parameter rows split over the ranks of the default process group (one process per
    GPU): contiguous row ranges, no data-path collective, one gather of the costs; every rank returns the full
    vector, equal to the single-process one.  `_batch_costs` lets the CPU gloo test stand in for the device."""
    from qsb import distributed as D
    values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, max(config.num_params, 1))
    world, rank = D.world_info()
    lo, hi = D.shard_bounds(values.shape[0], world, rank)
    fn = _batch_costs or batch_costs
    mine = np.asarray(fn(config, cost_fn, values[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    return D.gather_concat(mine)


# ---- gradient estimation ------------------------------------------------------------------------
class GradientEstimator:
    """Gradients of parameterised circuits (optimizer.py:191-258): all shifted circuits in one batch."""

    @staticmethod
    def _shifted(values, delta):
        values = np.asarray(values, dtype=np.float64)
        P = len(values)
        rows = np.repeat(values[None, :], 2 * P, axis=0)
        idx = np.arange(P)
        rows[2 * idx, idx] += delta
        rows[2 * idx + 1, idx] -= delta
        return rows

    @staticmethod
    def parameter_shift(config, cost_fn, values, shift: float = np.pi / 2, seed=None) -> np.ndarray:
        """grad_i = [f(theta_i + shift) - f(theta_i - shift)] / (2 sin(shift))."""
        values = np.asarray(values, dtype=np.float64)
        if len(values) == 0:
            return np.zeros(0)
        costs = batch_costs(config, cost_fn, GradientEstimator._shifted(values, shift))
        return (costs[0::2] - costs[1::2]) * (1.0 / (2.0 * np.sin(shift)))

    @staticmethod
    def finite_difference(config, cost_fn, values, epsilon: float = 1e-4, seed=None) -> np.ndarray:
        values = np.asarray(values, dtype=np.float64)
        if len(values) == 0:
            return np.zeros(0)
        costs = batch_costs(config, cost_fn, GradientEstimator._shifted(values, epsilon))
        return (costs[0::2] - costs[1::2]) / (2 * epsilon)


# ---- Adam optimiser -------------------------------------------------------------------------------
@dataclass
class BarrenPlateauAnalysis:
    per_layer_variance: list
    per_layer_mean_variance: list
    per_qubit_variance: list
    depth_scaling: list
    overall_mean_variance: float
    overall_is_barren: bool
    threshold: float
    n_samples: int
    param_layer_map: list


@dataclass
class OptimizationResult:
    optimal_values: np.ndarray
    optimal_cost: float
    history: list
    converged: bool
    iterations: int


class CircuitOptimizer:
    """Adam optimiser for parameterised circuits (optimizer.py:290-421)."""

    def __init__(self, config: ParameterizedCircuitConfig, cost_fn: Callable, learning_rate: float = 0.1,
                 beta1: float = 0.9, beta2: float = 0.999, max_iterations: int = 100, tolerance: float = 1e-6,
                 gradient_method: str = "parameter_shift"):
        self._config = config
        self._cost_fn = cost_fn
        self._lr = learning_rate
        self._beta1 = beta1
        self._beta2 = beta2
        self._max_iter = max_iterations
        self._tol = tolerance
        self._grad_method = gradient_method
        n = config.num_params
        self._values = config.get_values().copy()
        self._m = np.zeros(n)
        self._v = np.zeros(n)
        self._t = 0
        self._history = []
        self._stop_requested = False

    @property
    def values(self) -> np.ndarray:
        return self._values.copy()

    @property
    def history(self) -> list:
        return self._history

    def request_stop(self) -> None:
        self._stop_requested = True

    def step(self, seed=None):
        self._t += 1
        if self._grad_method == "parameter_shift":
            grad = GradientEstimator.parameter_shift(self._config, self._cost_fn, self._values, seed=seed)
        else:
            grad = GradientEstimator.finite_difference(self._config, self._cost_fn, self._values, seed=seed)
        self._m = self._beta1 * self._m + (1 - self._beta1) * grad
        self._v = self._beta2 * self._v + (1 - self._beta2) * grad ** 2
        m_hat = self._m / (1 - self._beta1 ** self._t)
        v_hat = self._v / (1 - self._beta2 ** self._t)
        self._values -= self._lr * m_hat / (np.sqrt(v_hat) + 1e-8)
        cost = float(batch_costs(self._config, self._cost_fn, self._values[None, :])[0])
        self._history.append((self._values.copy(), cost))
        return self._values.copy(), cost

    def run(self, callback=None, seed=None) -> OptimizationResult:
        self._stop_requested = False
        converged = False
        for i in range(self._max_iter):
            if self._stop_requested:
                break
            values, cost = self.step(seed=seed)
            if callback is not None:
                callback(i, values, cost)
            if len(self._history) >= 2 and abs(cost - self._history[-2][1]) < self._tol:
                converged = True
                break
        best = min(range(len(self._history)), key=lambda k: self._history[k][1])
        return OptimizationResult(optimal_values=self._history[best][0], optimal_cost=self._history[best][1],
                                  history=self._history, converged=converged, iterations=len(self._history))

    # ---- barren plateaus: n_samples * 2P circuits in one batch (optimizer.py:423-559) ------------------
    def _gradient_samples(self, n_samples, seed):
        rng = np.random.default_rng(seed)
        P = self._config.num_params
        points = np.empty((n_samples, P))
        for s in range(n_samples):
            points[s] = rng.uniform(-np.pi, np.pi, size=P)
            rng.integers(0, 2 ** 63)            # the reference draws a per-sample seed here (unused by shots=0 runs)
        if P == 0:
            return np.zeros((n_samples, 0))
        shift = np.pi / 2
        rows = np.concatenate([GradientEstimator._shifted(points[s], shift) for s in range(n_samples)])
        costs = batch_costs(self._config, self._cost_fn, rows).reshape(n_samples, 2 * P)
        return (costs[:, 0::2] - costs[:, 1::2]) * (1.0 / (2.0 * np.sin(shift)))

    def detect_barren_plateau(self, n_samples: int = 50, seed=None) -> dict:
        grads = self._gradient_samples(n_samples, seed)
        per_param_var = np.var(grads, axis=0)
        mean_var = float(np.mean(per_param_var))
        return {"mean_variance": mean_var, "per_param": per_param_var.tolist(), "is_barren": mean_var < 1e-4}

    def detect_barren_plateau_layered(self, n_samples: int = 50, seed=None) -> BarrenPlateauAnalysis:
        circuit = self._config.circuit
        g2l = circuit.gate_to_layer_map()
        param_layer_map, param_qubit_map = [], []
        for b in self._config.bindings:
            gate = circuit.gates[b.gate_index]
            param_layer_map.append(g2l[b.gate_index])
            param_qubit_map.append(gate.target_qubits[0] if gate.target_qubits else 0)
        per_param_var = np.var(self._gradient_samples(n_samples, seed), axis=0)
        layer_indices, qubit_indices = {}, {}
        for pi, layer in enumerate(param_layer_map):
            layer_indices.setdefault(layer, []).append(pi)
        for pi, q in enumerate(param_qubit_map):
            qubit_indices.setdefault(q, []).append(pi)
        per_layer_variance, per_layer_mean, depth_scaling = [], [], []
        for layer in sorted(layer_indices):
            vs = [float(per_param_var[pi]) for pi in layer_indices[layer]]
            per_layer_variance.append(vs)
            per_layer_mean.append(float(np.mean(vs)))
            depth_scaling.append((layer, float(np.mean(vs))))
        max_qubit = max(qubit_indices) if qubit_indices else 0
        per_qubit = [float(np.mean([per_param_var[pi] for pi in qubit_indices[q]])) if q in qubit_indices else 0.0
                     for q in range(max_qubit + 1)]
        overall = float(np.mean(per_param_var))
        return BarrenPlateauAnalysis(per_layer_variance=per_layer_variance, per_layer_mean_variance=per_layer_mean,
                                     per_qubit_variance=per_qubit, depth_scaling=depth_scaling,
                                     overall_mean_variance=overall, overall_is_barren=overall < 1e-4, threshold=1e-4,
                                     n_samples=n_samples, param_layer_map=param_layer_map)
