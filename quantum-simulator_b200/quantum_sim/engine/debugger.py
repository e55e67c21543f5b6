"""Circuit debugger (the reference's debugger.py) on batched device launches.

Same public names as the reference (debugger.py:20-528): DebugSnapshot, NoiseImpactResult, NoiseAttribution,
CircuitDebugger.  The reference walks the circuit column by column with two StateVectors per trial and copies
them after every column; here the ideal run and ALL noisy trials are one launch each of the tile executor with a
snapshot per column (`record_steps` semantics, simulator.py:70-71), and the per-column metrics are device
reductions over the snapshots: overlaps for the fidelities, all 1-qubit reduced density matrices in one call for
the per-qubit Uhlmann fidelities (the 2x2 square roots stay on the host, as in the reference).

`entropy` fields: the reference computes `von_neumann_entropy` of a PURE state through a 2^n x 2^n eigvalsh and
gets rounding noise of order 1e-16 (golden file: -6.4e-16 ... 3.2e-16); the exact value 0.0 is reported here.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from qsb import runtime
from .analysis import StateAnalysis
from .circuit import QuantumCircuit
from .gate_registry import GateRegistry
from .gates import GateType
from .simulator import Simulator
from .state_vector import StateVector


@dataclass
class DebugSnapshot:
    """State captured at a single execution point."""
    column_index: int
    state: StateVector
    ideal_state: StateVector | None
    gate_labels: list
    fidelity: float
    cumulative_fidelity: float
    entropy: float


@dataclass
class NoiseImpactResult:
    """Noise impact for a single gate column."""
    column_index: int
    gate_labels: list
    fidelity_before: float
    fidelity_after: float
    fidelity_drop: float
    entropy_before: float
    entropy_after: float
    entropy_change: float
    per_qubit_fidelity: list
    mean_delta_fidelity: float = 0.0
    std_delta_fidelity: float = 0.0


@dataclass
class NoiseAttribution:
    """Per-gate noise attribution (debugger.py:50-69)."""
    delta_fidelity: list
    delta_fidelity_std: list
    total_fidelity_loss: float
    column_attribution_pct: list
    per_qubit_attribution: list
    gate_labels: list
    is_recovery: list = field(default_factory=list)
    no_measurable_loss: bool = False


class CircuitDebugger:
    """Caches per-column states for forward/backward stepping; noise impact / attribution over trials."""

    def __init__(self):
        self._snapshots = []
        self._position = 0
        self._breakpoints = set()
        self._registry = GateRegistry.instance()

    # ---- helpers -------------------------------------------------------------------------------------------
    def _labels(self, circuit):
        out = []
        for column_gates in circuit.get_ordered_gates():
            labels = []
            for g in column_gates:
                gd = self._registry.get(g.gate_name)
                if gd.gate_type not in (GateType.MEASUREMENT, GateType.BARRIER):
                    labels.append(f"{g.gate_name}({','.join(str(q) for q in g.target_qubits)})")
            out.append(labels)
        return out

    @staticmethod
    def _column_snapshots(circuit, noise_model, uniforms, count):
        """Device buffer complex128[count][columns][2^n] with the state after every column of `count` runs."""
        sim = Simulator(noise_model)
        dp, _ = sim._program(circuit, record_steps=True)
        c = runtime.ctx()
        dim = 2 ** circuit.num_qubits
        ns = dp.prog.n_snapshots
        snaps = c.alloc(max(count * ns, 1) * dim * 16)
        kw = {}
        if dp.prog.n_draws:
            kw.update(uniforms=c.to_device(np.ascontiguousarray(uniforms, dtype=np.float64)), uniforms_stride=dp.prog.n_draws)
        if ns:
            c.run(dp, count, default_basis=sim._basis(circuit), snapshots=snaps, store=False, **kw)
        return c, snaps, ns, dp.prog.n_draws

    # ---- stepping ----------------------------------------------------------------------------------------------
    def run_full_debug(self, circuit: QuantumCircuit, noise_model=None, seed=None) -> list:
        """Execute the circuit and cache the state after every column (debugger.py:94-174)."""
        self._snapshots.clear()
        self._position = 0
        n = circuit.num_qubits
        dim = 2 ** n
        labels = self._labels(circuit)
        initial = StateVector.from_initial_states(circuit.initial_states)
        self._snapshots.append(DebugSnapshot(column_index=-1, state=initial.copy(),
                                             ideal_state=initial.copy() if noise_model else None, gate_labels=[],
                                             fidelity=1.0, cumulative_fidelity=1.0, entropy=0.0))
        if not labels:
            return self._snapshots
        c, ideal, ns, _ = self._column_snapshots(circuit, None, None, 1)
        ideal_host = ideal.download(np.complex128, (ns, dim))
        if noise_model is None:
            for col in range(ns):
                self._snapshots.append(DebugSnapshot(column_index=col, state=StateVector._from_host(n, ideal_host[col].copy()),
                                                     ideal_state=None, gate_labels=labels[col], fidelity=1.0,
                                                     cumulative_fidelity=1.0, entropy=0.0))
            return self._snapshots
        d = Simulator(noise_model)._program(circuit, record_steps=True)[0].prog.n_draws
        draws = noise_model._rng.random(d) if d else None                     # the model's own stream, as noise.apply draws
        _, noisy, _, _ = self._column_snapshots(circuit, noise_model, draws[None] if d else None, 1)
        noisy_host = noisy.download(np.complex128, (ns, dim))
        ov = c.alloc(ns * 16)
        c.overlap(n, ideal, 0, noisy, 0, 1, ns, ov)
        fid = np.abs(ov.download(np.complex128, (ns,))) ** 2
        c.overlap(n, noisy, 0, initial._device(), 0, 0, ns, ov)
        cum = np.abs(ov.download(np.complex128, (ns,))) ** 2
        for col in range(ns):
            self._snapshots.append(DebugSnapshot(column_index=col, state=StateVector._from_host(n, noisy_host[col].copy()),
                                                 ideal_state=StateVector._from_host(n, ideal_host[col].copy()),
                                                 gate_labels=labels[col], fidelity=float(fid[col]),
                                                 cumulative_fidelity=float(cum[col]), entropy=0.0))
        return self._snapshots

    @property
    def snapshots(self) -> list:
        return self._snapshots

    @property
    def position(self) -> int:
        return self._position

    @position.setter
    def position(self, value: int) -> None:
        if self._snapshots:
            self._position = max(0, min(value, len(self._snapshots) - 1))

    @property
    def current_snapshot(self):
        return self._snapshots[self._position] if self._snapshots else None

    @property
    def num_steps(self) -> int:
        return len(self._snapshots)

    def step_forward(self):
        if not self._snapshots or self._position >= len(self._snapshots) - 1:
            return None
        self._position += 1
        return self._snapshots[self._position]

    def step_backward(self):
        if not self._snapshots or self._position <= 0:
            return None
        self._position -= 1
        return self._snapshots[self._position]

    def goto_step(self, step: int):
        if not self._snapshots:
            return None
        self._position = max(0, min(step, len(self._snapshots) - 1))
        return self._snapshots[self._position]

    # ---- breakpoints ----------------------------------------------------------------------------------------------
    def add_breakpoint(self, column: int) -> None:
        self._breakpoints.add(column)

    def remove_breakpoint(self, column: int) -> None:
        self._breakpoints.discard(column)

    def toggle_breakpoint(self, column: int) -> bool:
        if column in self._breakpoints:
            self._breakpoints.discard(column)
            return False
        self._breakpoints.add(column)
        return True

    @property
    def breakpoints(self) -> set:
        return self._breakpoints

    def clear_breakpoints(self) -> None:
        self._breakpoints.clear()

    def run_to_breakpoint(self):
        if not self._snapshots:
            return None
        for i in range(self._position + 1, len(self._snapshots)):
            if self._snapshots[i].column_index in self._breakpoints:
                self._position = i
                return self._snapshots[i]
        self._position = len(self._snapshots) - 1
        return self._snapshots[self._position]

    # ---- trials as one batch ------------------------------------------------------------------------------------------
    def _trial_metrics(self, circuit, noise_model, n_trials, seed):
        """fid[trial][col] = |<ideal_col|noisy_col>|^2 and pq[trial][col][qubit] = Uhlmann fidelity of the 1-qubit
        reduced density matrices, for all trials in one batch.  Trial t runs with the noise generator seeded with
        child seed t of `seed` (debugger.py:300-301)."""
        n = circuit.num_qubits
        base_rng = np.random.default_rng(seed)
        seeds = [int(base_rng.integers(0, 2 ** 63)) for _ in range(n_trials)]
        c, ideal, ns, _ = self._column_snapshots(circuit, None, None, 1)
        d = Simulator(noise_model)._program(circuit, record_steps=True)[0].prog.n_draws
        uniforms = np.empty((n_trials, max(d, 1)))
        for t in range(n_trials):
            if d:
                uniforms[t] = np.random.default_rng(seeds[t]).random(d)
        _, noisy, _, _ = self._column_snapshots(circuit, noise_model, uniforms if d else None, n_trials)
        fid = np.empty((n_trials, ns))
        ov = c.alloc(max(ns, 1) * 16)
        for t in range(n_trials):
            c.overlap(n, noisy, t * ns, ideal, 0, 1, ns, ov)
            fid[t] = np.abs(ov.download(np.complex128, (ns,))) ** 2
        r_ideal = c.alloc(max(ns, 1) * n * 4 * 16)
        r_noisy = c.alloc(max(n_trials * ns, 1) * n * 4 * 16)
        c.rdm_all(n, ideal, 0, ns, r_ideal, None)
        c.rdm_all(n, noisy, 0, n_trials * ns, r_noisy, None)
        ri = r_ideal.download(np.complex128, (1, ns, n, 2, 2))
        rn = r_noisy.download(np.complex128, (n_trials, ns, n, 2, 2))
        pq = StateAnalysis.density_fidelity(np.broadcast_to(ri, rn.shape), rn)
        if n_trials:
            noise_model.set_seed(seeds[-1])          # the reference leaves the model seeded with the last child seed ...
            if d:
                noise_model._rng.random(d)           # ... and advanced by one trial's draws
        return fid, pq, ns

    def compute_noise_impact(self, circuit: QuantumCircuit, noise_model, n_trials: int = 50, seed=None) -> list:
        """Per-column fidelity drop due to noise, averaged over trials (debugger.py:261-364)."""
        if noise_model is None:
            return []
        labels = self._labels(circuit)
        if not labels:
            return []
        fid, pq, ns = self._trial_metrics(circuit, noise_model, n_trials, seed)
        before = np.concatenate([np.ones((n_trials, 1)), fid[:, :-1]], axis=1)   # state before column c = after column c-1
        drops = before - fid
        results = []
        for col in range(ns):
            fb = fa = 0.0
            pqf = np.zeros(circuit.num_qubits)
            for t in range(n_trials):                 # trial-by-trial accumulation, as the reference sums
                fb += before[t, col]
                fa += fid[t, col]
                pqf = pqf + pq[t, col]
            fb /= n_trials
            fa /= n_trials
            results.append(NoiseImpactResult(column_index=col, gate_labels=labels[col], fidelity_before=float(fb),
                                             fidelity_after=float(fa), fidelity_drop=float(fb - fa), entropy_before=0.0,
                                             entropy_after=0.0, entropy_change=0.0,
                                             per_qubit_fidelity=(pqf / n_trials).tolist(),
                                             mean_delta_fidelity=float(np.mean(drops[:, col])),
                                             std_delta_fidelity=float(np.std(drops[:, col]))))
        return results

    def compute_noise_attribution(self, circuit: QuantumCircuit, noise_model, reference_state=None, n_trials: int = 50,
                                  seed=None) -> NoiseAttribution:
        """Per-column growth of the fidelity gap between the ideal and the noisy trajectory (debugger.py:366-477)."""
        labels = self._labels(circuit)
        num_cols = len(labels)
        n_qubits = circuit.num_qubits
        if num_cols == 0:
            return NoiseAttribution([], [], 0.0, [], [], [], [], True)
        fid, pq, ns = self._trial_metrics(circuit, noise_model, n_trials, seed)
        gap = 1.0 - fid
        contrib = gap - np.concatenate([np.zeros((n_trials, 1)), gap[:, :-1]], axis=1)
        mean_contrib = np.mean(contrib, axis=0).tolist()
        std_contrib = np.std(contrib, axis=0).tolist()
        total_loss = float(np.sum(mean_contrib))
        is_recovery = [d < -1e-12 for d in mean_contrib]
        positive_sum = sum(max(0.0, d) for d in mean_contrib)
        no_loss = positive_sum <= 1e-12
        attr_pct = [max(0.0, d) / positive_sum * 100.0 for d in mean_contrib] if not no_loss else [0.0] * num_cols
        acc = np.zeros((num_cols, n_qubits))
        for t in range(n_trials):
            acc += 1.0 - pq[t]
        return NoiseAttribution(delta_fidelity=mean_contrib, delta_fidelity_std=std_contrib, total_fidelity_loss=total_loss,
                                column_attribution_pct=attr_pct, per_qubit_attribution=(acc / n_trials).tolist(),
                                gate_labels=labels, is_recovery=is_recovery, no_measurable_loss=no_loss)

    # ---- state diff ---------------------------------------------------------------------------------------------------
    @staticmethod
    def compute_state_diff(snap_a: DebugSnapshot, snap_b: DebugSnapshot) -> dict:
        """Compare two snapshots (debugger.py:480-528): fidelity and probabilities on the device, the top-10 listing
        on the host."""
        n = snap_a.state.num_qubits
        fid = StateAnalysis.process_fidelity(snap_a.state, snap_b.state)
        prob_a, prob_b = snap_a.state.probabilities, snap_b.state.probabilities
        data_a, data_b = snap_a.state.data, snap_b.state.data
        amp_diffs = np.abs(data_a - data_b)
        amplitude_diffs = []
        for idx in np.argsort(amp_diffs)[::-1][:min(10, len(amp_diffs))]:
            if amp_diffs[idx] < 1e-10:
                break
            amplitude_diffs.append((int(idx), format(idx, f"0{n}b"), complex(data_a[idx]), complex(data_b[idx]),
                                    float(amp_diffs[idx])))
        return {"fidelity": float(fid), "tvd": float(0.5 * np.sum(np.abs(prob_a - prob_b))),
                "amplitude_diffs": amplitude_diffs, "entropy_diff": snap_b.entropy - snap_a.entropy,
                "prob_diffs": np.abs(prob_a - prob_b)}
