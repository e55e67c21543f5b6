"""Simulator -- the reference's run entry points (simulator.py:28-199) as single device launches.

`run`, `run_with_noise` and `ensemble_density_matrix` each lower the whole circuit (and the noise model)
to one program and execute it for 1 / shots / n_trials states in one launch of the trajectory kernel;
the per-gate Python loop of the reference disappears.  Random streams are the reference's: Kraus draws
from `NoiseModel._rng` (one double per draw, gate-major / channel / target order), child seeds via
`rng.integers(0, 2**63)`, `measure_all` via one `rng.random()` per shot.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Generator

import numpy as np

from qsb import runtime
from qsb.lowering import lower_circuit
from .circuit import QuantumCircuit, GateInstance
from .gate_registry import GateRegistry
from .gates import GateType
from .measurement import MeasurementEngine, MeasurementBasis
from .state_vector import StateVector

_PIPE_SHOTS = (64, 2048)        # first / largest pipelined slice of run_with_noise, in trajectories
_CHUNK_BYTES = 8 << 30          # state bytes resident per launch when batching shots / trials


@dataclass
class SimulationResult:
    final_state: StateVector
    measurement_counts: dict
    step_states: list | None = None
    num_shots: int = 1024
    seed: int | None = None
    reference_state: StateVector | None = None


def _circuit_key(circuit):
    # registry generation: any GateRegistry.register() after the built-ins invalidates cached programs (a custom gate
    # re-registered with a new matrix, or a built-in name overridden, must never hit a stale program)
    return (circuit.num_qubits, GateRegistry.instance().generation,
            tuple((g.gate_name, tuple(g.target_qubits), tuple(float(p) for p in g.params), g.column)
                  for g in circuit.gates))


class Simulator:
    """Executes a QuantumCircuit; optional NoiseModel applied after every gate.

    Two backend switches beyond the reference's signature ("config / flags"):
      precision = "c128" (default: the reference's complex128, amplitudes to 1e-12) | "c64" (complex64 states on the
                  device, 2^14-amplitude tiles, tolerance 1e-5; results come back as complex128 arrays)
      rng_mode  = "reference" (default: Kraus draws are taken from NoiseModel._rng exactly like the reference's
                  `choice` calls, so seeded counts are bit-identical) | "philox" (counter-based Philox4x32-10 in the
                  kernel: draw d of trajectory t = f(philox_seed, t, d); no host generator, no upload -- the
                  throughput mode; trajectories are numbered consecutively over the calls of this Simulator)."""

    def __init__(self, noise_model=None, precision="c128", rng_mode="reference", philox_seed=0):
        if precision not in ("c128", "c64"):
            raise ValueError("precision must be 'c128' or 'c64'")
        if rng_mode not in ("reference", "philox"):
            raise ValueError("rng_mode must be 'reference' or 'philox'")
        self._gate_registry = GateRegistry.instance()
        self._noise_model = noise_model
        self._precision = precision
        self._rng_mode = rng_mode
        self._philox_seed = int(philox_seed)
        self._philox_next = 0            # global index of the next trajectory (Philox counter)

    # ---- backend plumbing ------------------------------------------------------------------
    def _ctx(self):
        return runtime.ctx(self._precision)

    def _amp(self):
        return (8, np.complex64) if self._precision == "c64" else (16, np.complex128)

    def _draw_kwargs(self, c, dp, count, uniforms=None):
        """Kraus draws of `count` trajectories: the reference's stream (uploaded) or the in-kernel Philox stream."""
        d = dp.prog.n_draws
        if not d:
            return {}
        if self._rng_mode == "philox":
            kw = dict(seed=self._philox_seed, traj_offset=self._philox_next)
            self._philox_next += count
            return kw
        if uniforms is None:
            uniforms = self._noise_model._rng.random(count * d).reshape(count, d)
        return dict(uniforms=c.to_device(uniforms), uniforms_stride=d)

    def _host_states(self, buf, shape):
        ab, dt = self._amp()
        arr = buf.download(dt, shape)
        return arr if dt is np.complex128 else arr.astype(np.complex128)

    # ---- lowering ------------------------------------------------------------------------
    def _program(self, circuit, record_steps=False, with_noise=True):
        nm = self._noise_model if with_noise else None
        n = circuit.num_qubits
        layout = StateVector.layout
        key = ("circuit", layout, _circuit_key(circuit), record_steps,
               nm._signature() if nm is not None and hasattr(nm, "_signature") else None)
        has_meas = [False]
        max_bits = 14 if self._precision == "c64" else 13       # complex64 tiles hold twice the amplitudes

        def build():
            channels = (lambda name: nm._channel_specs(name)) if nm is not None else None
            prog, hm = lower_circuit(n, circuit.get_ordered_gates(), self._gate_registry, channels,
                                     record_steps=record_steps, layout=layout, max_local_bits=max_bits)
            prog.meta["has_measurement"] = hm
            return prog

        dp = runtime.cached_program(key, build, self._precision)
        return dp, dp.prog.meta["has_measurement"]

    @staticmethod
    def _basis(circuit) -> int:
        return StateVector._basis_index(circuit.initial_states)

    # ---- reference API ---------------------------------------------------------------------
    def run(self, circuit: QuantumCircuit, shots: int = 1024, record_steps: bool = False,
            seed: int | None = None, rng: np.random.Generator | None = None,
            measurement_basis: MeasurementBasis = MeasurementBasis.Z) -> SimulationResult:
        if rng is None:
            rng = np.random.default_rng(seed)
        n = circuit.num_qubits
        if n != len(circuit.initial_states):
            n = len(circuit.initial_states)       # from_initial_states sizes the register (simulator.py:53)
        dp, has_meas = self._program(circuit, record_steps)
        c = self._ctx()
        ab, _ = self._amp()
        dim = 2 ** n
        state_buf = c.alloc(ab * dim)
        kw = self._draw_kwargs(c, dp, 1)
        snaps = None
        if dp.prog.n_snapshots:
            snaps = c.alloc(dp.prog.n_snapshots * dim * ab)
            kw.update(snapshots=snaps)
        c.run(dp, 1, states=state_buf, default_basis=self._basis(circuit), **kw)
        state = (StateVector._from_device(n, state_buf) if self._precision == "c128"
                 else StateVector._from_host(n, self._host_states(state_buf, (dim,))))
        step_states = None
        if record_steps:
            step_states = []
            if snaps is not None:
                host = self._host_states(snaps, (dp.prog.n_snapshots, dim))
                step_states = [StateVector._from_host(n, host[i].copy()) for i in range(dp.prog.n_snapshots)]

        if has_meas or shots > 0:
            readout_err = getattr(self._noise_model, "readout_error", None) if self._noise_model is not None else None
            counts = MeasurementEngine.sample_with_basis(state, shots, basis=measurement_basis,
                                                         readout_error=readout_err, rng=rng)
        else:
            counts = {}
        return SimulationResult(final_state=state, measurement_counts=counts, step_states=step_states,
                                num_shots=shots, seed=seed)

    def run_step_by_step(self, circuit: QuantumCircuit,
                         rng: np.random.Generator | None = None) -> Generator:
        """Yields (state, column_index) after each column, starting with (initial, -1)  (simulator.py:93-108).

        Lazy like the reference: a column is applied when the caller asks for it, gate by gate through
        `_apply_gate_instance` + `NoiseModel.apply` (one launch each on the device-resident state), so the noise
        model's draws are taken from `NoiseModel._rng` column by column, and an unknown gate name in a late column
        raises only when that column is reached, after the earlier columns were yielded.  Complex64 and Philox
        simulators keep the one-launch form (`_steps_one_launch`): those modes have no reference stream to follow."""
        if self._precision != "c128" or self._rng_mode != "reference":
            yield from self._steps_one_launch(circuit)
            return
        state = StateVector.from_initial_states(circuit.initial_states)
        yield state.copy(), -1
        for col_idx, column_gates in enumerate(circuit.get_ordered_gates()):
            for gate_inst in column_gates:
                gate_def = self._gate_registry.get(gate_inst.gate_name)
                if gate_def.gate_type in (GateType.MEASUREMENT, GateType.BARRIER):
                    continue
                self._apply_gate_instance(state, gate_inst)
                if self._noise_model is not None:
                    self._noise_model.apply(state, gate_inst)
            yield state.copy(), col_idx

    def _steps_one_launch(self, circuit: QuantumCircuit) -> Generator:
        """All columns in one launch with a snapshot per column (eager: every draw is taken up front)."""
        yield StateVector.from_initial_states(circuit.initial_states), -1
        n = circuit.num_qubits
        dp, _ = self._program(circuit, record_steps=True)
        if dp.prog.n_snapshots == 0:
            return
        c = self._ctx()
        ab, _ = self._amp()
        dim = 2 ** n
        kw = self._draw_kwargs(c, dp, 1)
        snaps = c.alloc(dp.prog.n_snapshots * dim * ab)
        c.run(dp, 1, default_basis=self._basis(circuit), snapshots=snaps, **kw)
        host = self._host_states(snaps, (dp.prog.n_snapshots, dim))
        for i in range(dp.prog.n_snapshots):
            yield StateVector._from_host(n, host[i].copy()), i

    def _apply_gate_instance(self, state: StateVector, gate: GateInstance):
        gate_def = self._gate_registry.get(gate.gate_name)
        state.apply_gate(gate_def.matrix_func(*gate.params), gate.target_qubits)

    def _trajectory_batch(self, circuit, uniforms, count):
        """Run `count` trajectories of `circuit` (+ noise) and return (ctx, device states buffer)."""
        dp, _ = self._program(circuit)
        c = self._ctx()
        ab, _ = self._amp()
        dim = 2 ** circuit.num_qubits
        states = c.alloc(count * dim * ab)
        kw = self._draw_kwargs(c, dp, count, uniforms) if (uniforms is not None or self._rng_mode == "philox") else {}
        c.run(dp, count, states=states, default_basis=self._basis(circuit), **kw)
        return c, states

    def run_with_noise(self, circuit: QuantumCircuit, shots: int = 1024, seed: int | None = None,
                       rng: np.random.Generator | None = None) -> SimulationResult:
        """`shots` independent noisy trajectories, one basis index each."""
        if self._noise_model is None:
            return self.run(circuit, shots, seed=seed, rng=rng)
        if rng is None:
            rng = np.random.default_rng(seed)
        n = circuit.num_qubits
        idx = self._noisy_indices(circuit, shots, self._noise_model._rng, rng)
        from qsb.distributed import merge_counts_in_shot_order
        counts = merge_counts_in_shot_order(idx, n)            # keys in order of first occurrence (simulator.py:144-145)
        final_state = StateVector.from_initial_states(circuit.initial_states)   # placeholder, as in the reference
        return SimulationResult(final_state=final_state, measurement_counts=counts, num_shots=shots, seed=seed)

    def _noisy_indices(self, circuit, shots, noise_rng, meas_rng) -> np.ndarray:
        """Sampled basis index of `shots` noisy trajectories, in shot order (the device leg of run_with_noise and of a
        rank's share of run_with_noise_sharded).  `noise_rng` hands out d doubles per shot in shot order, never
        reseeded (noise.py:253-259); `meas_rng` one double per shot (state_vector.py:107-113)."""
        n = circuit.num_qubits
        dim = 2 ** n
        dp, _ = self._program(circuit)
        d = dp.prog.n_draws
        ab, _ = self._amp()
        chunk = max(1, min(shots, _CHUNK_BYTES // (ab * dim)))
        c = self._ctx()
        philox = self._rng_mode == "philox"
        # Draws are generated straight into pinned staging memory in slices of _PIPE_SHOTS trajectories and
        # uploaded without a host wait, so the generator works on slice k+1 while the GPU runs slice k.
        # Slice sizes grow 8x (draws are ~20x cheaper than trajectories), so few launches pay a ragged last wave.
        sub = max(1, min(chunk, _PIPE_SHOTS[1]))
        stage = [c.staging(("run_with_noise", k), (sub, max(d, 1))) for k in range(2)] if not philox else None
        dev_u = [c.alloc(sub * max(d, 1) * 8) for _ in range(2)] if not philox else None
        done = [None, None]
        states = c.alloc(chunk * dim * ab)
        basis = self._basis(circuit)
        out_idx = np.empty(shots, dtype=np.int64)
        for lo in range(0, shots, chunk):
            cnt = min(chunk, shots - lo)
            s0, m, j = 0, min(_PIPE_SHOTS[0], sub), 0
            while s0 < cnt:
                k, m = j & 1, min(m, cnt - s0)
                kw = {}
                if d and philox:
                    kw.update(seed=self._philox_seed, traj_offset=self._philox_next)
                    self._philox_next += m
                elif d:
                    if done[k] is not None:
                        done[k].wait()                       # slice j-2 has left this staging buffer
                    noise_rng.random(out=stage[k][:m].reshape(-1))
                    dev_u[k].upload_async(stage[k][:m])
                    done[k] = (done[k] or c.event()).record()
                    kw.update(uniforms=dev_u[k], uniforms_stride=d)
                c.run(dp, m, states=states, first=s0, default_basis=basis, async_=True, **kw)
                s0, m, j = s0 + m, min(8 * m, sub), j + 1
            u = c.to_device(meas_rng.random(cnt))
            out = c.alloc(cnt * 8)
            c.sample_index(n, states, 0, cnt, u, out)
            out.download(np.int64, (cnt,), out=out_idx[lo:lo + cnt])
        return out_idx

    # ---- sharded over the ranks of the default process group (one process per GPU, torchrun) ---------------------
    def run_with_noise_sharded(self, circuit: QuantumCircuit, shots: int = 1024, seed: int | None = None,
                               rng: np.random.Generator | None = None, _indices_fn=None) -> SimulationResult:
        """`run_with_noise` with the shots split over the ranks.  Both generators are POSITIONED, not replayed: rank r
        jumps the noise stream over the lo*d doubles and the measurement stream over the lo doubles that earlier shots
        consume (PCG64 `advance`), so every shot draws exactly what the reference's single loop gives it
        (simulator.py:134-145); the per-shot indices meet in one gather and the counts dict is rebuilt in shot order.
        Identical to `run_with_noise` on one process, key order included.  `_indices_fn` is for the CPU gloo test."""
        from qsb import distributed as D
        if self._noise_model is None:
            return self.run(circuit, shots, seed=seed, rng=rng)
        world, rank = D.world_info()
        n = circuit.num_qubits
        d = self._program(circuit)[0].prog.n_draws if _indices_fn is None else _indices_fn.n_draws
        lo, hi = D.shard_bounds(shots, world, rank)
        noise_rng = D.positioned_rng(self._noise_model._rng, lo * d)
        meas_rng = D.positioned_rng(rng if rng is not None else seed, lo)
        cnt = hi - lo
        if _indices_fn is None:                    # same pipelined device leg as run_with_noise, on this rank's slice
            idx = self._noisy_indices(circuit, cnt, noise_rng, meas_rng) if cnt else np.zeros(0, dtype=np.int64)
        else:
            uniforms = noise_rng.random(cnt * d).reshape(cnt, d) if d else None
            measure_u = meas_rng.random(cnt)
            idx = np.asarray(_indices_fn(circuit, uniforms, measure_u), dtype=np.int64) if cnt else np.zeros(0, dtype=np.int64)
        all_idx = D.gather_concat(idx)
        # leave both generators where the single loop would have left them
        if d:
            self._noise_model._rng.bit_generator.advance(shots * d)
        if rng is not None:
            rng.bit_generator.advance(shots)
        counts = D.merge_counts_in_shot_order(all_idx, n)
        final_state = StateVector.from_initial_states(circuit.initial_states)
        return SimulationResult(final_state=final_state, measurement_counts=counts, num_shots=shots, seed=seed)

    def ensemble_density_matrix_sharded(self, circuit: QuantumCircuit, n_trials: int = 50, seed: int | None = None) -> np.ndarray:
        """`ensemble_density_matrix` with the trials split over the ranks: every rank walks the child-seed chain
        (simulator.py:175-182), accumulates its trials on its GPU and ONE all-reduce sums the partial rho's."""
        import torch
        from qsb import distributed as D
        world, rank = D.world_info()
        rng = np.random.default_rng(seed)
        n = circuit.num_qubits
        dim = 2 ** n
        dp, _ = self._program(circuit)
        d = dp.prog.n_draws
        c = self._ctx()
        seeds = [int(rng.integers(0, 2 ** 63)) for _ in range(n_trials)]
        t_rho = torch.zeros(2 * dim * dim, dtype=torch.float64, device=torch.device("cuda", c.device))
        torch.cuda.current_stream(t_rho.device).synchronize()
        rho = c.wrap(t_rho.data_ptr(), 16 * dim * dim)
        lo_r, hi_r = D.shard_bounds(n_trials, world, rank)
        chunk = max(1, min(max(hi_r - lo_r, 1), _CHUNK_BYTES // (16 * dim)))
        for lo in range(lo_r, hi_r, chunk):
            cnt = min(chunk, hi_r - lo)
            uniforms = None
            if d:
                uniforms = np.empty((cnt, d), dtype=np.float64)
                for i in range(cnt):
                    uniforms[i] = np.random.default_rng(seeds[lo + i]).random(d)
            _, states = self._trajectory_batch(circuit, uniforms, cnt)
            c.rho_accumulate(n, states, 0, cnt, 1.0 / n_trials, rho)
        c.sync()
        D.allreduce_sum_(t_rho)
        if self._noise_model is not None and n_trials > 0:
            self._noise_model.set_seed(seeds[-1])
            if d:
                self._noise_model._rng.random(d)
        return t_rho.cpu().numpy().view(np.complex128).reshape(dim, dim)

    def ensemble_density_matrix(self, circuit: QuantumCircuit, n_trials: int = 50,
                                seed: int | None = None) -> np.ndarray:
        """rho = (1/N) sum_i |psi_i><psi_i| over N stochastic trajectories with per-trial child seeds."""
        rng = np.random.default_rng(seed)
        n = circuit.num_qubits
        dim = 2 ** n
        dp, _ = self._program(circuit)
        d = dp.prog.n_draws
        c = self._ctx()
        rho = c.alloc(16 * dim * dim).zero()
        seeds = [int(rng.integers(0, 2 ** 63)) for _ in range(n_trials)]     # drawn even without noise
        chunk = max(1, min(n_trials, _CHUNK_BYTES // (self._amp()[0] * dim)))
        for lo in range(0, n_trials, chunk):
            cnt = min(chunk, n_trials - lo)
            uniforms = None
            if d and self._rng_mode == "reference":
                uniforms = np.empty((cnt, d), dtype=np.float64)
                for i in range(cnt):
                    uniforms[i] = np.random.default_rng(seeds[lo + i]).random(d)
            _, states = self._trajectory_batch(circuit, uniforms, cnt)
            c.rho_accumulate(n, states, 0, cnt, 1.0 / n_trials, rho)
        if self._noise_model is not None and n_trials > 0 and self._rng_mode == "reference":
            self._noise_model.set_seed(seeds[-1])      # the reference leaves the model seeded with the last child seed
            if d:
                self._noise_model._rng.random(d)
        return rho.download(np.complex128, (dim, dim))
