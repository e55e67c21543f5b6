"""Property test (CPU): random circuits, noise models, cluster shapes and worker counts through the host compiler and
the executor's op loop (host emulator: the same `qsb_exec.cuh`), against the oracle.  Targets the index arithmetic
that no fixed case pins: multi-pair exchanges on tiny tiles (a single group per CTA), rank-bit flushes by halves,
pending matrices parked on rank bits, snapshot permutations after arbitrary axis scrambles."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import qsim_oracle as O
from emu_util import emu_run
from qsb.lowering import lower_circuit
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.gate_registry import GateRegistry

REG = GateRegistry.instance()
ONE = ["H", "X", "Y", "Z", "S", "T", "Rx", "Ry", "Rz", "Phase", "U3"]
TWO = ["CNOT", "CZ", "SWAP"]
THREE = ["Toffoli", "Fredkin"]
NOISE = ["bit_flip", "phase_flip", "depolarizing", "amplitude_damping"]


@st.composite
def cases(draw):
    n = draw(st.integers(3, 7))
    gmax = min(3, n - 3)
    gbits = draw(st.sampled_from([0, gmax, gmax, draw(st.integers(0, gmax))]))       # biased towards the largest cluster
    seed = draw(st.integers(0, 2 ** 31 - 1))
    n_gates = draw(st.integers(1, 30))
    noisy = draw(st.booleans())
    workers = draw(st.sampled_from([1, 2]))
    return n, gbits, seed, n_gates, noisy, workers


@settings(max_examples=250, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(cases())
def test_random_programs_match_the_oracle(case):
    _check(case, None)


@pytest.mark.gpu
@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(cases())
def test_random_programs_match_the_oracle_on_the_gpu(case):
    """The same generator through the C ABI and the CUDA kernel (clusters of 1..8 CTAs, tiles of 1..7 bits)."""
    from gpu_util import gpu_run
    _check(case, gpu_run)


@st.composite
def big_cases(draw):
    n = draw(st.sampled_from([14, 15, 16]))
    gbits = n - 13
    return n, gbits, draw(st.integers(0, 2 ** 31 - 1)), draw(st.integers(15, 45)), draw(st.booleans()), 1


@pytest.mark.gpu
@settings(max_examples=10, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(big_cases())
def test_random_production_shapes_on_the_gpu(case):
    """14-16 qubits: 2^13-amplitude tiles, 256 workers, clusters of 2 / 4 / 8 -- the shapes the benchmark runs."""
    from gpu_util import gpu_run
    _check(case, gpu_run)


@pytest.mark.gpu
@settings(max_examples=80, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(cases())
def test_random_programs_in_complex64_on_the_gpu(case):
    """complex64 mode (tolerance 1e-5; a Kraus branch may differ from the complex128 oracle only when a uniform lands
    within float rounding of a threshold -- such a trajectory is skipped, not compared)."""
    from gpu_util import gpu_run
    _check(case, lambda *a, **k: gpu_run(*a, precision="c64", **k), tol=1e-5, strict_branches=False)


def _check(case, gpu_run, tol=1e-12, strict_branches=True):
    n, gbits, seed, n_gates, noisy, workers = case
    rng = np.random.default_rng(seed)
    names = [x for x in ONE + TWO + THREE if x in O._FIXED or x in O.NUM_PARAMS or x in ("CNOT", "CZ", "SWAP", "Toffoli", "Fredkin")]
    gates = []
    for i in range(n_gates):
        name = names[int(rng.integers(0, len(names)))] if rng.random() < 0.5 else (TWO + THREE)[int(rng.integers(0, 5))]
        k = 1 if name in ONE else 2 if name in TWO else 3
        targets = rng.permutation(n)[:k].tolist()
        params = [float(x) for x in rng.uniform(-np.pi, np.pi, O.NUM_PARAMS.get(name, 0))]
        gates.append((name, targets, params, i // 2))
    noise = None
    if noisy:
        noise = {"global": [(NOISE[int(rng.integers(0, 4))], float(rng.uniform(0.05, 0.6))) for _ in range(int(rng.integers(1, 3)))],
                 "gate": {"CNOT": [("amplitude_damping", 0.5)]} if rng.random() < 0.5 else {}, "readout": None}
    initial = [int(b) for b in rng.integers(0, 2, n)]
    qc = QuantumCircuit(n, initial_states=list(initial))
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    channels = (lambda name: [(k, p, None) for k, p in O.channels_for(noise, name)]) if noise else None
    prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, channels, record_steps=True, local_bits=n - gbits)
    T = 2
    draws = rng.random((T, max(prog.n_draws, 1)))
    basis = sum(1 << (n - 1 - i) for i, b in enumerate(initial) if b)
    if gpu_run is None:
        out = emu_run(prog, count=T, T=workers, uniforms=draws if prog.n_draws else None, default_basis=basis, want_branches=True)
    else:
        out = gpu_run(prog, count=T, uniforms=draws if prog.n_draws else None, default_basis=basis, want_branches=True)
    for t in range(T):
        psi, steps, branches, _ = O.run_state(n, gates, initial, noise, draws[t] if prog.n_draws else None, record_steps=True)
        if prog.n_draws:
            same = out["branches"][t][:len(branches)].tolist() == branches
            assert same or not strict_branches, case
            if not same:
                continue
        assert np.max(np.abs(out["states"][t] - psi)) < tol, case
        if steps is not None and prog.n_snapshots:
            assert np.max(np.abs(out["snapshots"][t] - np.array(steps))) < tol, case
