"""The drop-in engine API (quantum_sim.engine of this repo) against golden outputs of the reference.
These read like the reference's own calls; every array op behind them runs on the GPU.  GPU only."""

import numpy as np
import pytest

from conftest import as_gates, as_noise
from oracle import qsim_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def E():
    import types
    from quantum_sim.engine import circuit, state_vector, simulator, noise, measurement, analysis, gates
    ns = types.SimpleNamespace()
    for mod in (circuit, state_vector, simulator, noise, measurement, analysis, gates):
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(ns, k, v)
    ns.NOISE = {"bit_flip": noise.BitFlipNoise, "phase_flip": noise.PhaseFlipNoise,
                "depolarizing": noise.DepolarizingNoise, "amplitude_damping": noise.AmplitudeDampingNoise}
    return ns


def circuit_of(E, n, gates, initial=None):
    qc = E.QuantumCircuit(n, initial_states=list(initial) if initial else [])
    for g in gates:
        qc.add_gate(E.GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    return qc


def model_of(E, spec, seed=None):
    nm = E.NoiseModel()
    for kind, p in spec.get("global", []):
        nm.add_global_noise(E.NOISE[kind](p))
    for name, chans in (spec.get("gate") or {}).items():
        for kind, p in chans:
            nm.add_gate_noise(name, E.NOISE[kind](p))
    if spec.get("readout"):
        nm.set_readout_error(E.ReadoutError(*spec["readout"]))
    if seed is not None:
        nm.set_seed(seed)
    return nm


def sv_of(E, psi, n):
    sv = E.StateVector(n)
    sv.data = psi
    return sv


def test_config1_ghz3(E, golden):
    j, a = golden
    qc = circuit_of(E, 3, as_gates(j["ghz3"]["gates"]))
    res = E.Simulator().run(qc, shots=0)
    assert res.measurement_counts == {}
    assert np.max(np.abs(res.final_state.data - a["ghz3_state"])) < TOL
    for b in "ZXY":
        r = E.Simulator().run(qc, shots=1024, seed=42, measurement_basis=E.MeasurementBasis[b])
        assert r.measurement_counts == j["ghz3"][f"counts_{b}"], b
    nm = model_of(E, {"readout": (0.02, 0.05)})
    assert E.Simulator(nm).run(qc, shots=1024, seed=42).measurement_counts == j["ghz3"]["counts_readout"]
    d = E.ReadoutError(0.02, 0.05).apply_to_distribution(res.final_state.probabilities, 3)
    assert np.max(np.abs(d - a["ghz3_readout_dist"])) < 1e-15
    mi = [E.StateAnalysis.mutual_information(res.final_state, i, k) for i in range(3) for k in range(i + 1, 3)]
    assert np.allclose(mi, j["ghz3"]["mi"], atol=1e-12)
    assert abs(E.StateAnalysis.entanglement_entropy(res.final_state, [0]) - j["ghz3"]["entropy_q0"]) < 1e-12
    nm = model_of(E, {"global": [("depolarizing", 0.1)]}, seed=7)
    assert E.Simulator(nm).run_with_noise(qc, shots=200, seed=42).measurement_counts == j["ghz3"]["run_with_noise"]


def test_apply_gate_scramble_every_target_list(E, golden):
    j, a = golden
    n = j["sigma"]["n"]
    for i, t in enumerate(j["sigma"]["targets"]):
        k = len(t)
        sv = sv_of(E, a["sigma_in"][i], n)
        sv.apply_gate(a["sigma_mat"][i][:4 ** k].reshape(2 ** k, 2 ** k), t)
        assert np.max(np.abs(sv.data - a["sigma_out"][i])) < TOL, t
    with pytest.raises(ValueError):
        E.StateVector(3).apply_gate(np.eye(2), [3])
    with pytest.raises(ValueError):
        E.StateVector(17)
    with pytest.raises(ValueError):
        E.StateVector(2).data = np.zeros(3)


def test_random_circuits_run_and_steps(E, golden):
    j, a = golden
    for rec in j["random_circuits"]:
        qc = circuit_of(E, rec["n"], as_gates(rec["gates"]), rec["initial"])
        res = E.Simulator().run(qc, shots=0, record_steps=True)
        assert np.max(np.abs(res.final_state.data - a[rec["tag"]])) < TOL
        steps = np.array([s.data for s in res.step_states])
        assert np.max(np.abs(steps - a[rec["tag"] + "_steps"])) < TOL
    rec = j["random_circuits"][5]
    qc = circuit_of(E, rec["n"], as_gates(rec["gates"]), rec["initial"])
    seq = list(E.Simulator().run_step_by_step(qc))
    assert seq[0][1] == -1 and [i for _, i in seq[1:]] == list(range(len(seq) - 1))
    assert np.max(np.abs(np.array([s.data for s, _ in seq[1:]]) - a[rec["tag"] + "_steps"])) < TOL


def test_run_step_by_step_is_lazy_like_the_reference(E, golden):
    """simulator.py:93-108 is a generator over columns: the noise model's draws are taken column by column, and an
    unknown gate in a late column raises only once that column is reached."""
    j, a = golden
    rec = next(r for r in j["noisy"] if len({g[3] for g in r["gates"]}) >= 3)
    n, gates, noise = rec["n"], as_gates(rec["gates"]), as_noise(rec["noise"])
    qc = circuit_of(E, n, gates)
    nm = model_of(E, noise, seed=rec["noise_seed"])
    it = E.Simulator(nm).run_step_by_step(qc)
    s0, i0 = next(it)
    assert i0 == -1 and abs(s0.data[0] - 1.0) < TOL
    ref_rng = np.random.default_rng(rec["noise_seed"])
    assert nm._rng.bit_generator.state == ref_rng.bit_generator.state           # nothing drawn yet
    s1, i1 = next(it)
    first_col = min(g[3] for g in gates)
    ref_rng.random(O.draw_count(n, [g for g in gates if g[3] == first_col], noise))
    assert i1 == 0 and nm._rng.bit_generator.state == ref_rng.bit_generator.state   # ... only the first column's draws
    last = s1
    for st, _ in it:
        last = st
    assert np.max(np.abs(last.data - a[rec["tag"]])) < TOL
    # a gate name the registry does not know, in the last column
    qc2 = circuit_of(E, n, gates)
    qc2.add_gate(E.GateInstance("NoSuchGate", [0], [], max(g[3] for g in gates) + 1))
    it2 = E.Simulator().run_step_by_step(qc2)
    seen = 0
    with pytest.raises(KeyError):
        for _ in it2:
            seen += 1
    assert seen == 1 + len({g[3] for g in gates})                                # the initial state and every good column


def test_per_gate_api_matches_batched_run(E, golden):
    """simulator._apply_gate_instance + NoiseModel.apply called gate by gate (the reference's own loop)
    gives the same trajectory as the single-launch run."""
    j, a = golden
    for rec in j["noisy"][:12]:
        n = rec["n"]
        qc = circuit_of(E, n, as_gates(rec["gates"]))
        nm = model_of(E, as_noise(rec["noise"]), seed=rec["noise_seed"])
        sim = E.Simulator(nm)
        state = E.StateVector.from_initial_states(qc.initial_states)
        for col in qc.get_ordered_gates():
            for g in col:
                sim._apply_gate_instance(state, g)
                nm.apply(state, g)
        assert np.max(np.abs(state.data - a[rec["tag"]])) < TOL, rec["tag"]


def test_noisy_runs(E, golden):
    j, a = golden
    for rec in j["noisy"]:
        qc = circuit_of(E, rec["n"], as_gates(rec["gates"]))
        nm = model_of(E, as_noise(rec["noise"]), seed=rec["noise_seed"])
        res = E.Simulator(nm).run(qc, shots=0, record_steps=True)
        assert np.max(np.abs(res.final_state.data - a[rec["tag"]])) < TOL, rec["tag"]
        steps = np.array([s.data for s in res.step_states])
        assert np.max(np.abs(steps - a[rec["tag"] + "_steps"])) < TOL


def test_ensemble_density_matrix(E, golden):
    j, a = golden
    for key, n in (("ens4", 4), ("ens5", 5)):
        rec = j[key]
        nm = model_of(E, as_noise(rec["noise"]))
        rho = E.Simulator(nm).ensemble_density_matrix(circuit_of(E, n, as_gates(rec["gates"])),
                                                      rec["n_trials"], seed=rec["seed"])
        assert np.max(np.abs(rho - a[key + "_rho"])) < TOL
    assert abs(E.StateAnalysis.purity_dm(a["ens4_rho"]) - j["ens4"]["purity"]) < 1e-13
    rho = E.Simulator().ensemble_density_matrix(circuit_of(E, 5, as_gates(j["ens5"]["gates"])), 3, seed=9)
    assert np.max(np.abs(rho - a["ens5_clean_rho"])) < TOL


def test_run_with_noise_counts(E, golden):
    j, _ = golden
    for rec in j["run_with_noise"]:
        nm = model_of(E, as_noise(rec["noise"]), seed=rec["noise_seed"])
        res = E.Simulator(nm).run_with_noise(circuit_of(E, rec["n"], as_gates(rec["gates"])),
                                             shots=rec["shots"], seed=rec["seed"])
        assert res.measurement_counts == rec["counts"]
        assert list(res.measurement_counts) == list(rec["counts"])        # same insertion order


def test_measurement_engine(E, golden):
    j, a = golden
    for rec in j["measurement"]:
        n, psi = rec["n"], a[rec["tag"]]
        for key, want in rec["counts"].items():
            basis, mode = key.split("_")
            ro = None if mode == "None" else E.ReadoutError(0.1, 0.07)
            got = E.MeasurementEngine.sample_with_basis(
                sv_of(E, psi, n), 500, basis=E.MeasurementBasis[basis], readout_error=ro,
                readout_mode="shot" if mode == "None" else mode, rng=np.random.default_rng(77))
            assert got == want and list(got) == list(want), key
        assert sv_of(E, psi, n).measure_all(np.random.default_rng(3)) == rec["measure_all"]
        sv = sv_of(E, psi, n)
        r = np.random.default_rng(4)
        assert [sv.measure_qubit(q, r) for q in range(n)] == rec["measure_qubit"]
        assert np.max(np.abs(sv.data - a[rec["tag"] + "_after_mq"])) < TOL
        bl = [list(sv_of(E, psi, n).get_bloch_coordinates(q)) for q in range(n)]
        assert np.allclose(bl, rec["bloch"], atol=1e-13)
        rdm = np.array([sv_of(E, psi, n).get_reduced_density_matrix(q) for q in range(n)])
        assert np.max(np.abs(rdm - a[rec["tag"] + "_rdm1"])) < TOL


def test_analysis(E, golden):
    from quantum_sim.engine.analysis import all_pairs_mutual_information
    j, a = golden
    for rec in j["analysis"]:
        n, psi = rec["n"], a[rec["tag"]]
        sv = sv_of(E, psi, n)
        assert np.allclose(all_pairs_mutual_information(sv), rec["mi"], atol=1e-11)
        rdm2 = np.array([E.StateAnalysis.partial_trace(sv, [i, k]) for i in range(n) for k in range(i + 1, n)])
        assert np.max(np.abs(rdm2 - a[rec["tag"] + "_rdm2"])) < TOL
        ent = [E.StateAnalysis.entanglement_entropy(sv, [q]) for q in range(n)]
        assert np.allclose(ent, rec["entropy_1q"], atol=1e-12)
        for ev in rec["expect"]:
            obs = np.array([[1]], dtype=complex)
            for ch in ev["label"]:
                obs = np.kron(obs, {"X": E.X_MATRIX, "Y": E.Y_MATRIX, "Z": E.Z_MATRIX}[ch])
            v = E.StateAnalysis.expectation_value(sv, obs, ev["qubits"])
            assert abs(v - complex(ev["re"], ev["im"])) < TOL, ev
        assert abs(E.StateAnalysis.state_fidelity(psi, a[rec["tag"] + "_phi"]) - rec["fidelity"]) < TOL
        assert np.max(np.abs(sv.data - psi)) == 0          # analysis leaves the state untouched


def test_reference_validation_suite_shapes(E):
    """The assertions of the reference's test_validation.py tests 1, 2, 7 (Bell, normalisation, CPTP)."""
    qc = circuit_of(E, 2, [("H", [0], [], 0), ("CNOT", [0, 1], [], 1)])
    st = E.Simulator().run(qc, shots=0).final_state
    amp = np.abs(st.data)
    assert abs(amp[0] - 1 / np.sqrt(2)) < 1e-8 and abs(amp[3] - 1 / np.sqrt(2)) < 1e-8 and amp[1] < 1e-8 and amp[2] < 1e-8
    assert abs(E.StateAnalysis.mutual_information(st, 0, 1) - 2.0) < 0.01
    assert abs(E.StateAnalysis.entanglement_entropy(st, [0]) - 1.0) < 0.01
    qc3 = circuit_of(E, 3, [("H", [0], [], 0), ("CNOT", [0, 1], [], 1), ("Rz", [2], [1.234], 0), ("Ry", [1], [0.567], 2)])
    nm = E.NoiseModel()
    nm.add_global_noise(E.DepolarizingNoise(0.05))
    for sim in (E.Simulator(), E.Simulator(nm)):
        d = sim.run(qc3, shots=0).final_state.data
        assert abs(np.sum(np.abs(d) ** 2) - 1.0) < 1e-8
    for gamma in (0.0, 0.3, 1.0):
        nm = E.NoiseModel()
        nm.add_global_noise(E.AmplitudeDampingNoise(gamma))
        d = E.Simulator(nm).run(circuit_of(E, 1, [("X", [0], [], 0)]), shots=0).final_state.data
        assert abs(np.sum(np.abs(d) ** 2) - 1.0) < 1e-8
        if gamma == 1.0:
            assert abs(abs(d[0]) - 1.0) < 1e-8


def _noisy_circuit(n=10, depth=6):
    from qsb.workloads import layered_circuit
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
    qc = QuantumCircuit(n)
    for name, targets, params, col in layered_circuit(n, depth, 31):
        qc.add_gate(GateInstance(name, list(targets), list(params), col))
    nm = NoiseModel()
    nm.add_global_noise(DepolarizingNoise(0.02))
    nm.add_global_noise(AmplitudeDampingNoise(0.05))
    return qc, nm


def test_simulator_complex64_precision_through_the_engine_api():
    """Simulator(..., precision="c64") (SURVEY section 5's precision switch): complex64 states on the device, results
    as complex128 arrays, within BASELINE's 1e-5 of the complex128 run on the same draws."""
    from quantum_sim.engine.simulator import Simulator
    qc, nm = _noisy_circuit()
    nm.set_seed(5)
    a = Simulator(nm).run(qc, shots=0, record_steps=True)
    nm.set_seed(5)
    b = Simulator(nm, precision="c64").run(qc, shots=0, record_steps=True)
    assert b.final_state.data.dtype == np.complex128
    assert np.max(np.abs(a.final_state.data - b.final_state.data)) < 1e-5
    assert np.max(np.abs(a.final_state.data - b.final_state.data)) > 0          # it really ran in another precision
    for x, y in zip(a.step_states, b.step_states):
        assert np.max(np.abs(x.data - y.data)) < 1e-5
    nm.set_seed(6)
    ca = Simulator(nm).run_with_noise(qc, shots=300, seed=9).measurement_counts
    nm.set_seed(6)
    cb = Simulator(nm, precision="c64").run_with_noise(qc, shots=300, seed=9).measurement_counts
    assert sum(cb.values()) == 300
    same = sum(min(ca.get(k, 0), cb.get(k, 0)) for k in set(ca) | set(cb))
    assert same >= 297                                   # a shot may flip only when a uniform sits within float rounding of a threshold
    ra = Simulator(nm).ensemble_density_matrix(qc, 40, seed=3)
    rb = Simulator(nm, precision="c64").ensemble_density_matrix(qc, 40, seed=3)
    assert np.max(np.abs(ra - rb)) < 1e-5 and abs(np.trace(rb).real - 1.0) < 1e-5
    with pytest.raises(ValueError):
        Simulator(nm, precision="fp16")


def test_simulator_philox_mode_through_the_engine_api():
    """Simulator(..., rng_mode="philox"): the in-kernel counter-based stream.  Trajectory t, draw d uses
    philox_uniform(seed, t, d) -- feeding exactly those numbers through the reference-draw path gives bit-identical
    states, trajectories are numbered consecutively over calls, and a different seed gives other trajectories."""
    from qsb.stream import philox_uniform
    from quantum_sim.engine.simulator import Simulator
    qc, nm = _noisy_circuit(8, 5)
    seed, count = 1234567, 12
    ph = Simulator(nm, rng_mode="philox", philox_seed=seed)
    ref = Simulator(nm)
    d = ref._program(qc)[0].prog.n_draws
    n = qc.num_qubits
    U = np.array([[philox_uniform(seed, t, k) for k in range(d)] for t in range(2 * count)])
    for call in range(2):                                 # the second call continues at trajectory `count`
        _, sp = ph._trajectory_batch(qc, None, count)
        _, sr = ref._trajectory_batch(qc, U[call * count:(call + 1) * count], count)
        assert np.array_equal(sp.download(np.complex128, (count, 2 ** n)), sr.download(np.complex128, (count, 2 ** n)))
    other = Simulator(nm, rng_mode="philox", philox_seed=seed + 1)
    _, so = other._trajectory_batch(qc, None, count)
    _, s0 = Simulator(nm, rng_mode="philox", philox_seed=seed)._trajectory_batch(qc, None, count)
    assert not np.array_equal(so.download(np.complex128, (count, 2 ** n)), s0.download(np.complex128, (count, 2 ** n)))
    # the public entry points run in this mode without touching the noise model's generator
    state = nm._rng.bit_generator.state
    res = Simulator(nm, rng_mode="philox", philox_seed=7).run_with_noise(qc, shots=500, seed=1)
    assert sum(res.measurement_counts.values()) == 500
    rho = Simulator(nm, rng_mode="philox", philox_seed=7).ensemble_density_matrix(qc, 64, seed=1)
    assert abs(np.trace(rho).real - 1.0) < 1e-12 and np.max(np.abs(rho - rho.conj().T)) < 1e-15
    assert nm._rng.bit_generator.state == state
