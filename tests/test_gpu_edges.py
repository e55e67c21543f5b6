"""Edge cases of the engine API against the real reference (tests/golden/golden_edges.*, written by
`tests/golden/make_golden.py --edges`): 1 and 16 qubits, empty / measure-only circuits, shots = 0, channel
probabilities 0 and 1, gate-specific noise only, readout extremes, collapse sequences, error classes and texts."""
import json
import os

import numpy as np
import pytest

from test_gpu_engine_api import E, circuit_of, model_of, sv_of   # noqa: F401  (fixture + helpers)

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


@pytest.fixture(scope="module")
def edges():
    with open(os.path.join(HERE, "golden", "golden_edges.json")) as f:
        j = json.load(f)
    a = np.load(os.path.join(HERE, "golden", "golden_edges.npz"))
    return j, a


RUN_CASES = ["one_qubit", "one_qubit_y", "empty", "empty_shots0", "measure_only", "measure_shots0", "ghz16", "max16_layer",
             "sparse_columns", "p0_noise", "p1_bitflip", "p1_phaseflip", "gamma1_damping", "p1_depol", "gate_noise_only",
             "readout_extremes"]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", RUN_CASES)
def test_run_edge_cases(E, edges, tag):
    j, a = edges
    rec = j[tag]
    gates = [(g[0], g[1], g[2], g[3]) for g in rec["gates"]]
    qc = circuit_of(E, rec["n"], gates, rec["initial"])
    nm = model_of(E, rec["noise"], rec["noise_seed"]) if rec["noise"] else None
    res = E.Simulator(nm).run(qc, shots=rec["shots"], seed=rec["seed"], record_steps=True,
                              measurement_basis=E.MeasurementBasis[rec["basis"]])
    assert np.max(np.abs(res.final_state.data - a[tag])) < TOL
    assert res.measurement_counts == rec["counts"]
    assert list(res.measurement_counts) == list(rec["counts"])                  # same insertion order
    if rec["n_steps"] is None:
        assert res.step_states is None
    else:
        assert len(res.step_states) == rec["n_steps"]
        if rec["n_steps"]:
            got = np.array([s.data for s in res.step_states])
            assert np.max(np.abs(got - a[tag + "_steps"])) < TOL


@pytest.mark.gpu
def test_run_with_noise_and_ensemble_small_counts(E, edges):
    j, a = edges
    from qsb.workloads import ghz
    qc = circuit_of(E, 3, ghz(3))
    for shots in (0, 1, 5):
        nm = model_of(E, {"global": [("depolarizing", 0.2)]}, 11)
        r = E.Simulator(nm).run_with_noise(qc, shots=shots, seed=12)
        assert r.measurement_counts == j[f"rwn_{shots}"]["counts"] and r.num_shots == j[f"rwn_{shots}"]["num_shots"]
    for tag, spec in (("ens_nonoise", None), ("ens_noise", {"global": [("depolarizing", 0.3)]})):
        for trials in (1, 3):
            nm = model_of(E, spec, 1) if spec else None
            rho = E.Simulator(nm).ensemble_density_matrix(qc, n_trials=trials, seed=8)
            assert np.max(np.abs(rho - a[f"{tag}_{trials}"])) < TOL


@pytest.mark.gpu
def test_collapse_sequences_and_resets(E, edges):
    j, a = edges
    rng = np.random.default_rng(77)

    def rand_state(seed, n):
        r = np.random.default_rng(seed)
        v = r.normal(size=2 ** n) + 1j * r.normal(size=2 ** n)
        return v / np.linalg.norm(v)

    sv = sv_of(E, rand_state(1, 4), 4)
    seq = []
    for q in (2, 0, 3, 1, 2):
        seq.append(int(sv.measure_qubit(q, rng)))
        assert np.max(np.abs(sv.data - a[f"collapse_{len(seq)}"])) < TOL
    assert seq == j["collapse_outcomes"]
    sv = sv_of(E, rand_state(2, 3), 3)
    assert sv.measure_all(np.random.default_rng(3)) == j["measure_all"]
    assert np.max(np.abs(sv.data - a["measure_all_state"])) < TOL
    sv = E.StateVector.from_initial_states([1, 1, 0, 1])
    assert np.array_equal(sv.data, a["from_initial"])
    sv.reset()
    assert np.array_equal(sv.data, a["reset_default"])
    sv.reset([0, 1, 1, 0])
    assert np.array_equal(sv.data, a["reset_states"])
    bl = [list(map(float, E.StateVector.from_initial_states([0, 1]).get_bloch_coordinates(q))) for q in (0, 1)]
    assert np.max(np.abs(np.array(bl) - np.array(j["bloch"]))) < TOL


@pytest.mark.gpu
def test_error_classes_and_texts(E, edges):
    j, _ = edges
    calls = {
        "sv_0": lambda: E.StateVector(0),
        "sv_17": lambda: E.StateVector(17),
        "data_shape": lambda: setattr(E.StateVector(2), "data", np.zeros(3, dtype=complex)),
        "qubit_range": lambda: E.StateVector(2).apply_gate(np.eye(2, dtype=complex), [2]),
        "qubit_negative": lambda: E.StateVector(2).apply_gate(np.eye(2, dtype=complex), [-1]),
        "unknown_gate": lambda: E.Simulator().run(circuit_of(E, 2, [("Nope", [0], [], 0)]), shots=0),
        "noise_p": lambda: E.DepolarizingNoise(1.5),
        "noise_p_neg": lambda: E.AmplitudeDampingNoise(-0.1),
        "readout_p": lambda: E.ReadoutError(2.0, 0.0),
    }
    for tag, fn in calls.items():
        want = j["errors"][tag]
        try:
            fn()
            got = None
        except Exception as e:                               # noqa: BLE001
            got = [type(e).__name__, str(e)]
        assert got == want, (tag, got, want)


@pytest.mark.gpu
def test_dense_operators_on_more_than_three_qubits(E):
    """apply_gate with dense k = 4..8 qubit matrices that are not Kronecker products (the reference accepts any k):
    `qsb_apply_dense`, scramble included, against the real reference (tests/golden/make_golden_densek.py)."""
    sys_path = os.path.join(HERE, "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_densek", os.path.join(sys_path, "make_golden_densek.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(sys_path, "golden_densek.npz"))
    for ci, (n, targets, m, psi) in enumerate(mod.densek_inputs()):
        sv = sv_of(E, psi, n)
        sv.apply_gate(m, targets)
        scale = np.max(np.abs(gold[f"c{ci}_out"]))
        assert np.max(np.abs(sv.data - gold[f"c{ci}_out"])) < 1e-12 * scale, (n, targets)
        sv.apply_gate(m.conj().T, targets[::-1])
        scale = np.max(np.abs(gold[f"c{ci}_out2"]))
        assert np.max(np.abs(sv.data - gold[f"c{ci}_out2"])) < 1e-12 * scale, (n, targets)
    # textbook layout: plain tensor action, no scramble
    n, targets, m, psi = mod.densek_inputs()[1]
    E.StateVector.layout = "textbook"
    try:
        sv = sv_of(E, psi, n)
        sv.apply_gate(m, targets)
        t = psi.reshape([2] * n)
        k = len(targets)
        want = np.tensordot(m.reshape([2] * (2 * k)), t, axes=(list(range(k, 2 * k)), targets))
        want = np.moveaxis(want, list(range(k)), targets).reshape(-1)
        assert np.max(np.abs(sv.data - want)) < 1e-12 * np.max(np.abs(want))
    finally:
        E.StateVector.layout = "reference"
    with pytest.raises(NotImplementedError):
        sv_of(E, np.ones(2 ** 10) / 32.0, 10).apply_gate(np.random.default_rng(0).normal(size=(512, 512)), list(range(9)))
