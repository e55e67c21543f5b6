"""concurrence, EntanglementEventDetector, ConvergenceAnalysis, BenchmarkAnalysis (analysis.py:194-219, :255-621) on the
device backend against the real reference (tests/golden/make_golden_analysis2.py)."""
import json
import os

import numpy as np
import pytest

from test_gpu_engine_api import E, circuit_of, model_of   # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "golden_analysis2.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def run5(E, gold):
    g = gold["conc"]
    qc = circuit_of(E, g["n"], [(x[0], x[1], x[2], x[3]) for x in g["gates"]])
    return E.Simulator().run(qc, shots=0, record_steps=True)


def test_concurrence(E, gold, run5):
    for a, b, want in gold["conc"]["values"]:
        # lambda_2..4 are square roots of eigenvalues that are zero up to rounding: 1e-16 in rho becomes 1e-8 in C
        assert abs(E.StateAnalysis.concurrence(run5.final_state, a, b) - want) < 1e-7, (a, b)
    bell = circuit_of(E, 2, [("H", [0], [], 0), ("CNOT", [0, 1], [], 1)])
    assert abs(E.StateAnalysis.concurrence(E.Simulator().run(bell, shots=0).final_state, 0, 1) - gold["conc_bell"]) < 1e-12


def _rows(evs):
    return [[e.step, list(e.qubit_pair), e.event_type.value, e.magnitude, e.entropy_before, e.entropy_after] for e in evs]


def _same_events(got, want):
    assert len(got) == len(want), (got, want)
    for g, w in zip(got, want):
        assert g[:3] == w[:3], (g, w)
        assert np.max(np.abs(np.array(g[3:]) - np.array(w[3:]))) < 1e-10, (g, w)


def test_entanglement_event_detector(E, gold, run5):
    for rec in gold["detector"]:
        det = E.EntanglementEventDetector(**rec["kw"])
        for i, st in enumerate(run5.step_states):
            _same_events(_rows(det.process_step(st, i)), rec["per_step"][i])
        hist = det.get_pair_history(3, 1)
        assert [h[0] for h in hist] == [h[0] for h in rec["history_3_1"]]
        assert np.max(np.abs(np.array([h[1] for h in hist]) - np.array([h[1] for h in rec["history_3_1"]]))) < 1e-10
        assert len(det.get_all_pair_histories()) == 10
        det.reset()
        _same_events(_rows(det.process_step(run5.step_states[-1], 99)), rec["after_reset"])
        assert len(det.get_timeline()) == rec["n_timeline"]
    assert E.EntanglementEventType.CREATION.value == "creation"
    assert E.EntanglementEvent(1, (0, 1), E.EntanglementEventType.INCREASE, 0.1, 0.0, 0.1).qubit_pair == (0, 1)


def test_convergence_analysis(E, gold, run5):
    sv = run5.final_state
    rows = E.ConvergenceAnalysis.shot_convergence(sv, [10, 100, 1000, 5000], seed=5)
    for got, want in zip(rows, gold["shot_conv"]):
        assert got["shots"] == want["shots"]
        assert abs(got["tvd"] - want["tvd"]) < 1e-12 and abs(got["kl_divergence"] - want["kl_divergence"]) < 1e-10
    counts = E.MeasurementEngine.sample(sv, 300, rng=np.random.default_rng(8))
    assert counts == gold["counts300"]
    p = sv.probabilities
    assert abs(E.ConvergenceAnalysis.tvd(p, counts, 300) - gold["tvd"]) < 1e-12
    assert abs(E.ConvergenceAnalysis.kl_divergence(p, counts, 300) - gold["kl"]) < 1e-10
    assert abs(E.ConvergenceAnalysis.kl_divergence(p, counts, 300, epsilon=1e-3) - gold["kl_eps"]) < 1e-10


def test_benchmark_analysis(E, gold):
    assert E.BenchmarkAnalysis.quantum_volume(max_qubits=4, num_trials=5, seed=3) == gold["qv_ideal"]
    nm = E.NoiseModel()
    nm.add_global_noise(E.DepolarizingNoise(0.2))
    nm.set_seed(17)
    assert E.BenchmarkAnalysis.quantum_volume(max_qubits=3, num_trials=6, noise_model=nm, seed=4) == gold["qv_noisy"]
    rows = E.BenchmarkAnalysis.gate_timing(range(2, 5), E.H_MATRIX, lambda nq: [nq - 1], repetitions=3)
    assert [r["num_qubits"] for r in rows] == [2, 3, 4] and all(r["mean_time_ms"] > 0 for r in rows)
