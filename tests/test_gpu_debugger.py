"""CircuitDebugger against values the real reference produced (tests/golden/make_golden.py --debugger)."""

import json
import os

import numpy as np
import pytest

from conftest import as_gates, as_noise

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dbg_golden():
    return (json.load(open(os.path.join(GOLDEN_DIR, "golden_debugger.json"))),
            np.load(os.path.join(GOLDEN_DIR, "golden_debugger.npz")))


def _setup(j, seed=None):
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine import noise as N
    qc = QuantumCircuit(j["n"])
    for g in as_gates(j["gates"]):
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    cls = {"bit_flip": N.BitFlipNoise, "phase_flip": N.PhaseFlipNoise, "depolarizing": N.DepolarizingNoise,
           "amplitude_damping": N.AmplitudeDampingNoise}
    nm = N.NoiseModel()
    for kind, p in j["noise"]["global"]:
        nm.add_global_noise(cls[kind](p))
    for name, chans in j["noise"]["gate"].items():
        for kind, p in chans:
            nm.add_gate_noise(name, cls[kind](p))
    if seed is not None:
        nm.set_seed(seed)
    return qc, nm


def test_full_debug_snapshots_and_stepping(dbg_golden):
    from quantum_sim.engine.debugger import CircuitDebugger
    j, a = dbg_golden
    qc, nm = _setup(j, seed=j["full_debug"]["noise_seed"])
    dbg = CircuitDebugger()
    snaps = dbg.run_full_debug(qc, nm, seed=3)
    fd = j["full_debug"]
    assert [s.column_index for s in snaps] == fd["columns"]
    assert [s.gate_labels for s in snaps] == fd["labels"]
    assert np.max(np.abs(np.array([s.fidelity for s in snaps]) - np.array(fd["fidelity"]))) < 1e-12
    assert np.max(np.abs(np.array([s.cumulative_fidelity for s in snaps]) - np.array(fd["cumulative_fidelity"]))) < 1e-12
    assert np.max(np.abs(np.array([s.entropy for s in snaps]) - np.array(fd["entropy"]))) < 1e-9
    assert np.max(np.abs(np.stack([s.state.data for s in snaps]) - a["dbg_states"])) < 1e-12
    assert np.max(np.abs(np.stack([s.ideal_state.data for s in snaps]) - a["dbg_ideal"])) < 1e-12
    # stepping / breakpoints
    assert dbg.current_snapshot.column_index == -1 and dbg.step_backward() is None
    assert dbg.step_forward().column_index == 0
    dbg.add_breakpoint(2)
    assert dbg.run_to_breakpoint().column_index == 2
    assert dbg.toggle_breakpoint(2) is False and dbg.run_to_breakpoint().column_index == fd["columns"][-1]
    assert dbg.goto_step(1).column_index == 0 and dbg.num_steps == len(fd["columns"])
    d = CircuitDebugger.compute_state_diff(snaps[1], snaps[-1])
    sd = j["state_diff"]
    assert abs(d["fidelity"] - sd["fidelity"]) < 1e-12 and abs(d["tvd"] - sd["tvd"]) < 1e-12
    assert np.max(np.abs(d["prob_diffs"] - a["dbg_prob_diffs"])) < 1e-12
    assert [x[0] for x in d["amplitude_diffs"]] == [x[0] for x in sd["amplitude_diffs"]]
    # noiseless run
    snaps0 = CircuitDebugger().run_full_debug(qc, None, seed=3)
    assert all(s.fidelity == 1.0 and s.ideal_state is None for s in snaps0)
    assert np.max(np.abs(np.stack([s.state.data for s in snaps0]) - a["dbg_states_noiseless"])) < 1e-12


def test_noise_impact_and_attribution(dbg_golden):
    from quantum_sim.engine.debugger import CircuitDebugger
    j, _ = dbg_golden
    qc, nm = _setup(j)
    imp = CircuitDebugger().compute_noise_impact(qc, nm, n_trials=7, seed=5)
    assert len(imp) == len(j["noise_impact"])
    for got, want in zip(imp, j["noise_impact"]):
        assert got.column_index == want["column_index"] and got.gate_labels == want["gate_labels"]
        for f in ("fidelity_before", "fidelity_after", "fidelity_drop", "mean_delta_fidelity", "std_delta_fidelity"):
            assert abs(getattr(got, f) - want[f]) < 1e-12, f
        for f in ("entropy_before", "entropy_after", "entropy_change"):
            assert abs(getattr(got, f) - want[f]) < 1e-9, f
        assert np.max(np.abs(np.array(got.per_qubit_fidelity) - np.array(want["per_qubit_fidelity"]))) < 1e-7
    qc, nm = _setup(j)
    att = CircuitDebugger().compute_noise_attribution(qc, nm, n_trials=7, seed=6)
    want = j["noise_attribution"]
    assert np.max(np.abs(np.array(att.delta_fidelity) - np.array(want["delta_fidelity"]))) < 1e-12
    assert np.max(np.abs(np.array(att.delta_fidelity_std) - np.array(want["delta_fidelity_std"]))) < 1e-12
    assert abs(att.total_fidelity_loss - want["total_fidelity_loss"]) < 1e-12
    assert np.max(np.abs(np.array(att.column_attribution_pct) - np.array(want["column_attribution_pct"]))) < 1e-9
    assert np.max(np.abs(np.array(att.per_qubit_attribution) - np.array(want["per_qubit_attribution"]))) < 1e-7
    assert att.gate_labels == want["gate_labels"] and att.is_recovery == want["is_recovery"]
    assert att.no_measurable_loss == want["no_measurable_loss"]
