"""QEC cycles behind the reference's qec.py API (BASELINE config 4) against goldens produced by the real
reference (tests/golden/make_golden.py case_qec): single cycles through the per-state methods AND through the
batched path, full threshold sweeps, and scripts/qec_threshold.py's threshold rule."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("logical_rate", "success_rate", "avg_fidelity", "logical_z_fidelity", "decoder_success_rate",
          "projection_logical_rate")


def _codes():
    from quantum_sim.engine.qec import SteaneCode, BitFlipCode, PhaseFlipCode
    return {"steane": SteaneCode, "bit_flip": BitFlipCode, "phase_flip": PhaseFlipCode}


def test_single_cycles_match_reference(golden):
    from quantum_sim.engine.qec import QECSimulator
    j, a = golden
    codes = _codes()
    assert np.max(np.abs(codes["steane"]().encode(0).data - a["steane_enc0"])) == 0
    assert np.max(np.abs(codes["steane"]().encode(1).data - a["steane_enc1"])) == 0
    groups = {}
    for rec in j["qec_cycles"]:
        groups.setdefault((rec["code"], rec["noise_type"], rec["p"]), []).append(rec)
    for (code, ntype, p), recs in groups.items():
        sim = QECSimulator(codes[code]())
        # per-state path (run_cycle) on a few, batched path (run_cycles) on all
        for rec in recs[:3]:
            r = sim.run_cycle(rec["logical"], ntype, p, seed=rec["seed"])
            assert r.syndrome == rec["syndrome"]
            assert [list(c) for c in r.correction_applied] == rec["corrections"]
            assert abs(r.fidelity_before - rec["fidelity_before"]) < 1e-12
            assert abs(r.fidelity_after - rec["fidelity_after"]) < 1e-12
            assert abs(r.logical_z_expectation - rec["z_exp"]) < 1e-12
            assert bool(r.logical_error_detected) == rec["logical_error"]
        b = sim.run_cycles([r["logical"] for r in recs], ntype, p, [r["seed"] for r in recs])
        for t, rec in enumerate(recs):
            assert b["syndrome"][t].tolist() == rec["syndrome"], (code, ntype, rec["seed"])
            assert [list(c) for c in b["corrections"][t]] == rec["corrections"]
            assert abs(b["fidelity_before"][t] - rec["fidelity_before"]) < 1e-12
            assert abs(b["fidelity_after"][t] - rec["fidelity_after"]) < 1e-12
            assert abs(b["z_exp"][t] - rec["z_exp"]) < 1e-12
            assert bool(b["logical_error"][t]) == rec["logical_error"]


def test_threshold_sweeps_match_reference(golden):
    from quantum_sim.engine.qec import QECSimulator
    j, _ = golden
    codes = _codes()
    for sw in j["qec_sweeps"]:
        pts = QECSimulator(codes[sw["code"]]()).threshold_sweep(sw["probs"], sw["trials"], sw["noise_type"], sw["seed"])
        assert len(pts) == len(sw["points"])
        for got, want in zip(pts, sw["points"]):
            assert got.physical_rate == want["physical_rate"]
            for f in FIELDS:
                assert abs(getattr(got, f) - want[f]) < 1e-12, (sw["code"], sw["noise_type"], want["physical_rate"], f)


def test_projection_logical_error_consistent():
    from quantum_sim.engine.qec import QECSimulator, SteaneCode
    sim = QECSimulator(SteaneCode())
    r = sim.projection_logical_error(1, "depolarizing", 0.05, n_trials=24, seed=9)
    rng = np.random.default_rng(9)
    seeds = [int(rng.integers(0, 2 ** 63)) for _ in range(24)]
    fids = [sim.run_cycle(1, "depolarizing", 0.05, seed=s).fidelity_after for s in seeds]
    assert abs(r["mean_fidelity"] - sum(fids) / 24) < 1e-12
    assert r["n_trials"] == 24


def test_philox_threshold_sweep_agrees_with_the_reference_stream_sweep_statistically():
    """threshold_sweep_philox (counter-based bulk draws, the throughput mode of config 4) estimates the same quantities
    as the reference-stream sweep: with 6000 trials per point the two differ by sampling error only, and the
    sweep is reproducible and independent of the batch size used to key... (the key includes the batch index, so the
    batch size is part of the stream's definition -- a fixed batch gives bit-identical repeats)."""
    from quantum_sim.engine.qec import QECSimulator, SteaneCode
    sim = QECSimulator(SteaneCode())
    probs = [0.01, 0.08, 0.25]
    a = sim.threshold_sweep_philox(probs, n_trials=6000, noise_type="depolarizing", seed=5, batch=2048)
    b = sim.threshold_sweep_philox(probs, n_trials=6000, noise_type="depolarizing", seed=5, batch=2048)
    r = sim.threshold_sweep(probs, n_trials=1500, noise_type="depolarizing", seed=5)
    for x, y, z in zip(a, b, r):
        assert x == y
        assert x.physical_rate == z.physical_rate
        # binomial standard errors: sqrt(p(1-p)/6000) + sqrt(p(1-p)/1500) <= 0.02; allow 5 sigma
        assert abs(x.logical_rate - z.logical_rate) < 0.1 and abs(x.avg_fidelity - z.avg_fidelity) < 0.1
        assert abs(x.decoder_success_rate - z.decoder_success_rate) < 0.1
    assert a[0].logical_rate < a[1].logical_rate < a[2].logical_rate


def test_lean_batches_return_the_same_scalars():
    """`run_cycles(..., lean=True)` (the throughput sweep's mode: no per-trial Python lists) gives the same fidelity /
    <Z_L> / logical-error arrays as the full form, for every built-in code."""
    from quantum_sim.engine import qec as Q
    rng = np.random.default_rng(3)
    for code in (Q.BitFlipCode(), Q.PhaseFlipCode(), Q.SteaneCode()):
        sim = Q.QECSimulator(code)
        u = rng.random((300, code.data_qubits))
        lg = rng.integers(0, 2, 300)
        full = sim.run_cycles(lg.tolist(), "depolarizing", 0.15, None, uniforms=u)
        lean = sim.run_cycles(lg, "depolarizing", 0.15, None, uniforms=u, lean=True)
        for key in ("fidelity_before", "fidelity_after", "z_exp", "logical_error"):
            assert np.array_equal(np.asarray(full[key]), np.asarray(lean[key])), (type(code).__name__, key)
        assert lean["corrections"] is None and len(full["corrections"]) == 300
