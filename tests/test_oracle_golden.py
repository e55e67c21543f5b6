"""Pin the CPU oracle (oracle/qsim_oracle.py) to outputs of the real reference
engine frozen by tests/golden/make_golden.py.  CPU only."""

import os

import numpy as np
import pytest

from oracle import qsim_oracle as O
from conftest import as_gates, as_noise, GOLDEN_DIR

TOL = 1e-13   # oracle vs reference amplitudes (target for the CUDA path is 1e-12)


def test_ghz3(golden):
    j, a = golden
    g = as_gates(j["ghz3"]["gates"])
    psi = O.run_state(3, g)[0]
    assert np.max(np.abs(psi - a["ghz3_state"])) < TOL
    for b in "ZXY":
        _, counts, _ = O.run(3, g, shots=1024, seed=42, basis=b)
        assert counts == j["ghz3"][f"counts_{b}"], b
    _, counts, _ = O.run(3, g, noise={"global": [], "gate": {}, "readout": (0.02, 0.05)},
                         shots=1024, seed=42)
    assert counts == j["ghz3"]["counts_readout"]
    d = O.readout_distribution(O.probabilities(psi), 3, 0.02, 0.05)
    assert np.max(np.abs(d - a["ghz3_readout_dist"])) < 1e-15
    assert np.allclose(O.all_pairs_mi(psi, 3), j["ghz3"]["mi"], atol=1e-12)
    assert abs(O.entanglement_entropy(psi, 3, [0]) - j["ghz3"]["entropy_q0"]) < 1e-12
    c = O.run_with_noise(3, g, None, {"global": [("depolarizing", 0.1)]}, 7, 200, 42)
    assert c == j["ghz3"]["run_with_noise"]


def test_sigma_every_target_list(golden):
    j, a = golden
    n = j["sigma"]["n"]
    for i, t in enumerate(j["sigma"]["targets"]):
        k = len(t)
        u = a["sigma_mat"][i][:4 ** k].reshape(2 ** k, 2 ** k)
        out = O.apply_gate(a["sigma_in"][i], n, u, t)
        assert np.max(np.abs(out - a["sigma_out"][i])) < TOL, t
    for rec in j["bigk"]:
        out = O.apply_gate(a[rec["tag"] + "_in"], rec["n"], a[rec["tag"] + "_mat"], rec["targets"])
        assert np.max(np.abs(out - a[rec["tag"] + "_out"])) < TOL


def test_random_circuits_all_registry_gates(golden):
    j, a = golden
    for rec in j["random_circuits"]:
        psi, steps, _, _ = O.run_state(rec["n"], as_gates(rec["gates"]), rec["initial"],
                                       record_steps=True)
        assert np.max(np.abs(psi - a[rec["tag"]])) < TOL
        assert np.max(np.abs(np.array(steps) - a[rec["tag"] + "_steps"])) < TOL


def test_layered16_checksum(golden):
    from qsb.workloads import layered_circuit
    j, a = golden
    g = layered_circuit(16, 64, 2026)
    assert len(g) == j["layered16"]["n_gates"] == 683
    psi = O.run_state(16, g)[0]
    assert np.max(np.abs(psi[a["layered16_idx"]] - a["layered16_amps"])) < TOL
    assert abs(np.sum(np.abs(psi) ** 2) - j["layered16"]["norm2"]) < 1e-12
    assert int(np.argmax(np.abs(psi))) == j["layered16"]["argmax"] == 51629
    vals = np.random.default_rng(2027).uniform(-np.pi, np.pi, (4096, 473))
    psi = O.run_state(16, O.bind_values(g, vals[4095]))[0]
    assert np.max(np.abs(psi[a["layered16_idx"]] - a["layered16_bound4095_amps"])) < TOL


def test_noisy_trajectories(golden):
    j, a = golden
    for rec in j["noisy"]:
        g, noise = as_gates(rec["gates"]), as_noise(rec["noise"])
        draws = np.random.default_rng(rec["noise_seed"]).random(O.draw_count(rec["n"], g, noise))
        psi, steps, _, _ = O.run_state(rec["n"], g, None, noise, draws, record_steps=True)
        assert np.max(np.abs(psi - a[rec["tag"]])) < TOL, rec["tag"]
        assert np.max(np.abs(np.array(steps) - a[rec["tag"] + "_steps"])) < TOL


def test_config3_trajectories(golden):
    from qsb.workloads import layered_circuit, config3_noise
    j, a = golden
    g, noise = layered_circuit(12, 16, 2026), config3_noise()
    assert O.draw_count(12, g, noise) == 384
    for seed in j["cfg3_traj_seeds"]:
        draws = np.random.default_rng(seed).random(384)
        psi = O.run_state(12, g, None, noise, draws)[0]
        assert np.max(np.abs(psi - a[f"cfg3_traj_{seed}"])) < TOL


def test_ensemble_density_matrix(golden):
    j, a = golden
    for key in ("ens4", "ens5"):
        rec = j[key]
        n = 4 if key == "ens4" else 5
        rho = O.ensemble_density_matrix(n, as_gates(rec["gates"]), None, as_noise(rec["noise"]),
                                        rec["n_trials"], rec["seed"])
        assert np.max(np.abs(rho - a[key + "_rho"])) < TOL
    assert abs(O.purity_dm(a["ens4_rho"]) - j["ens4"]["purity"]) < 1e-14
    rho = O.ensemble_density_matrix(5, as_gates(j["ens5"]["gates"]), None, None, 3, 9)
    assert np.max(np.abs(rho - a["ens5_clean_rho"])) < TOL


def test_run_with_noise_counts(golden):
    j, _ = golden
    for rec in j["run_with_noise"]:
        c = O.run_with_noise(rec["n"], as_gates(rec["gates"]), None, as_noise(rec["noise"]),
                             rec["noise_seed"], rec["shots"], rec["seed"])
        assert c == rec["counts"]


def test_measurement_and_readout(golden):
    j, a = golden
    for rec in j["measurement"]:
        n, psi = rec["n"], a[rec["tag"]]
        for key, want in rec["counts"].items():
            basis, mode = key.split("_")
            ro = None if mode == "None" else (0.1, 0.07)
            got = O.sample_with_basis(psi, n, 500, basis, ro, "shot" if mode == "None" else mode,
                                      np.random.default_rng(77))
            assert got == want, key
        idx = O.measure_all_index(psi, np.random.default_rng(3).random())
        assert format(idx, f"0{n}b") == rec["measure_all"]
        rdm = np.array([O.reduced_density_matrix_1q(psi, n, q) for q in range(n)])
        assert np.max(np.abs(rdm - a[rec["tag"] + "_rdm1"])) < TOL
    out = O.readout_distribution(a["readout8_in"], 8, 0.03, 0.11)
    assert np.max(np.abs(out - a["readout8_out"])) < 1e-15
    p = np.random.default_rng(1007).random(2 ** 16)
    p /= p.sum()
    out = O.readout_distribution(p, 16, 0.02, 0.05)
    assert np.max(np.abs(out[a["readout16_idx"]] - a["readout16_out"])) < 1e-15


def test_analysis(golden):
    j, a = golden
    for rec in j["analysis"]:
        n, psi = rec["n"], a[rec["tag"]]
        assert np.allclose(O.all_pairs_mi(psi, n), rec["mi"], atol=1e-11)
        rdm2 = np.array([O.partial_trace(psi, n, [i, k]) for i in range(n) for k in range(i + 1, n)])
        assert np.max(np.abs(rdm2 - a[rec["tag"] + "_rdm2"])) < TOL
        assert np.allclose([O.entanglement_entropy(psi, n, [q]) for q in range(n)],
                           rec["entropy_1q"], atol=1e-12)
        for ev in rec["expect"]:
            obs = np.array([[1]], dtype=complex)
            for ch in ev["label"]:
                obs = np.kron(obs, O.gate_matrix(ch))
            v = O.expectation_value(psi, n, obs, ev["qubits"])
            assert abs(v - complex(ev["re"], ev["im"])) < TOL, ev
        assert abs(O.state_fidelity(psi, a[rec["tag"] + "_phi"]) - rec["fidelity"]) < TOL
    from qsb.workloads import ghz
    _, steps, _, _ = O.run_state(4, ghz(4), record_steps=True)
    got = [O.all_pairs_mi(s, 4) for s in steps]
    assert np.allclose(got, j["ghz4_layer_mi"], atol=1e-12)


def test_qec_cycles_and_sweeps(golden):
    j, a = golden
    assert np.max(np.abs(O.steane_encode(0) - a["steane_enc0"])) == 0
    assert np.max(np.abs(O.steane_encode(1) - a["steane_enc1"])) == 0
    for rec in j["qec_cycles"]:
        r = O.qec_cycle(rec["code"], rec["logical"], rec["noise_type"], rec["p"], rec["seed"])
        assert r["syndrome"] == rec["syndrome"], rec
        assert [list(c) for c in r["corrections"]] == rec["corrections"]
        assert abs(r["fidelity_before"] - rec["fidelity_before"]) < 1e-12
        assert abs(r["fidelity_after"] - rec["fidelity_after"]) < 1e-12
        assert abs(r["z_exp"] - rec["z_exp"]) < 1e-12
        assert r["logical_error"] == rec["logical_error"]
    for sw in j["qec_sweeps"]:
        pts = O.threshold_sweep(sw["code"], sw["probs"], sw["trials"], sw["noise_type"], sw["seed"])
        for got, want in zip(pts, sw["points"]):
            for k, v in want.items():
                assert abs(got[k] - v) < 1e-12, (sw["code"], k)


def test_vqe_gradient_shape(golden):
    j, _ = golden
    rec = j["vqe_grad"]
    g, vals = as_gates(rec["gates"]), np.array(rec["values"])
    terms = [(c, l, q) for c, l, q in rec["terms"]]
    cost = lambda v: O.vqe_cost(O.run_state(4, O.bind_values(g, v))[0], 4, terms)
    assert abs(cost(vals) - rec["cost"]) < 1e-12
    grad = []
    for i in range(len(vals)):
        p, m = vals.copy(), vals.copy()
        p[i] += np.pi / 2
        m[i] -= np.pi / 2
        grad.append((cost(p) - cost(m)) / (2 * np.sin(np.pi / 2)))
    assert np.allclose(grad, rec["grad"], atol=1e-12)


def test_config3_rho_golden_if_present():
    path = os.path.join(GOLDEN_DIR, "golden_cfg3_rho.npz")
    if not os.path.exists(path):
        pytest.skip("slow golden not generated")
    # covered on the GPU path (tests/test_gpu_parity.py); here only check the fixture is sane
    s = np.load(path)
    assert abs(s["cfg3_rho_diag"].sum() - 1.0) < 1e-10


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert O.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
                           [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = O.philox_uniforms(12345, 7, 9)
    assert np.all((u >= 0) & (u < 1)) and len(set(u.tolist())) == 9
