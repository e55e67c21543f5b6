"""Freeze outputs of the REAL reference engine into tests/golden/.

Run in the build container only (the reference lives at /root/reference and
does not travel to the GPU box):

    python tests/golden/make_golden.py            # fast cases  (~1 min)
    python tests/golden/make_golden.py --slow     # + config-3 rho (12 q, 500 trials)

It imports `quantum_sim.engine` from /root/reference, runs the calls named in
SURVEY.md section 8 on seeded inputs and stores inputs and outputs:
  golden.json  scalars, counts, gate lists, syndromes
  golden.npz   amplitude vectors / matrices (complex128, float64)
"""

from __future__ import annotations

import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")                               # the reference's quantum_sim wins ...
sys.path.append(os.path.join(ROOT, "quantum-simulator_b200"))       # ... ours only provides qsb.workloads

from quantum_sim.engine.circuit import QuantumCircuit, GateInstance          # noqa: E402
from quantum_sim.engine.state_vector import StateVector                      # noqa: E402
from quantum_sim.engine.simulator import Simulator                           # noqa: E402
from quantum_sim.engine.noise import (NoiseModel, BitFlipNoise, PhaseFlipNoise,  # noqa: E402
                                      DepolarizingNoise, AmplitudeDampingNoise, ReadoutError)
from quantum_sim.engine.measurement import MeasurementEngine, MeasurementBasis  # noqa: E402
from quantum_sim.engine.analysis import StateAnalysis                        # noqa: E402
from quantum_sim.engine.gate_registry import GateRegistry                    # noqa: E402
from quantum_sim.engine.qec import QECSimulator, SteaneCode, BitFlipCode, PhaseFlipCode  # noqa: E402
from quantum_sim.engine.optimizer import (ParameterizedCircuitConfig, CostFunction,  # noqa: E402
                                          GradientEstimator)

from qsb.workloads import layered_circuit, ghz, config3_noise                # noqa: E402

assert "/root/reference" in sys.modules["quantum_sim"].__file__, "wrong quantum_sim on path"

J = {}   # json payload
A = {}   # npz payload

NOISE_CLS = {"bit_flip": BitFlipNoise, "phase_flip": PhaseFlipNoise,
             "depolarizing": DepolarizingNoise, "amplitude_damping": AmplitudeDampingNoise}


def circuit(n, gates, initial=None):
    qc = QuantumCircuit(n, initial_states=list(initial) if initial else [])
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    return qc


def noise_model(spec, seed=None):
    nm = NoiseModel()
    for kind, p in spec.get("global", []):
        nm.add_global_noise(NOISE_CLS[kind](p))
    for name, chans in spec.get("gate", {}).items():
        for kind, p in chans:
            nm.add_gate_noise(name, NOISE_CLS[kind](p))
    if spec.get("readout"):
        nm.set_readout_error(ReadoutError(*spec["readout"]))
    if seed is not None:
        nm.set_seed(seed)
    return nm


def rand_state(rng, n):
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    return v / np.linalg.norm(v)


def rand_unitary(rng, dim):
    q, r = np.linalg.qr(rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def sv_from(psi, n):
    sv = StateVector(n)
    sv.data = psi
    return sv


def rand_circuit(rng, n, n_gates, names):
    reg = GateRegistry.instance()
    gates = []
    for i in range(n_gates):
        while True:
            name = names[int(rng.integers(0, len(names)))]
            gd = reg.get(name)
            if gd.num_qubits <= n:
                break
        targets = rng.permutation(n)[:gd.num_qubits].tolist()
        params = [float(x) for x in rng.uniform(-np.pi, np.pi, gd.num_params)]
        gates.append((name, targets, params, i // 3))
    return gates


# ---------------------------------------------------------------- config 1: GHZ-3
def case_ghz3():
    g = ghz(3)
    qc = circuit(3, g)
    res = Simulator().run(qc, shots=0)
    A["ghz3_state"] = res.final_state.data
    out = {"gates": g}
    for b in "ZXY":
        r = Simulator().run(qc, shots=1024, seed=42, measurement_basis=MeasurementBasis[b])
        out[f"counts_{b}"] = r.measurement_counts
    nm = noise_model({"readout": (0.02, 0.05)})
    out["counts_readout"] = Simulator(nm).run(qc, shots=1024, seed=42).measurement_counts
    A["ghz3_readout_dist"] = ReadoutError(0.02, 0.05).apply_to_distribution(
        res.final_state.probabilities, 3)
    out["mi"] = [StateAnalysis.mutual_information(res.final_state, i, j)
                 for i in range(3) for j in range(i + 1, 3)]
    out["entropy_q0"] = StateAnalysis.entanglement_entropy(res.final_state, [0])
    nm = noise_model({"global": [("depolarizing", 0.1)]}, seed=7)
    out["run_with_noise"] = Simulator(nm).run_with_noise(qc, shots=200, seed=42).measurement_counts
    J["ghz3"] = out


# ---------------------------------------------------------------- sigma cases
def case_sigma():
    rng = np.random.default_rng(1001)
    n = 5
    cases = []
    ins, mats, outs = [], [], []
    for k in (1, 2, 3):
        for targets in itertools.permutations(range(n), k):
            psi = rand_state(rng, n)
            u = rand_unitary(rng, 2 ** k)
            sv = sv_from(psi, n)
            sv.apply_gate(u, list(targets))
            cases.append(list(targets))
            ins.append(psi)
            mats.append(np.pad(u.reshape(-1), (0, 64 - u.size)))
            outs.append(sv.data)
    J["sigma"] = {"n": n, "targets": cases}
    A["sigma_in"] = np.array(ins)
    A["sigma_mat"] = np.array(mats)
    A["sigma_out"] = np.array(outs)
    # 4- and 5-qubit dense operators (expectation_value passes kron strings of any length)
    big = []
    for k, targets in ((4, [3, 0, 5, 1]), (5, [2, 6, 0, 4, 1]), (4, [0, 1, 2, 3])):
        psi = rand_state(rng, 7)
        u = rand_unitary(rng, 2 ** k)
        sv = sv_from(psi, 7)
        sv.apply_gate(u, targets)
        tag = f"bigk_{len(big)}"
        A[tag + "_in"], A[tag + "_mat"], A[tag + "_out"] = psi, u, sv.data
        big.append({"n": 7, "targets": targets, "tag": tag})
    J["bigk"] = big


# ---------------------------------------------------------------- random circuits
def case_random_circuits():
    rng = np.random.default_rng(1002)
    names = [g for g in GateRegistry.instance().gate_names()]
    out = []
    for n in range(3, 11):
        for rep in range(3):
            gates = rand_circuit(rng, n, 30, names)
            initial = [int(x) for x in rng.integers(0, 2, n)]
            qc = circuit(n, gates, initial)
            res = Simulator().run(qc, shots=0, record_steps=True)
            tag = f"rc_{n}_{rep}"
            A[tag] = res.final_state.data
            A[tag + "_steps"] = np.array([s.data for s in res.step_states])
            out.append({"n": n, "gates": gates, "initial": initial, "tag": tag})
    J["random_circuits"] = out


# ---------------------------------------------------------------- config 2 checksum
def case_layered16():
    g = layered_circuit(16, 64, 2026)
    t0 = time.perf_counter()
    res = Simulator().run(circuit(16, g), shots=0)
    dt = time.perf_counter() - t0
    psi = res.final_state.data
    idx = np.random.default_rng(5).integers(0, 2 ** 16, 256)
    A["layered16_idx"] = idx
    A["layered16_amps"] = psi[idx]
    J["layered16"] = {"norm2": float(np.sum(np.abs(psi) ** 2)), "a0": [psi[0].real, psi[0].imag],
                      "argmax": int(np.argmax(np.abs(psi))), "n_gates": len(g),
                      "ref_seconds_here": dt}
    # one bound parameter set of config 2's batch
    cfg = ParameterizedCircuitConfig.auto_detect(circuit(16, g))
    vals = np.random.default_rng(2027).uniform(-np.pi, np.pi, (4096, cfg.num_params))
    for row in (0, 4095):
        psi = Simulator().run(cfg.bind_values(vals[row]), shots=0).final_state.data
        A[f"layered16_bound{row}_amps"] = psi[idx]
    J["layered16"]["num_params"] = cfg.num_params


# ---------------------------------------------------------------- noisy trajectories
NOISY_SPECS = [
    {"global": [("depolarizing", 0.2)]},
    {"global": [("amplitude_damping", 0.3)]},
    {"global": [("bit_flip", 0.25), ("phase_flip", 0.15)]},
    {"global": [("depolarizing", 0.05), ("amplitude_damping", 0.1)],
     "gate": {"CNOT": [("bit_flip", 0.3)], "H": [("amplitude_damping", 0.5)]}},
    {"global": [("amplitude_damping", 1.0)]},
    {"global": [("depolarizing", 1.0)]},
]


def case_noisy():
    rng = np.random.default_rng(1003)
    names = ["H", "X", "Y", "S", "T", "Rx", "Ry", "Rz", "U3", "CNOT", "CZ", "SWAP", "Toffoli", "Fredkin"]
    out = []
    for ci, spec in enumerate(NOISY_SPECS):
        for n in (3, 5, 6):
            gates = rand_circuit(rng, n, 12, names)
            qc = circuit(n, gates)
            for seed in (11, 12, 13):
                nm = noise_model(spec, seed=seed)
                res = Simulator(nm).run(qc, shots=0, record_steps=True)
                tag = f"noisy_{ci}_{n}_{seed}"
                A[tag] = res.final_state.data
                A[tag + "_steps"] = np.array([s.data for s in res.step_states])
                out.append({"n": n, "gates": gates, "noise": spec, "noise_seed": seed, "tag": tag})
    J["noisy"] = out
    # 12-qubit config-3 trajectories (2 seeds), sampled amplitudes only
    g = layered_circuit(12, 16, 2026)
    qc = circuit(12, g)
    idx = np.random.default_rng(6).integers(0, 2 ** 12, 128)
    A["cfg3_idx"] = idx
    c3 = []
    for seed in (101, 102):
        nm = noise_model(config3_noise(), seed=seed)
        psi = Simulator(nm).run(qc, shots=0).final_state.data
        A[f"cfg3_traj_{seed}"] = psi
        c3.append(seed)
    J["cfg3_traj_seeds"] = c3


# ---------------------------------------------------------------- ensemble rho / run_with_noise
def case_ensemble():
    g = [("H", [0], [], 0), ("CNOT", [0, 1], [], 1), ("CNOT", [1, 2], [], 2), ("CNOT", [2, 3], [], 3)]
    nm = noise_model({"global": [("depolarizing", 0.05)]})
    rho = Simulator(nm).ensemble_density_matrix(circuit(4, g), 50, seed=42)
    A["ens4_rho"] = rho
    J["ens4"] = {"gates": g, "noise": {"global": [("depolarizing", 0.05)]}, "n_trials": 50, "seed": 42,
                 "purity": StateAnalysis.purity_dm(rho), "trace": float(np.real(np.trace(rho)))}
    rng = np.random.default_rng(1004)
    g5 = rand_circuit(rng, 5, 15, ["H", "Rx", "Ry", "U3", "CNOT", "CZ", "Toffoli"])
    spec = {"global": [("depolarizing", 0.05), ("amplitude_damping", 0.1)]}
    rho = Simulator(noise_model(spec)).ensemble_density_matrix(circuit(5, g5), 40, seed=9)
    A["ens5_rho"] = rho
    J["ens5"] = {"gates": g5, "noise": spec, "n_trials": 40, "seed": 9}
    rho = Simulator().ensemble_density_matrix(circuit(5, g5), 3, seed=9)    # no noise model
    A["ens5_clean_rho"] = rho
    # run_with_noise
    rw = []
    for spec, ns, seed, shots in ((spec, 21, 5, 150), ({"global": [("amplitude_damping", 0.4)]}, 3, 8, 100)):
        nm = noise_model(spec, seed=ns)
        cnt = Simulator(nm).run_with_noise(circuit(5, g5), shots=shots, seed=seed).measurement_counts
        rw.append({"n": 5, "gates": g5, "noise": spec, "noise_seed": ns, "seed": seed,
                   "shots": shots, "counts": cnt})
    J["run_with_noise"] = rw


# ---------------------------------------------------------------- measurement / readout
def case_measurement():
    rng = np.random.default_rng(1005)
    out = []
    for n in (2, 4, 7):
        psi = rand_state(rng, n)
        tag = f"meas_{n}"
        A[tag] = psi
        rec = {"n": n, "tag": tag, "counts": {}}
        for basis in "ZXY":
            for mode in ("shot", "distribution", None):
                ro = ReadoutError(0.1, 0.07) if mode else None
                cnt = MeasurementEngine.sample_with_basis(
                    sv_from(psi, n), 500, basis=MeasurementBasis[basis], readout_error=ro,
                    readout_mode=mode or "shot", rng=np.random.default_rng(77))
                rec["counts"][f"{basis}_{mode}"] = cnt
        sv = sv_from(psi, n)
        rec["measure_all"] = sv.measure_all(np.random.default_rng(3))
        sv = sv_from(psi, n)
        r = np.random.default_rng(4)
        rec["measure_qubit"] = [sv.measure_qubit(q, r) for q in range(n)]
        A[tag + "_after_mq"] = sv.data
        rec["bloch"] = [list(sv_from(psi, n).get_bloch_coordinates(q)) for q in range(n)]
        A[tag + "_rdm1"] = np.array([sv_from(psi, n).get_reduced_density_matrix(q) for q in range(n)])
        out.append(rec)
    J["measurement"] = out
    p = np.random.default_rng(1006).random(2 ** 8)
    p /= p.sum()
    A["readout8_in"] = p
    A["readout8_out"] = ReadoutError(0.03, 0.11).apply_to_distribution(p.copy(), 8)
    p = np.random.default_rng(1007).random(2 ** 16)
    p /= p.sum()
    q = ReadoutError(0.02, 0.05).apply_to_distribution(p.copy(), 16)
    idx = np.arange(0, 2 ** 16, 257)
    A["readout16_idx"], A["readout16_out"] = idx, q[idx]


# ---------------------------------------------------------------- analysis
def case_analysis():
    rng = np.random.default_rng(1008)
    out = []
    for n in (2, 4, 6, 8):
        psi = rand_state(rng, n)
        tag = f"ana_{n}"
        A[tag] = psi
        sv = sv_from(psi, n)
        rec = {"n": n, "tag": tag}
        rec["mi"] = [StateAnalysis.mutual_information(sv, i, j) for i in range(n) for j in range(i + 1, n)]
        A[tag + "_rdm2"] = np.array([StateAnalysis.partial_trace(sv, [i, j])
                                     for i in range(n) for j in range(i + 1, n)])
        rec["entropy_1q"] = [StateAnalysis.entanglement_entropy(sv, [q]) for q in range(n)]
        ev = []
        for label, qubits in (("Z", [0]), ("X", [n - 1]), ("Y", [n // 2]), ("ZZ", [0, n - 1]),
                              ("XY", [n - 1, 0]), ("YZ", [n // 2, 0])) + \
                             ((("XYZ", [n - 1, 0, 1]), ("ZXYY", [3, 1, 0, 2])) if n >= 4 else ()):
            obs = np.array([[1]], dtype=complex)
            from quantum_sim.engine.gates import X_MATRIX, Y_MATRIX, Z_MATRIX
            pm = {"X": X_MATRIX, "Y": Y_MATRIX, "Z": Z_MATRIX}
            for ch in label:
                obs = np.kron(obs, pm[ch])
            v = StateAnalysis.expectation_value(sv, obs, qubits)
            ev.append({"label": label, "qubits": qubits, "re": v.real, "im": v.imag})
        rec["expect"] = ev
        phi = rand_state(rng, n)
        A[tag + "_phi"] = phi
        rec["fidelity"] = StateAnalysis.state_fidelity(psi, phi)
        out.append(rec)
    J["analysis"] = out
    # per-layer MI of a GHZ-4 run (record_steps semantics of config 3)
    res = Simulator().run(circuit(4, ghz(4)), shots=0, record_steps=True)
    J["ghz4_layer_mi"] = [[StateAnalysis.mutual_information(s, i, j) for i in range(4) for j in range(i + 1, 4)]
                          for s in res.step_states]


# ---------------------------------------------------------------- QEC
def case_qec():
    out = []
    codes = {"steane": SteaneCode, "bit_flip": BitFlipCode, "phase_flip": PhaseFlipCode}
    for code, ntype, p, seeds in (("steane", "depolarizing", 0.15, range(200, 212)),
                                  ("steane", "bit_flip", 0.2, range(300, 304)),
                                  ("steane", "phase_flip", 0.2, range(400, 404)),
                                  ("bit_flip", "bit_flip", 0.3, range(500, 506)),
                                  ("phase_flip", "phase_flip", 0.3, range(600, 606)),
                                  ("phase_flip", "depolarizing", 0.4, range(700, 704))):
        sim = QECSimulator(codes[code]())
        for s in seeds:
            for logical in (0, 1):
                r = sim.run_cycle(logical, ntype, p, seed=s)
                out.append({"code": code, "noise_type": ntype, "p": p, "seed": s, "logical": logical,
                            "syndrome": r.syndrome, "corrections": [list(c) for c in r.correction_applied],
                            "fidelity_before": r.fidelity_before, "fidelity_after": r.fidelity_after,
                            "z_exp": r.logical_z_expectation, "logical_error": bool(r.logical_error_detected)})
    J["qec_cycles"] = out
    A["steane_enc0"] = SteaneCode().encode(0).data
    A["steane_enc1"] = SteaneCode().encode(1).data
    sweeps = []
    probs = np.linspace(0.001, 0.3, 15).tolist()
    for code, ntype, trials in (("steane", "depolarizing", 20), ("bit_flip", "bit_flip", 40),
                                ("phase_flip", "phase_flip", 40), ("bit_flip", "depolarizing", 30)):
        t0 = time.perf_counter()
        pts = QECSimulator(codes[code]()).threshold_sweep(probs, trials, ntype, 42)
        sweeps.append({"code": code, "noise_type": ntype, "trials": trials, "seed": 42, "probs": probs,
                       "ref_seconds_here": time.perf_counter() - t0,
                       "points": [vars(pt) for pt in pts]})
    J["qec_sweeps"] = sweeps


# ---------------------------------------------------------------- VQE shape
def case_vqe():
    sys.path.insert(0, "/root/reference/scripts")
    import vqe_benchmark
    r = vqe_benchmark.run_benchmark(3, 2, "heisenberg", 0.1, 5, 42)
    J["vqe_script"] = {k: r[k] for k in ("n_params", "cost_trace", "optimal_cost", "actual_iterations")}
    J["vqe_script"]["args"] = {"qubits": 3, "layers": 2, "hamiltonian": "heisenberg", "lr": 0.1,
                               "iters": 5, "seed": 42}
    # batch shape: parameter-shift gradient of a 4-qubit layered circuit, ZZ chain cost
    g = layered_circuit(4, 3, 7)
    cfg = ParameterizedCircuitConfig.auto_detect(circuit(4, g))
    vals = np.random.default_rng(8).uniform(-np.pi, np.pi, cfg.num_params)
    terms = [(-1.0, "ZZ", [i, i + 1]) for i in range(3)] + [(0.5, "XY", [3, 0]), (0.25, "Y", [2])]
    cost = CostFunction.vqe_hamiltonian(terms)
    grad = GradientEstimator.parameter_shift(cfg, cost, vals)
    c0 = cost(Simulator().run(cfg.bind_values(vals), shots=0).final_state)
    J["vqe_grad"] = {"gates": g, "values": vals.tolist(), "terms": terms, "grad": grad.tolist(), "cost": c0}


# ---------------------------------------------------------------- config 3 rho (slow)
def case_cfg3_rho():
    g = layered_circuit(12, 16, 2026)
    nm = noise_model(config3_noise())
    t0 = time.perf_counter()
    rho = Simulator(nm).ensemble_density_matrix(circuit(12, g), 500, seed=42)
    dt = time.perf_counter() - t0
    idx = np.random.default_rng(9).integers(0, 4096, (512, 2))
    S = {"cfg3_rho_diag": np.real(np.diag(rho)).copy(), "cfg3_rho_idx": idx,
         "cfg3_rho_samples": rho[idx[:, 0], idx[:, 1]], "cfg3_rho_row7": rho[7].copy()}
    meta = {"trace": float(np.real(np.trace(rho))), "purity": float(np.real(np.vdot(rho, rho))),
            "ref_seconds_here": dt, "n_trials": 500, "seed": 42}
    np.savez_compressed(os.path.join(HERE, "golden_cfg3_rho.npz"), **S)
    with open(os.path.join(HERE, "golden_cfg3_rho.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("cfg3 rho", meta)


# ---------------------------------------------------------------- debugger
def case_debugger():
    """CircuitDebugger: step snapshots, noise impact, noise attribution, state diff -> golden_debugger.{json,npz}."""
    from quantum_sim.engine.debugger import CircuitDebugger
    JD, AD = {}, {}
    g = layered_circuit(5, 4, 31)
    spec = {"global": [("depolarizing", 0.05), ("amplitude_damping", 0.1)], "gate": {"CNOT": [("bit_flip", 0.2)]}}
    JD["gates"], JD["noise"], JD["n"] = g, spec, 5
    dbg = CircuitDebugger()
    nm = noise_model(spec, seed=11)
    snaps = dbg.run_full_debug(circuit(5, g), nm, seed=3)
    JD["full_debug"] = {"noise_seed": 11, "columns": [s.column_index for s in snaps], "labels": [s.gate_labels for s in snaps],
                        "fidelity": [s.fidelity for s in snaps], "cumulative_fidelity": [s.cumulative_fidelity for s in snaps],
                        "entropy": [s.entropy for s in snaps]}
    AD["dbg_states"] = np.stack([s.state.data for s in snaps])
    AD["dbg_ideal"] = np.stack([s.ideal_state.data for s in snaps])
    d = CircuitDebugger.compute_state_diff(snaps[1], snaps[-1])
    JD["state_diff"] = {"fidelity": d["fidelity"], "tvd": d["tvd"], "entropy_diff": d["entropy_diff"],
                        "amplitude_diffs": [[a[0], a[1], a[2].real, a[2].imag, a[3].real, a[3].imag, a[4]] for a in d["amplitude_diffs"]]}
    AD["dbg_prob_diffs"] = d["prob_diffs"]
    snaps0 = CircuitDebugger().run_full_debug(circuit(5, g), None, seed=3)
    JD["full_debug_noiseless"] = {"fidelity": [s.fidelity for s in snaps0], "entropy": [s.entropy for s in snaps0]}
    AD["dbg_states_noiseless"] = np.stack([s.state.data for s in snaps0])
    imp = CircuitDebugger().compute_noise_impact(circuit(5, g), noise_model(spec), n_trials=7, seed=5)
    JD["noise_impact"] = [vars(r) for r in imp]
    att = CircuitDebugger().compute_noise_attribution(circuit(5, g), noise_model(spec), n_trials=7, seed=6)
    JD["noise_attribution"] = vars(att)
    with open(os.path.join(HERE, "golden_debugger.json"), "w") as f:
        json.dump(JD, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_debugger.npz"), **AD)
    print("debugger golden written")


# ---------------------------------------------------------------- edge cases (golden_edges.*)
def case_edges():
    """Smallest / largest registers, empty and measure-only circuits, shots = 0, probabilities 0 and 1, collapse
    sequences, error texts."""
    JE, AE = {}, {}

    def run_case(tag, n, gates, initial=None, shots=64, seed=5, noise=None, noise_seed=None, basis="Z"):
        qc = circuit(n, gates, initial)
        nm = noise_model(noise, noise_seed) if noise else None
        res = Simulator(nm).run(qc, shots=shots, seed=seed, record_steps=True, measurement_basis=MeasurementBasis[basis])
        AE[tag] = res.final_state.data
        JE[tag] = {"n": n, "gates": gates, "initial": initial, "shots": shots, "seed": seed, "noise": noise,
                   "noise_seed": noise_seed, "basis": basis, "counts": res.measurement_counts,
                   "n_steps": len(res.step_states) if res.step_states is not None else None}
        if res.step_states:
            AE[tag + "_steps"] = np.array([s.data for s in res.step_states])

    run_case("one_qubit", 1, [("H", [0], [], 0), ("Rz", [0], [0.3], 1), ("Measure", [0], [], 2)], shots=100)
    run_case("one_qubit_y", 1, [("Rx", [0], [1.1], 0)], shots=50, basis="Y")
    run_case("empty", 4, [], initial=[1, 0, 1, 1], shots=10)
    run_case("empty_shots0", 3, [], shots=0)
    run_case("measure_only", 3, [("Measure", [1], [], 0), ("Barrier", [0], [], 1)], initial=[0, 1, 0], shots=7)
    run_case("measure_shots0", 2, [("H", [0], [], 0), ("Measure", [0], [], 1)], shots=0)
    ghz16 = [("H", [0], [], 0)] + [("CNOT", [0, q], [], q) for q in range(1, 16)]
    run_case("ghz16", 16, ghz16, shots=200, seed=9)
    run_case("max16_layer", 16, layered_circuit(16, 3, 5), initial=[i % 2 for i in range(16)], shots=32, seed=3)
    run_case("sparse_columns", 3, [("X", [2], [], 5), ("H", [0], [], 5), ("CNOT", [0, 1], [], 40), ("Y", [1], [], 2)], shots=20)
    run_case("p0_noise", 3, ghz(3), noise={"global": [("depolarizing", 0.0), ("amplitude_damping", 0.0), ("bit_flip", 0.0)]},
             noise_seed=1, shots=30)
    run_case("p1_bitflip", 3, ghz(3), noise={"global": [("bit_flip", 1.0)]}, noise_seed=2, shots=30)
    run_case("p1_phaseflip", 2, [("H", [0], [], 0), ("H", [1], [], 0)], noise={"global": [("phase_flip", 1.0)]}, noise_seed=2, shots=30)
    run_case("gamma1_damping", 3, [("X", [0], [], 0), ("H", [1], [], 0), ("CNOT", [1, 2], [], 1)],
             noise={"global": [("amplitude_damping", 1.0)]}, noise_seed=3, shots=30)
    run_case("p1_depol", 2, [("H", [0], [], 0), ("CNOT", [0, 1], [], 1)], noise={"global": [("depolarizing", 1.0)]},
             noise_seed=4, shots=30)
    run_case("gate_noise_only", 3, ghz(3) + [("H", [2], [], 5)], noise={"gate": {"H": [("bit_flip", 0.5)], "CNOT": [("amplitude_damping", 0.5)]}},
             noise_seed=6, shots=30)
    run_case("readout_extremes", 2, [("X", [0], [], 0)], noise={"readout": (1.0, 0.0)}, shots=16)

    # run_with_noise with shots = 0 and 1
    qc = circuit(3, ghz(3))
    for shots in (0, 1, 5):
        nm = noise_model({"global": [("depolarizing", 0.2)]}, 11)
        r = Simulator(nm).run_with_noise(qc, shots=shots, seed=12)
        JE[f"rwn_{shots}"] = {"counts": r.measurement_counts, "num_shots": r.num_shots}
    # ensemble with 0 / 1 trials, with and without noise
    for tag, nm in (("ens_nonoise", None), ("ens_noise", noise_model({"global": [("depolarizing", 0.3)]}, 1))):
        for trials in (1, 3):
            AE[f"{tag}_{trials}"] = Simulator(nm).ensemble_density_matrix(qc, n_trials=trials, seed=8)

    # collapse sequences
    rng = np.random.default_rng(77)
    sv = StateVector(4)
    sv.data = rand_state(np.random.default_rng(1), 4)
    seq = []
    for q in (2, 0, 3, 1, 2):
        seq.append(int(sv.measure_qubit(q, rng)))
        AE[f"collapse_{len(seq)}"] = sv.data.copy()
    JE["collapse_outcomes"] = seq
    sv = StateVector(3)
    sv.data = rand_state(np.random.default_rng(2), 3)
    JE["measure_all"] = sv.measure_all(np.random.default_rng(3))
    AE["measure_all_state"] = sv.data.copy()
    sv = StateVector.from_initial_states([1, 1, 0, 1])
    AE["from_initial"] = sv.data.copy()
    sv.reset()
    AE["reset_default"] = sv.data.copy()
    sv.reset([0, 1, 1, 0])
    AE["reset_states"] = sv.data.copy()
    JE["bloch"] = [list(map(float, StateVector.from_initial_states([0, 1]).get_bloch_coordinates(q))) for q in (0, 1)]

    # error texts
    errs = {}
    def err(tag, fn):
        try:
            fn()
            errs[tag] = None
        except Exception as e:
            errs[tag] = [type(e).__name__, str(e)]
    err("sv_0", lambda: StateVector(0))
    err("sv_17", lambda: StateVector(17))
    err("data_shape", lambda: setattr(StateVector(2), "data", np.zeros(3, dtype=complex)))
    err("qubit_range", lambda: StateVector(2).apply_gate(np.eye(2, dtype=complex), [2]))
    err("qubit_negative", lambda: StateVector(2).apply_gate(np.eye(2, dtype=complex), [-1]))
    err("unknown_gate", lambda: Simulator().run(circuit(2, [("Nope", [0], [], 0)]), shots=0))
    err("noise_p", lambda: DepolarizingNoise(1.5))
    err("noise_p_neg", lambda: AmplitudeDampingNoise(-0.1))
    err("readout_p", lambda: ReadoutError(2.0, 0.0))
    JE["errors"] = errs
    JE["_meta"] = {"numpy": np.__version__, "generator": "tests/golden/make_golden.py --edges"}
    with open(os.path.join(HERE, "golden_edges.json"), "w") as f:
        json.dump(JE, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_edges.npz"), **AE)
    print("edge-case golden written:", len(JE), "json entries,", len(AE), "arrays")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--debugger", action="store_true", help="only (re)write golden_debugger.*")
    ap.add_argument("--edges", action="store_true", help="only (re)write golden_edges.*")
    ap.add_argument("--slow", action="store_true")
    ap.add_argument("--only-slow", action="store_true")
    args = ap.parse_args()
    if args.debugger:
        case_debugger()
        return
    if args.edges:
        case_edges()
        return
    if not args.only_slow:
        for fn in (case_ghz3, case_sigma, case_random_circuits, case_layered16, case_noisy,
                   case_ensemble, case_measurement, case_analysis, case_qec, case_vqe):
            t0 = time.perf_counter()
            fn()
            print(f"{fn.__name__}: {time.perf_counter() - t0:.1f}s")
        J["_meta"] = {"numpy": np.__version__, "generator": "tests/golden/make_golden.py",
                      "reference": "justinbrianhwang/Quantum-Simulator @ /root/reference"}
        with open(os.path.join(HERE, "golden.json"), "w") as f:
            json.dump(J, f, indent=1)
        np.savez_compressed(os.path.join(HERE, "golden.npz"), **A)
    if args.slow or args.only_slow:
        case_cfg3_rho()


if __name__ == "__main__":
    main()
