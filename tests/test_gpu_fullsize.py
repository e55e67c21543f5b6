"""BASELINE.json configs at their FULL sizes on the device.

Config 3's ensemble rho is compared with entries the real reference produced (500 trials, 152 s of NumPy here:
tests/golden/make_golden.py case_cfg3_rho).  Where a full-size oracle run is not affordable the tests use
size-independent properties: unit norms, Hermiticity / trace / purity, circuit followed by its inverse."""

import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _circuit(n, gates):
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    qc = QuantumCircuit(n)
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    return qc


def test_config3_ensemble_rho_500_trials_matches_reference():
    from qsb.workloads import layered_circuit
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise, ReadoutError
    from quantum_sim.engine.simulator import Simulator
    s = np.load(os.path.join(GOLDEN_DIR, "golden_cfg3_rho.npz"))
    meta = json.load(open(os.path.join(GOLDEN_DIR, "golden_cfg3_rho.json")))
    nm = NoiseModel()
    nm.add_global_noise(DepolarizingNoise(0.01))
    nm.add_global_noise(AmplitudeDampingNoise(0.02))
    nm.set_readout_error(ReadoutError(0.02, 0.05))
    rho = Simulator(nm).ensemble_density_matrix(_circuit(12, layered_circuit(12, 16, 2026)), meta["n_trials"], seed=meta["seed"])
    assert rho.shape == (4096, 4096)
    assert np.max(np.abs(np.real(np.diag(rho)) - s["cfg3_rho_diag"])) < 1e-12
    idx = s["cfg3_rho_idx"]
    assert np.max(np.abs(rho[idx[:, 0], idx[:, 1]] - s["cfg3_rho_samples"])) < 1e-12
    assert np.max(np.abs(rho[7] - s["cfg3_rho_row7"])) < 1e-12
    assert abs(np.real(np.trace(rho)) - meta["trace"]) < 1e-12
    assert abs(np.real(np.vdot(rho, rho)) - meta["purity"]) < 1e-12
    assert np.max(np.abs(rho - rho.conj().T)) == 0.0          # exactly Hermitian, like the reference's sum of outers
    # readout transform of the ensemble's diagonal (config 3's last step) keeps a distribution
    out = nm.readout_error.apply_to_distribution(np.real(np.diag(rho)).copy(), 12)
    assert abs(out.sum() - 1.0) < 1e-12 and out.min() >= 0.0


def test_config2_parameter_batch_4096_sets_unit_norm():
    from qsb import capi
    from qsb.workloads import layered_circuit
    from quantum_sim.engine.optimizer import ParameterizedCircuitConfig
    n = 16
    cfg = ParameterizedCircuitConfig.auto_detect(_circuit(n, layered_circuit(n, 64, 2026)))
    vals = np.random.default_rng(2027).uniform(-np.pi, np.pi, (4096, cfg.num_params))
    c, states = cfg.run_batch(vals)
    out = c.alloc(4096 * 16)
    c.overlap(n, states, 0, states, 0, 1, 4096, out)
    norms = out.download(np.complex128, (4096,))
    assert np.max(np.abs(norms - 1.0)) < 1e-12


def test_noisy_16q_trajectories_unit_norm_and_branch_parity():
    from oracle import qsim_oracle as O
    from qsb import capi
    from qsb.lowering import lower_circuit
    from qsb.workloads import layered_circuit, config3_noise
    from quantum_sim.engine.gate_registry import GateRegistry
    n, T = 16, 240
    gates = layered_circuit(n, 64, 2026)
    noise = config3_noise()
    prog, _ = lower_circuit(n, _circuit(n, gates).get_ordered_gates(), GateRegistry.instance(),
                            lambda name: [(k, p, None) for k, p in O.channels_for(noise, name)])
    c = capi.get_context()
    draws = np.random.default_rng(99).random((T, prog.n_draws))
    draws[0, ::97] = 0.9995                                   # force a few amplitude-damping reductions / Pauli hits
    states = c.alloc(T * (16 << n))
    br = c.to_device(np.full((T, prog.n_draws), -1, dtype=np.int32))
    c.run(c.program(prog), T, states=states, uniforms=c.to_device(draws), uniforms_stride=prog.n_draws,
          branches=br, branches_stride=prog.n_draws)
    out = c.alloc(T * 16)
    c.overlap(n, states, 0, states, 0, 1, T, out)
    assert np.max(np.abs(out.download(np.complex128, (T,)) - 1.0)) < 1e-12
    # one full 16-qubit trajectory against the oracle (branches bit-exact, amplitudes 1e-12): ~25 s of NumPy
    psi, _, branches, _ = O.run_state(n, gates, None, noise, draws[0])
    got = states.download(np.complex128, (T, 2 ** n))[0]
    assert br.download(np.int32, (T, prog.n_draws))[0].tolist() == branches
    assert np.max(np.abs(got - psi)) < 1e-12


def test_config3_per_layer_mutual_information_matches_reference():
    """BASELINE config 3: per-trajectory, per-layer mutual information of all 66 pairs on the 12-qubit noisy
    circuit -- trajectories with per-column snapshots, the single-read DMMA RDM kernel and device entropies -- against
    the REAL reference's StateAnalysis.mutual_information (tests/golden/make_golden_cfg3_mi.py: 2 noise seeds x 2
    layers x 66 pairs)."""
    import json
    import os
    from qsb import capi
    from qsb.workloads import layered_circuit
    from quantum_sim.engine.analysis import all_pairs_mutual_information_device, StateAnalysis
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
    from quantum_sim.engine.simulator import Simulator
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cfg3_mi.json")) as f:
        gold = json.load(f)
    n = gold["n"]
    qc = QuantumCircuit(n)
    for name, targets, params, col in layered_circuit(n, gold["depth"], gold["circuit_seed"]):
        qc.add_gate(GateInstance(name, list(targets), list(params), col))
    ctx = capi.get_context()
    for s in gold["noise_seeds"]:
        nm = NoiseModel()
        nm.add_global_noise(DepolarizingNoise(0.01))
        nm.add_global_noise(AmplitudeDampingNoise(0.02))
        nm.set_seed(s)
        res = Simulator(nm).run(qc, shots=0, record_steps=True, seed=s)
        states = np.stack([st.data for st in res.step_states])
        mi = all_pairs_mutual_information_device(n, ctx.to_device(states), 0, len(states))
        for layer in gold["layers"]:
            want = np.array(gold["mi"][f"{s}:{layer}"])
            assert np.max(np.abs(mi[layer] - want)) < 1e-9, (s, layer)
            # and through the per-pair API of the reference
            k = 0
            for i in range(n):
                for j in range(i + 1, n):
                    if k % 13 == 0:
                        assert abs(StateAnalysis.mutual_information(res.step_states[layer], i, j) - want[k]) < 1e-9
                    k += 1
