"""Executor + host compiler vs the oracle and the reference's golden vectors.

Every case runs twice: `emu` = the test-only host emulator (CPU; same `qsb_exec.cuh` op loop, same
`qsb_op` programs, OS threads instead of CUDA threads) and `gpu` = the real CUDA path through the
C ABI of libqsb.so (marked gpu)."""

import numpy as np
import pytest

from oracle import qsim_oracle as O
from conftest import as_gates, as_noise
from emu_util import emu_run
from qsb.lowering import lower_circuit
from qsb.compiler import Lowering
from qsb.workloads import layered_circuit, config3_noise
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.gate_registry import GateRegistry

TOL = 1e-12     # complex128 amplitude tolerance of BASELINE.json's north_star
REG = GateRegistry.instance()


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def run(request):
    if request.param == "emu":
        return emu_run
    from gpu_util import gpu_run
    return gpu_run


def make_circuit(n, gates, initial=None):
    qc = QuantumCircuit(n, initial_states=list(initial) if initial else [])
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    return qc


def channels_of_factory(noise):
    if noise is None:
        return None
    def f(name):
        return [(k, p, None) for k, p in O.channels_for(noise, name)]
    return f


def basis_index(initial):
    n = len(initial)
    return sum(1 << (n - 1 - i) for i, b in enumerate(initial) if b)


@pytest.mark.parametrize("gbits", [0, 1, 2, 3])
def test_random_circuits_vs_oracle(run, golden, gbits):
    j, a = golden
    for rec in j["random_circuits"]:
        n = rec["n"]
        m = n - gbits
        if m < 3:
            continue
        qc = make_circuit(n, as_gates(rec["gates"]), rec["initial"])
        prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, record_steps=True, local_bits=m)
        out = run(prog, T=[1, 2, 3][n % 3], default_basis=basis_index(rec["initial"]))
        assert np.max(np.abs(out["states"][0] - a[rec["tag"]])) < TOL, rec["tag"]
        assert np.max(np.abs(out["snapshots"][0] - a[rec["tag"] + "_steps"])) < TOL, rec["tag"]
        if gbits:
            assert prog.n_remaps >= 0


def test_sigma_dense_ops(run, golden):
    j, a = golden
    n = j["sigma"]["n"]
    for i, t in enumerate(j["sigma"]["targets"]):
        k = len(t)
        u = a["sigma_mat"][i][:4 ** k].reshape(2 ** k, 2 ** k)
        for m in (n, n - 2):
            lw = Lowering(n)
            lw.matrix(u, t)
            out = run(lw.finish(m), T=2, states=a["sigma_in"][i][None])
            assert np.max(np.abs(out["states"][0] - a["sigma_out"][i])) < TOL, (t, m)


def test_kron_string_observable(run, golden):
    j, a = golden
    rng = np.random.default_rng(3)
    n = 7
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    for label, targets in (("XYZX", [5, 0, 3, 1]), ("ZZYXI", [6, 2, 0, 4, 1])):
        obs = np.array([[1]], dtype=complex)
        for ch in label:
            obs = np.kron(obs, O.gate_matrix(ch))
        lw = Lowering(n)
        lw.matrix(obs, targets)
        out = run(lw.finish(5), T=2, states=psi[None])
        assert np.max(np.abs(out["states"][0] - O.apply_gate(psi, n, obs, targets))) < TOL
    with pytest.raises(NotImplementedError):
        lw = Lowering(n)
        lw.matrix(a["bigk_0_mat"], [3, 0, 5, 1])


@pytest.mark.parametrize("gbits", [0, 2])
def test_noisy_trajectories_reference_draws(run, golden, gbits):
    j, a = golden
    for rec in j["noisy"]:
        n = rec["n"]
        m = n - gbits
        if m < 3:
            continue
        g, noise = as_gates(rec["gates"]), as_noise(rec["noise"])
        qc = make_circuit(n, g)
        prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, channels_of_factory(noise),
                                record_steps=True, local_bits=m)
        assert prog.n_draws == O.draw_count(n, g, noise)
        draws = np.random.default_rng(rec["noise_seed"]).random(max(prog.n_draws, 1))
        out = run(prog, T=2, uniforms=draws[None], want_branches=True)
        _, _, branches, _ = O.run_state(n, g, None, noise, draws)
        assert out["branches"][0][:len(branches)].tolist() == branches, rec["tag"]
        assert np.max(np.abs(out["states"][0] - a[rec["tag"]])) < TOL, rec["tag"]
        assert np.max(np.abs(out["snapshots"][0] - a[rec["tag"] + "_steps"])) < TOL, rec["tag"]


@pytest.mark.parametrize("gbits", [0, 1, 3])
@pytest.mark.parametrize("gamma", [0.02, 0.3])
def test_amplitude_damping_draws_outside_the_certain_range(run, gbits, gamma):
    """Most draws are forced into [1 - gamma, 1), where the branch depends on P(q = 1) of the state: the executor
    first brackets it with diagonal weights (no flush; csrc/qsb_exec.cuh, QSB_AD_BOUNDS) and falls back to the exact
    marginal; either way branches and amplitudes are the reference's (noise.py:241-255)."""
    n = 7
    gates = layered_circuit(n, 6, 23)
    noise = {"global": [("depolarizing", 0.01), ("amplitude_damping", gamma)], "gate": {}}
    qc = make_circuit(n, gates)
    prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, channels_of_factory(noise), local_bits=n - gbits)
    rng = np.random.default_rng(100 + gbits)
    draws = rng.random((8, prog.n_draws))
    slow = rng.random(draws.shape) < 0.7
    draws = np.where(slow, 1.0 - gamma * rng.random(draws.shape), draws)
    out = run(prog, count=8, T=2, uniforms=draws, want_branches=True)
    n_k1 = 0
    for t in range(8):
        psi, _, br, _ = O.run_state(n, gates, None, noise, draws[t])
        assert out["branches"][t].tolist() == br
        assert np.max(np.abs(out["states"][t] - psi)) < TOL
        n_k1 += sum(1 for k, x in enumerate(br) if k % 2 == 1 and x == 1)
    assert n_k1 > 0                                  # K1 branches (rank-1 pending matrices) were exercised


def test_generic_kraus_matches_builtin_channels(run):
    rng = np.random.default_rng(11)
    n = 5
    gates = layered_circuit(n, 4, 3)
    noise = {"global": [("amplitude_damping", 0.3), ("depolarizing", 0.2)], "gate": {}}
    qc = make_circuit(n, gates)

    def generic(name):
        return [("generic", p, O.kraus_ops(k, p)) for k, p in O.channels_for(noise, name)]

    prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, generic, local_bits=4)
    draws = rng.random((6, prog.n_draws))
    out = run(prog, count=6, T=2, uniforms=draws, want_branches=True)
    for t in range(6):
        psi, _, br, _ = O.run_state(n, gates, None, noise, draws[t])
        assert out["branches"][t].tolist() == br
        assert np.max(np.abs(out["states"][t] - psi)) < TOL


def test_parameter_batch(run, golden):
    n = 6
    gates = layered_circuit(n, 5, 21)
    qc = make_circuit(n, gates)
    slots = O.param_slots(gates)
    offs, off = {}, 0
    for gi, g in enumerate(qc.gates):
        k = O.NUM_PARAMS.get(g.gate_name, 0)
        if k:
            offs[id(g)] = off
            off += k
    assert off == len(slots)
    prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, param_offsets=offs, local_bits=4)
    vals = np.random.default_rng(5).uniform(-np.pi, np.pi, (5, off))
    out = run(prog, count=5, T=3, params=vals)
    for t in range(5):
        psi = O.run_state(n, O.bind_values(gates, vals[t]))[0]
        assert np.max(np.abs(out["states"][t] - psi)) < TOL


def test_philox_mode_matches_oracle_uniforms(run):
    n = 5
    gates = layered_circuit(n, 3, 8)
    noise = {"global": [("depolarizing", 0.3), ("amplitude_damping", 0.4)], "gate": {}}
    qc = make_circuit(n, gates)
    prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, channels_of_factory(noise), local_bits=3)
    out = run(prog, count=4, T=2, seed=0xDEADBEEF12345, traj_offset=100, want_branches=True,
                  accum_probs=True)
    acc = np.zeros(2 ** n)
    for t in range(4):
        u = O.philox_uniforms(0xDEADBEEF12345, 100 + t, prog.n_draws)
        psi, _, br, _ = O.run_state(n, gates, None, noise, u)
        assert out["branches"][t].tolist() == br
        assert np.max(np.abs(out["states"][t] - psi)) < TOL
        acc += np.abs(psi) ** 2
    assert np.max(np.abs(out["probs"] - acc)) < 1e-12


def test_config3_trajectory_12q(run, golden):
    j, a = golden
    g, noise = layered_circuit(12, 16, 2026), config3_noise()
    qc = make_circuit(12, g)
    for m in (12, 10):
        prog, _ = lower_circuit(12, qc.get_ordered_gates(), REG, channels_of_factory(noise), local_bits=m)
        seed = j["cfg3_traj_seeds"][0]
        draws = np.random.default_rng(seed).random(prog.n_draws)
        out = run(prog, T=2, uniforms=draws[None])
        assert np.max(np.abs(out["states"][0] - a[f"cfg3_traj_{seed}"])) < TOL


def test_layered16_cluster8(run, golden):
    j, a = golden
    g = layered_circuit(16, 64, 2026)
    qc = make_circuit(16, g)
    prog, _ = lower_circuit(16, qc.get_ordered_gates(), REG)
    assert (prog.n, prog.m) == (16, 13)
    out = run(prog, T=1)
    psi = out["states"][0]
    assert np.max(np.abs(psi[a["layered16_idx"]] - a["layered16_amps"])) < TOL
    assert int(np.argmax(np.abs(psi))) == j["layered16"]["argmax"]


def run_stream(run, plan, psi0):
    """Run a StreamPlan pass by pass (every pass is one launch over all tiles of the state)."""
    cur = np.ascontiguousarray(psi0, dtype=np.complex128)[None]
    for k, prog in enumerate(plan.passes):
        last = k == len(plan.passes) - 1
        cur = run(prog, T=2, states=cur, out_of_place=last and plan.final_out_of_place)["states"]
    return cur[0]


@pytest.mark.parametrize("n,m", [(8, 4), (9, 4), (10, 5), (11, 5)])
def test_streaming_passes_vs_oracle(run, n, m):
    """n - m > 3: the state stays in memory and is streamed tile by tile, one pass per launch."""
    rng = np.random.default_rng(n * 31 + m)
    gates = layered_circuit(n, 6, 100 + n)
    qc = QuantumCircuit(n)
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    plan, _ = lower_circuit(n, qc.get_ordered_gates(), REG, local_bits=m, stream=True)
    assert len(plan.passes) >= 2
    psi0 = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi0 /= np.linalg.norm(psi0)
    ordered = [g for col in qc.get_ordered_gates() for g in col]     # the reference's execution order
    ref = psi0
    for g in ordered:
        ref = O.apply_gate(ref, n, O.gate_matrix(g.gate_name, g.params), list(g.target_qubits))
    got = run_stream(run, plan, psi0)
    assert np.max(np.abs(got - ref)) < TOL
    # textbook layout: no final reorder, everything in place
    plan_t, _ = lower_circuit(n, qc.get_ordered_gates(), REG, local_bits=m, stream=True, layout="textbook")
    assert not plan_t.final_out_of_place
    got_t = run_stream(run, plan_t, psi0)
    ref_t = psi0
    for g in ordered:
        ref_t = O.apply_textbook(ref_t, n, O.gate_matrix(g.gate_name, g.params), list(g.target_qubits))
    assert np.max(np.abs(got_t - ref_t)) < TOL


# ---- complex64 mode (BASELINE: reported separately, tolerance 1e-5) ---------------------------------------------
TOL64 = 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("gbits", [0, 1, 3])
def test_c64_random_circuits_and_snapshots(golden, gbits):
    from gpu_util import gpu_run
    j, a = golden
    for rec in j["random_circuits"]:
        n = rec["n"]
        m = n - gbits
        if m < 3:
            continue
        qc = make_circuit(n, as_gates(rec["gates"]), rec["initial"])
        prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, record_steps=True, local_bits=m)
        out = gpu_run(prog, default_basis=basis_index(rec["initial"]), precision="c64")
        assert np.max(np.abs(out["states"][0] - a[rec["tag"]])) < TOL64, rec["tag"]
        assert np.max(np.abs(out["snapshots"][0] - a[rec["tag"] + "_steps"])) < TOL64, rec["tag"]


@pytest.mark.gpu
def test_c64_noisy_16q_cluster4_and_reductions():
    """16 qubits in complex64: 2^14 amplitudes per CTA, clusters of 4; Philox-free reference draws.  Branch
    decisions can differ from complex128 only when a uniform lands within fp32 rounding of a threshold."""
    from gpu_util import gpu_run
    from qsb import capi
    n = 16
    gates = layered_circuit(n, 8, 2026)
    noise = config3_noise()
    qc = make_circuit(n, gates)
    prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, channels_of_factory(noise), local_bits=14, max_local_bits=14)
    assert (prog.n, prog.m) == (16, 14)
    draws = np.random.default_rng(5).random((3, prog.n_draws))
    out = gpu_run(prog, count=3, uniforms=draws, want_branches=True, precision="c64")
    for t in range(3):
        psi, _, br, _ = O.run_state(n, gates, None, noise, draws[t])
        assert out["branches"][t].tolist() == br
        assert np.max(np.abs(out["states"][t] - psi)) < TOL64
    # reductions on complex64 buffers accumulate in double
    ctx = capi.get_context(precision="c64")          # complex64 users have a context of their own
    psi = out["states"].astype(np.complex64)
    buf = ctx.to_device(psi)
    pr = ctx.alloc(3 * 2 ** n * 8)
    ctx.probabilities(n, buf, 0, 3, pr)
    assert np.max(np.abs(pr.download(np.float64, (3, 2 ** n)) - np.abs(psi.astype(np.complex128)) ** 2)) < 1e-12
    ov = ctx.alloc(3 * 16)
    ctx.overlap(n, buf, 0, buf, 0, 1, 3, ov)
    assert np.max(np.abs(ov.download(np.complex128, (3,)) - np.sum(np.abs(psi.astype(np.complex128)) ** 2, axis=1))) < 1e-12
    assert capi.get_context().precision == "c128"


# ---- REMAP ops that exchange several (rank bit, local bit) pairs in one pass ------------------------------------
@pytest.mark.parametrize("gbits", [2, 3])
def test_multi_pair_remaps(run, gbits):
    """Layered noisy circuit on clusters of 4 / 8: the planner packs up to three pair swaps into one REMAP op; the
    result equals the oracle and the single-pair plan (same draws, same branches)."""
    from qsb.compiler import REMAP
    n = 9
    m = n - gbits
    gates = layered_circuit(n, 12, 77)
    noise = config3_noise()
    qc = make_circuit(n, gates)
    lw_args = (n, qc.get_ordered_gates(), REG, channels_of_factory(noise))
    prog, _ = lower_circuit(*lw_args, local_bits=m)
    k = [int(o["b2"]) + 1 for o in prog.ops if int(o["kind"]) == REMAP]
    assert k and max(k) >= 2, k                     # at least one exchange moves two or three pairs at once
    if gbits == 3:
        assert max(k) == 3, k
    for o in prog.ops:
        if int(o["kind"]) == REMAP and int(o["b2"]) >= 1:
            aux = int(o["aux"])
            g = [int(o["b0"]), aux & 255, (aux >> 16) & 255][:int(o["b2"]) + 1]
            l = [int(o["b1"]), (aux >> 8) & 255, (aux >> 24) & 255][:int(o["b2"]) + 1]
            assert len(set(g)) == len(g) and len(set(l)) == len(l) and max(g) < gbits and max(l) < m
    draws = np.random.default_rng(3).random((2, prog.n_draws))
    out = run(prog, count=2, T=2, uniforms=draws, want_branches=True)
    for t in range(2):
        psi, _, br, _ = O.run_state(n, gates, None, noise, draws[t])
        assert out["branches"][t][:len(br)].tolist() == br
        assert np.max(np.abs(out["states"][t] - psi)) < TOL


# ---- every specialised sweep variant: (gate, class of the pending matrix of each of its bits) ---------------------
@pytest.mark.parametrize("gate,k", [("CNOT", 2), ("CZ", 2), ("SWAP", 2), ("Toffoli", 3), ("Fredkin", 3)])
@pytest.mark.parametrize("gbits", [0, 2])
def test_every_dense_mask_and_gate_of_the_specialised_sweeps(run, gate, k, gbits):
    """csrc/qsb_exec.cuh instantiates the sweep per mask of bits that carry a full 2x2; the other bits carry a real scale
    (a damping K0) or nothing, and permutation gates only move the store.  All 3^k class combinations of every gate:
    nothing pending / K0 pending (identity gate with certain-K0 amplitude damping) / Ry pending, on an entangled state."""
    import itertools
    n = 7
    targets = [5, 1, 3][:k]
    noise = {"global": [], "gate": {"I": [("amplitude_damping", 0.2)]}}
    for classes in itertools.product((0, 1, 2), repeat=k):
        gates, col = [], 0
        for q in range(n):
            gates.append(("Ry", [q], [0.3 + 0.2 * q], col))
        col += 1
        for q in range(0, n - 1, 2):
            gates.append(("CNOT", [q, q + 1], [], col))
        col += 1
        for q in range(1, n - 1, 2):
            gates.append(("CZ", [q, q + 1], [], col))
        gates.append(("CNOT", [0, n - 1], [], col + 1))          # every qubit met a 2-qubit gate: nothing is pending
        col += 2
        for q, c in zip(targets, classes):
            if c == 1:
                gates.append(("I", [q], [], col))
            elif c == 2:
                gates.append(("Ry", [q], [0.9 - 0.1 * q], col))
        gates.append((gate, targets, [], col + 1))
        qc = make_circuit(n, gates)
        prog, _ = lower_circuit(n, qc.get_ordered_gates(), REG, channels_of_factory(noise), local_bits=n - gbits)
        draws = np.full((1, max(prog.n_draws, 1)), 0.5)           # below 1 - gamma: K0 whatever the state
        out = run(prog, count=1, T=2, uniforms=draws)
        psi, _, _, _ = O.run_state(n, gates, None, noise, draws[0])
        assert np.max(np.abs(out["states"][0] - psi)) < TOL, (gate, classes)
