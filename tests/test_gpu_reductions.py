"""Reduction kernels of libqsb.so (probabilities, sampling, overlaps, parities, RDMs, rho, readout)
against the oracle, through the C ABI.  GPU only."""

import numpy as np
import pytest

from oracle import qsim_oracle as O

pytestmark = pytest.mark.gpu


def rand_states(rng, count, n):
    v = rng.normal(size=(count, 2 ** n)) + 1j * rng.normal(size=(count, 2 ** n))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


@pytest.fixture(scope="module")
def ctx():
    from qsb import capi
    return capi.get_context()


@pytest.mark.parametrize("n", [1, 3, 8, 13, 16])
def test_probabilities_and_sum(ctx, n):
    rng = np.random.default_rng(n)
    psi = rand_states(rng, 3, n)
    s = ctx.to_device(psi)
    out = ctx.alloc(3 * 2 ** n * 8)
    ctx.probabilities(n, s, 0, 3, out)
    p = out.download(np.float64, (3, 2 ** n))
    assert np.max(np.abs(p - np.abs(psi) ** 2)) < 1e-15
    acc = ctx.alloc(2 ** n * 8).zero()
    ctx.probabilities_sum(n, s, 1, 2, acc)
    assert np.max(np.abs(acc.download(np.float64, (2 ** n,)) - (np.abs(psi[1:]) ** 2).sum(0))) < 1e-14


@pytest.mark.parametrize("n", [2, 5, 12, 16])
def test_sample_index_matches_numpy_choice(ctx, n):
    rng = np.random.default_rng(100 + n)
    count = 64
    psi = rand_states(rng, count, n)
    psi[0, : 2 ** n // 2] = 0          # leading zero-probability block
    psi[0] /= np.linalg.norm(psi[0])
    psi[1] = 0
    psi[1, 2 ** n - 1] = 1             # all mass on the last index
    u = rng.random(count)
    u[2] = 0.0
    s, ub = ctx.to_device(psi), ctx.to_device(u)
    out = ctx.alloc(count * 8)
    ctx.sample_index(n, s, 0, count, ub, out)
    got = out.download(np.int64, (count,))
    want = [O.measure_all_index(psi[t], u[t]) for t in range(count)]
    assert got.tolist() == want


@pytest.mark.parametrize("n", [1, 4, 13, 16])
def test_overlap(ctx, n):
    rng = np.random.default_rng(200 + n)
    a, b = rand_states(rng, 5, n), rand_states(rng, 5, n)
    da, db = ctx.to_device(a), ctx.to_device(b)
    out = ctx.alloc(5 * 16)
    ctx.overlap(n, da, 0, db, 0, 1, 5, out)
    got = out.download(np.complex128, (5,))
    assert np.max(np.abs(got - np.array([np.vdot(a[t], b[t]) for t in range(5)]))) < 1e-13
    ctx.overlap(n, da, 1, db, 2, 0, 3, out)           # broadcast one b
    got = out.download(np.complex128, (5,))[:3]
    assert np.max(np.abs(got - np.array([np.vdot(a[t], b[2]) for t in (1, 2, 3)]))) < 1e-13


def test_masked_parity_steane_checks(ctx):
    n = 13
    rng = np.random.default_rng(300)
    psi = rand_states(rng, 4, n)
    checks = O.STEANE_CHECKS + [list(range(7))]
    masks = [sum(1 << (n - 1 - q) for q in c) for c in checks]
    s = ctx.to_device(psi)
    out = ctx.alloc(4 * len(masks) * 16)
    ctx.masked_parity(n, s, 0, 4, masks, out)
    got = out.download(np.float64, (4, len(masks), 2))
    for t in range(4):
        for k, c in enumerate(checks):
            e, o = O.z_parity_weights(psi[t], n, c)
            assert abs(got[t, k, 0] - e) < 1e-13 and abs(got[t, k, 1] - o) < 1e-13


@pytest.mark.parametrize("n", [2, 4, 5, 8, 12, 13, 14])
def test_rdm_all_pairs(ctx, n):
    """n = 4..13: one read of the state, all pairs as 8x8 real Grams on DMMA (qsb_rdm_gram_kernel); else the per-pair kernels."""
    rng = np.random.default_rng(400 + n)
    psi = rand_states(rng, 2, n)
    npairs = n * (n - 1) // 2
    s = ctx.to_device(psi)
    r1, r2 = ctx.alloc(2 * n * 4 * 16), ctx.alloc(2 * npairs * 16 * 16)
    ctx.rdm_all(n, s, 0, 2, r1, r2)
    g1 = r1.download(np.complex128, (2, n, 2, 2))
    g2 = r2.download(np.complex128, (2, npairs, 4, 4))
    for t in range(2):
        for q in range(n):
            assert np.max(np.abs(g1[t, q] - O.reduced_density_matrix_1q(psi[t], n, q))) < 1e-13
        k = 0
        for i in range(n):
            for jj in range(i + 1, n):
                assert np.max(np.abs(g2[t, k] - O.partial_trace(psi[t], n, [i, jj]))) < 1e-13
                k += 1


def test_rdm_gram_kernel_batches_and_single_outputs(ctx):
    """More states than CTAs (grid-stride), 1-qubit RDMs alone (n - 1 pairs carry them), 2-qubit RDMs alone, the scalar
    per-pair kernels as a second implementation, and complex64 input."""
    import os
    from qsb import capi
    n, count = 9, 700
    npairs = n * (n - 1) // 2
    rng = np.random.default_rng(77)
    psi = rand_states(rng, count, n)
    s = ctx.to_device(psi)
    r1, r2 = ctx.alloc(count * n * 4 * 16), ctx.alloc(count * npairs * 16 * 16)
    ctx.rdm_all(n, s, 0, count, r1, r2)
    g1, g2 = r1.download(np.complex128, (count, n, 2, 2)), r2.download(np.complex128, (count, npairs, 4, 4))
    for t in (0, 1, 147, 148, 443, 699):
        for q in range(n):
            assert np.max(np.abs(g1[t, q] - O.reduced_density_matrix_1q(psi[t], n, q))) < 1e-13
        k = 0
        for i in range(n):
            for jj in range(i + 1, n):
                assert np.max(np.abs(g2[t, k] - O.partial_trace(psi[t], n, [i, jj]))) < 1e-13
                k += 1
    assert np.max(np.abs(g2 - np.conj(np.swapaxes(g2, -1, -2)))) < 1e-15          # Hermitian to rounding
    only1 = ctx.alloc(count * n * 4 * 16).zero()
    ctx.rdm_all(n, s, 0, count, only1, None)
    assert np.array_equal(only1.download(np.complex128, (count, n, 2, 2)), g1)
    only2 = ctx.alloc(count * npairs * 16 * 16).zero()
    ctx.rdm_all(n, s, 0, count, None, only2)
    assert np.array_equal(only2.download(np.complex128, (count, npairs, 4, 4)), g2)
    os.environ["QSB_RDM_SCALAR"] = "1"                    # round 1's kernels: one CTA per (state, pair)
    try:
        a1, a2 = ctx.alloc(count * n * 4 * 16), ctx.alloc(count * npairs * 16 * 16)
        ctx.rdm_all(n, s, 0, count, a1, a2)
    finally:
        del os.environ["QSB_RDM_SCALAR"]
    assert np.max(np.abs(a1.download(np.complex128, (count, n, 2, 2)) - g1)) < 1e-14
    assert np.max(np.abs(a2.download(np.complex128, (count, npairs, 4, 4)) - g2)) < 1e-14
    c64 = capi.get_context(precision="c64")
    p64 = psi[:5].astype(np.complex64)
    b1, b2 = c64.alloc(5 * n * 4 * 16), c64.alloc(5 * npairs * 16 * 16)
    c64.rdm_all(n, c64.to_device(p64), 0, 5, b1, b2)
    w = p64.astype(np.complex128)
    got2 = b2.download(np.complex128, (5, npairs, 4, 4))
    assert np.max(np.abs(got2[3, 0] - O.partial_trace(w[3], n, [0, 1]))) < 1e-13


@pytest.mark.parametrize("n,count", [(3, 7), (6, 33), (8, 100)])
def test_rho_accumulate(ctx, n, count):
    rng = np.random.default_rng(500 + n)
    psi = rand_states(rng, count, n)
    s = ctx.to_device(psi)
    rho = ctx.alloc(4 ** n * 16).zero()
    ctx.rho_accumulate(n, s, 0, count, 1.0 / count, rho)
    got = rho.download(np.complex128, (2 ** n, 2 ** n))
    want = (psi.T @ psi.conj()) / count
    assert np.max(np.abs(got - want)) < 1e-14


@pytest.mark.parametrize("n", [1, 3, 8, 12, 14, 15, 16])
def test_readout_transform(ctx, n):
    """n <= 14: all axes + renormalisation in one shared-memory kernel; above: one launch per axis."""
    rng = np.random.default_rng(600 + n)
    p = rng.random((2, 2 ** n))
    p /= p.sum(1, keepdims=True)
    buf = ctx.to_device(p)
    ctx.readout_transform(n, buf, 2, 0.03, 0.11)
    got = buf.download(np.float64, (2, 2 ** n))
    for t in range(2):
        assert np.max(np.abs(got[t] - O.readout_distribution(p[t], n, 0.03, 0.11))) < 1e-15


def test_readout_golden(ctx, golden):
    j, a = golden
    buf = ctx.to_device(a["readout8_in"])
    ctx.readout_transform(8, buf, 1, 0.03, 0.11)
    assert np.max(np.abs(buf.download(np.float64, (256,)) - a["readout8_out"])) < 1e-15


def test_argument_errors(ctx):
    from qsb.compiler import Lowering
    small = ctx.alloc(64)
    with pytest.raises(ValueError):
        ctx.probabilities(8, small, 0, 1, small)
    with pytest.raises(ValueError):
        ctx.readout_transform(2, small, 1, 1.5, 0.0)
    lw = Lowering(3)
    lw.gate("X", [0])
    prog = lw.finish()
    dp = ctx.program(prog)
    with pytest.raises(ValueError):
        ctx.run(dp, 4, states=small)           # states buffer too small


@pytest.mark.gpu
def test_mutual_information_all_pairs_on_device(golden):
    """qsb_mi_all_pairs (device RDMs + Jacobi eigenvalues + entropies) against the host eigvalsh path and the
    reference's own mutual_information values frozen in the golden file."""
    from qsb import capi
    from quantum_sim.engine.analysis import all_pairs_mutual_information, all_pairs_mutual_information_device
    from quantum_sim.engine.state_vector import StateVector
    j, a = golden
    ctx = capi.get_context()
    rng = np.random.default_rng(12)
    for n in (2, 3, 6, 10, 12):
        count = 5
        psi = rng.normal(size=(count, 2 ** n)) + 1j * rng.normal(size=(count, 2 ** n))
        psi[0] = 0.0
        psi[0, 0] = psi[0, -1] = 1.0                      # GHZ: every pair has I = 1 bit, rank-deficient RDMs
        psi[1] = 0.0
        psi[1, 5 % 2 ** n] = 1.0                          # product state: I = 0, zero eigenvalues
        psi /= np.linalg.norm(psi, axis=1, keepdims=True)
        buf = ctx.to_device(psi)
        got = all_pairs_mutual_information_device(n, buf, 0, count)
        for t in range(count):
            sv = StateVector._from_host(n, psi[t].copy())
            want = all_pairs_mutual_information(sv)
            assert np.max(np.abs(got[t] - want)) < 1e-10, (n, t)
        assert np.max(np.abs(got[0] - (2.0 if n == 2 else 1.0))) < 1e-12 and np.max(np.abs(got[1])) < 1e-12
    # the reference's numbers: per-layer MI of a GHZ-4 run with record_steps (tests/golden/make_golden.py)
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.simulator import Simulator
    from qsb.workloads import ghz
    qc = QuantumCircuit(4)
    for g in ghz(4):
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    res = Simulator().run(qc, shots=0, record_steps=True)
    states = np.stack([s.data for s in res.step_states])
    got = all_pairs_mutual_information_device(4, ctx.to_device(states), 0, len(states))
    assert np.max(np.abs(got - np.array(j["ghz4_layer_mi"]))) < 1e-10


@pytest.mark.gpu
def test_partial_trace_general_subsystems():
    """StateAnalysis.partial_trace / entanglement_entropy for 3..6 kept qubits against the oracle's O(2^n) restatement."""
    from oracle import qsim_oracle as O
    from quantum_sim.engine.analysis import StateAnalysis
    from quantum_sim.engine.state_vector import StateVector
    rng = np.random.default_rng(21)
    for n, keep in ((5, [0, 2, 4]), (7, [1, 2, 3, 6]), (9, [0, 3, 4, 7, 8]), (10, [9, 1, 5, 2, 7, 0])):
        psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        psi /= np.linalg.norm(psi)
        sv = StateVector._from_host(n, psi.copy())
        got = StateAnalysis.partial_trace(sv, keep)
        t = psi.reshape([2] * n)
        ks = sorted(keep)
        m = np.moveaxis(t, ks, list(range(len(ks)))).reshape(2 ** len(ks), -1)
        want = m @ m.conj().T
        assert np.max(np.abs(got - want)) < 1e-13
        assert np.max(np.abs(got - got.conj().T)) == 0.0
        s = StateAnalysis.entanglement_entropy(sv, keep)
        w = np.linalg.eigvalsh(want)
        w = w[w > 1e-15]
        assert abs(s - float(-np.sum(w * np.log2(w)))) < 1e-10
