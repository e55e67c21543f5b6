"""Host compiler of the streamed passes (qsb/stream.py), checked on CPU: the plan -- host-fused sweeps, reorder passes,
exchanges -- is replayed with NumPy (a model of what csrc/qsb_stream.cuh computes per tile) and compared with the
oracle's gate-by-gate `apply_gate`.  The device side of the same plans is tested in tests/test_gpu_stream.py."""

import numpy as np
import pytest

from oracle import qsim_oracle as O
from qsb import stream as S
from qsb.compiler import Lowering
from qsb.workloads import layered_circuit
from quantum_sim.engine.gate_registry import GateRegistry
from test_bigstate import ordered, lower

REG = GateRegistry.instance()


def block_op_matrix(op, cdata, mat):
    """Operator of one block op; op.t[0] = most significant bit of its index."""
    if op.kind == S.B_MAT1:
        return np.asarray(op.U, dtype=np.complex128)
    if op.kind in (S.B_DENSE2, S.B_DENSE3):
        D = 4 if op.kind == S.B_DENSE2 else 8
        return cdata[mat:mat + 2 * D * D].view(np.complex128).reshape(D, D)
    return O.gate_matrix({S.B_CX: "CNOT", S.B_CZ: "CZ", S.B_SWAP: "SWAP", S.B_CCX: "Toffoli", S.B_CSWAP: "Fredkin"}[op.kind])


def replay(steps, n, g, psi, cdata):
    """Execute a stream plan with NumPy on the full 2^n vector viewed as 2^g shards."""
    L = n - g
    shards = [psi[r << L:(r + 1) << L].copy() for r in range(1 << g)]
    def exchange(shards):
        chunks = [s.reshape(1 << g, -1) for s in shards]
        return [np.concatenate([chunks[src][dst] for src in range(1 << g)]) for dst in range(1 << g)]

    for st in steps:
        if st.kind == "exchange":
            shards = exchange(shards)
            continue
        sp = st.spass
        assert sorted(sp.positions) == list(range(L)) and sorted(sp.positions_out) == list(range(L))
        assert sp.positions[:sp.l] == list(range(sp.l)) and sp.positions_out[:sp.l] == list(range(sp.l))
        assert len(sp.sweeps) <= S.MAX_SWEEPS and (st.kind != "reorder" or not sp.blocks)
        new = []
        for s in shards:
            t = s.reshape([2] * L)                                   # numpy axis a <-> position L-1-a
            t = t.transpose([L - 1 - sp.positions[j] for j in reversed(range(L))]).reshape(-1)   # index = slot bits
            assert len(sp.blocks) <= S.MAX_SWEEPS
            for bk in sp.blocks:                                      # what the device runs: register blocks of four slot bits
                assert len(set(bk.b)) == 4 and all(0 <= b < sp.m for b in bk.b) and len(bk.ops) <= S.MAX_BLOCK_OPS
                for op in bk.ops:
                    assert len(set(op.t)) == len(op.t) and all(0 <= x < 4 for x in op.t)
                    if op.kind in (S.B_DENSE2, S.B_DENSE3):
                        assert op.t == [3, 2, 1][:len(op.t)]
                    t = O.apply_textbook(t, L, block_op_matrix(op, cdata, bk.mat), [L - 1 - bk.b[x] for x in op.t])
            src = t.reshape([2] * L)                                  # axis a <-> slot L-1-a
            perm = [0] * L
            for j in range(L):
                perm[L - 1 - sp.positions_out[j]] = L - 1 - j
            new.append(np.ascontiguousarray(src.transpose(perm)).reshape(-1))
        shards = new
        if st.scatter:                      # the exchange rides in this pass's store (qsb_stream_run_scatter)
            assert g > 0 and all(p < L - g for p in sp.positions_out[sp.l:sp.l + sp.e])     # no box dimension on a peer bit
            shards = exchange(shards)
    return np.concatenate(shards)


def reference_state(n, gl, psi):
    ref = psi
    for name, targets, params in gl:
        ref = O.apply_gate(ref, n, O.gate_matrix(name, params), targets)
    return ref


@pytest.mark.parametrize("n,g,m,layout", [(7, 0, 4, "reference"), (8, 1, 4, "reference"), (9, 2, 5, "textbook"),
                                          (11, 3, 5, "reference"), (11, 0, 6, "textbook"), (12, 2, 6, "reference")])
@pytest.mark.parametrize("fuse_store", [True, False])
def test_stream_plan_replay_matches_oracle(n, g, m, layout, fuse_store):
    rng = np.random.default_rng(n)
    gl = ordered(n, layered_circuit(n, 6, 40 + n))
    lw = lower(n, gl, layout=layout)
    steps, pos_of, pending = S.plan(lw.items, lw.pool.array(), n, g, list(range(n)), local_bits=m, low_bits=2, box_bits=1,
                                    fuse_store=fuse_store)
    assert not pending
    kinds = [s.kind for s in steps]
    if g:
        assert ("exchange" in kinds) != fuse_store and any(s.scatter for s in steps) == fuse_store
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    got_mem = replay(steps, n, g, psi, lw.pool.array())
    if layout == "reference":
        ref = reference_state(n, gl, psi)
    else:
        ref = psi
        for name, targets, params in gl:
            ref = O.apply_textbook(ref, n, O.gate_matrix(name, params), targets)
    pos = [pos_of[lw.bit_of_axis[j]] for j in range(n)]                 # memory position of reference axis j
    got = np.ascontiguousarray(got_mem.reshape([2] * n).transpose([n - 1 - pos[j] for j in range(n)])).reshape(-1)
    assert np.max(np.abs(got - ref)) < 1e-12


def test_one_qubit_ops_cost_no_sweep_and_no_residency():
    """Only multi-qubit gates (plus the final flush) become sweeps; a circuit of 1-qubit gates is one flush pass."""
    n = 10
    gl = ordered(n, layered_circuit(n, 8, 3))
    lw = lower(n, gl, layout="textbook")
    steps, _, _ = S.plan(lw.items, lw.pool.array(), n, 0, list(range(n)), local_bits=6, low_bits=2, box_bits=1)
    n_multi = sum(1 for name, t, _ in gl if len(t) > 1)
    sweeps = [sw for st in steps for sw in st.spass.sweeps]
    assert sum(1 for sw in sweeps if sw.gate != S.G_NONE) == n_multi
    assert sum(1 for sw in sweeps if sw.gate == S.G_NONE) <= (n + 2) // 3 + len(steps)
    # ... and gates whose qubits fit four index bits share one shared-memory round trip
    assert sum(len(st.spass.blocks) for st in steps) < len(sweeps)
    lw1 = Lowering(n, layout="textbook")
    for q in range(n):
        lw1.gate("H", [q], [], REG.get("H").matrix_func)
        lw1.gate("Rz", [q], [0.3 * q], REG.get("Rz").matrix_func)
    steps1, _, _ = S.plan(lw1.items, lw1.pool.array(), n, 0, list(range(n)), local_bits=6, low_bits=2, box_bits=1)
    assert sum(len(st.spass.sweeps) for st in steps1) == (n + 2) // 3
    assert sum(len(st.spass.blocks) for st in steps1) == (n + 3) // 4
    psi = np.zeros(2 ** n, dtype=np.complex128)
    psi[0] = 1.0
    got = replay(steps1, n, 0, psi, lw1.pool.array())
    ref = psi
    for q in range(n):
        ref = O.apply_textbook(ref, n, O.gate_matrix("H", []), [q])
        ref = O.apply_textbook(ref, n, O.gate_matrix("Rz", [0.3 * q]), [q])
    assert np.max(np.abs(got - ref)) < 1e-12


def test_pauli_draws_parameters_and_dense_gates_in_a_stream_plan():
    """Pauli Kraus branches are decided by the host compiler from the reference's uniforms (or the Philox stream),
    parameterised rotations read the parameter row, dense 2- and 3-qubit matrices ride along as QSB_G_DENSE."""
    n = 9
    rng = np.random.default_rng(5)
    lw = Lowering(n, layout="reference")
    u2 = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    u3 = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]
    prm = rng.uniform(-3, 3, 6)
    script = []
    for q in range(n):
        lw.gate("H", [q], [], REG.get("H").matrix_func)
        lw.kraus("depolarizing", 0.6, q)
        script.append(("H", [q]))
        script.append(("depol", q))
    lw.matrix(u2, [7, 2]); script.append((u2, [7, 2]))
    lw.param_gate("Rx", [3], 0); script.append((O.gate_matrix("Rx", [prm[0]]), [3]))
    lw.param_gate("U3", [8], 1); script.append((O.gate_matrix("U3", list(prm[1:4])), [8]))
    lw.matrix(u3, [0, 8, 4]); script.append((u3, [0, 8, 4]))
    lw.param_gate("Rz", [4], 4); script.append((O.gate_matrix("Rz", [prm[4]]), [4]))
    lw.param_gate("Phase", [1], 5); script.append((O.gate_matrix("Phase", [prm[5]]), [1]))
    lw.gate("CNOT", [1, 5], [], None); script.append(("CNOT", [1, 5]))
    uni = rng.random(lw.n_draws)
    steps, pos_of, _ = S.plan(lw.items, lw.pool.array(), n, 1, list(range(n)), local_bits=5, low_bits=2, box_bits=1,
                              params=prm, uniforms=uni)
    assert any(st.scatter for st in steps)
    psi = np.zeros(2 ** n, dtype=np.complex128)
    psi[0] = 1.0
    got_mem = replay(steps, n, 1, psi, lw.pool.array())
    ref, d = psi, 0
    cdf = [0.4, 0.6, 0.8]                                      # choice() thresholds of depolarizing(0.6)
    for what, t in script:
        if isinstance(what, str) and what == "depol":
            idx = int(np.sum(np.array(cdf) <= uni[d]))
            d += 1
            # noise.py:241-255: the chosen branch went through apply_gate, so even the identity branch leaves the axis scramble
            ref = O.apply_gate(ref, n, O.gate_matrix("IXYZ"[idx]), [t])
        elif isinstance(what, str):
            ref = O.apply_gate(ref, n, O.gate_matrix(what), t)
        else:
            ref = O.apply_gate(ref, n, what, t)
    pos = [pos_of[lw.bit_of_axis[j]] for j in range(n)]
    got = np.ascontiguousarray(got_mem.reshape([2] * n).transpose([n - 1 - pos[j] for j in range(n)])).reshape(-1)
    assert np.max(np.abs(got - ref)) < 1e-12
    # the Philox stream of the host compiler is the kernels' (known answers from qsb_philox_uniform, test_executor)
    assert 0.0 <= S.philox_uniform(123, 7, 5) < 1.0


def test_geometry_defaults():
    assert S.choose_geometry(26) == (12, 5, 3)           # 64 KiB tiles, 512-byte rows, 4 KiB per TMA op, 16 ops per tile
    assert S.choose_geometry(27, low_bits=4) == (12, 4, 3)
    m, l, e = S.choose_geometry(20, local_bits=10)
    assert (m, l) == (10, 5) and m - l - e <= 5
