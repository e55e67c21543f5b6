"""The TMA tile pipeline for streamed states (csrc/qsb_stream.cuh behind qsb_stream_*; BASELINE config 5) on the B200:
against the oracle's gate-by-gate apply_gate at sizes the oracle finishes in seconds, against round 1's executor path,
and through size-independent properties at 26 qubits."""

import numpy as np
import pytest

from oracle import qsim_oracle as O
from qsb import stream as S
from qsb.compiler import Lowering
from qsb.workloads import layered_circuit
from quantum_sim.engine.gate_registry import GateRegistry
from test_bigstate import ordered, lower
from test_stream_plan import replay

pytestmark = pytest.mark.gpu
REG = GateRegistry.instance()


def oracle_state(n, gl, psi=None):
    if psi is None:
        psi = np.zeros(2 ** n, dtype=np.complex128)
        psi[0] = 1.0
    for name, targets, params in gl:
        psi = O.apply_gate(psi, n, O.gate_matrix(name, params), targets)
    return psi


@pytest.mark.parametrize("n,depth", [(17, 4), (18, 3), (20, 3)])
def test_tma_stream_vs_oracle(n, depth):
    from qsb.bigstate import BigState
    gl = ordered(n, layered_circuit(n, depth, 7 + n))
    st = BigState(n, engine="tma")
    st.apply_gates(gl)
    got = st.to_reference_order([st.local_shard()])
    assert np.max(np.abs(got - oracle_state(n, gl))) < 1e-12
    assert abs(st.norm2() - 1.0) < 1e-12


@pytest.mark.parametrize("m,l,e", [(12, 4, 3), (12, 5, 3), (11, 4, 2), (9, 3, 1), (8, 4, 0), (12, 6, 1)])
def test_tile_geometries(m, l, e):
    """Every (tile bits, row bits, box bits) geometry the API accepts gives the same state (18 qubits, oracle)."""
    from qsb import capi
    n = 18
    gl = ordered(n, layered_circuit(n, 3, 99))
    lw = lower(n, gl, layout="reference")
    cdata = lw.pool.array()
    steps, pos_of, _ = S.plan(lw.items, cdata, n, 0, list(range(n)), local_bits=m, low_bits=l, box_bits=e)
    assert all((st.spass.m, st.spass.l) == (m, l) for st in steps)
    ctx = capi.get_context()
    psi = np.zeros(2 ** n, dtype=np.complex128)
    psi[0] = 1.0
    buf = ctx.to_device(psi)
    handles = [ctx.stream_pass(st.spass, cdata) for st in steps]
    for h in handles:
        h.run(buf)
    ctx.sync()
    mem = buf.download(np.complex128, (2 ** n,))
    pos = [pos_of[lw.bit_of_axis[j]] for j in range(n)]
    got = np.ascontiguousarray(mem.reshape([2] * n).transpose([n - 1 - pos[j] for j in range(n)])).reshape(-1)
    assert np.max(np.abs(got - oracle_state(n, gl))) < 1e-12


def test_tma_engine_equals_executor_engine_22q():
    from qsb.bigstate import BigState
    n = 22
    gl = ordered(n, layered_circuit(n, 6, 5))
    a = BigState(n, engine="tma")
    a.apply_gates(gl)
    b = BigState(n, engine="executor")
    b.apply_gates(gl)
    va = a.to_reference_order([a.local_shard()])
    vb = b.to_reference_order([b.local_shard()])
    assert np.max(np.abs(va - vb)) < 1e-12
    assert abs(a.norm2() - 1.0) < 1e-12


def test_noise_parameters_dense_gates_through_the_pipeline():
    """Pauli draws (reference uniforms), parameter rows and dense 2-/3-qubit matrices, 17 qubits, vs the CPU replay of the
    same plan (itself pinned to the oracle in tests/test_stream_plan.py) and vs unit norm."""
    from qsb import capi
    n = 17
    rng = np.random.default_rng(11)
    lw = Lowering(n, layout="reference")
    u2 = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    u3 = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]
    prm = rng.uniform(-3, 3, 8)
    for q in range(n):
        lw.gate("H", [q], [], REG.get("H").matrix_func)
        lw.kraus("depolarizing", 0.5, q)
    lw.matrix(u2, [16, 2])
    lw.param_gate("Rx", [3], 0)
    lw.param_gate("U3", [8], 1)
    lw.matrix(u3, [0, 13, 4])
    lw.param_gate("Rz", [4], 4)
    lw.gate("Toffoli", [1, 15, 9], [], None)
    lw.gate("Fredkin", [10, 0, 14], [], None)
    lw.gate("SWAP", [11, 3], [], None)
    lw.param_gate("Phase", [1], 5)
    lw.gate("CZ", [12, 5], [], None)
    uni = rng.random(lw.n_draws)
    cdata = lw.pool.array()
    steps, pos_of, _ = S.plan(lw.items, cdata, n, 0, list(range(n)), params=prm, uniforms=uni)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    want = replay(steps, n, 0, psi, cdata)
    ctx = capi.get_context()
    buf = ctx.to_device(psi)
    handles = [ctx.stream_pass(st.spass, cdata) for st in steps]
    for h in handles:
        h.run(buf)
    ctx.sync()
    got = buf.download(np.complex128, (2 ** n,))
    assert np.max(np.abs(got - want)) < 1e-12
    assert abs(np.vdot(got, got).real - 1.0) < 1e-12


def test_reorder_pass_permutes_positions():
    """A pass without sweeps whose store side differs from its load side is a bit permutation of the index."""
    from qsb import capi
    L, m, l, e = 18, 12, 4, 3
    rng = np.random.default_rng(3)
    psi = rng.normal(size=2 ** L) + 1j * rng.normal(size=2 ** L)
    new_of_old = list(range(l)) + (l + rng.permutation(L - l)).tolist()
    sp = S.StreamPass(L, m, l, e, list(range(L)), new_of_old, [])
    ctx = capi.get_context()
    src, dst = ctx.to_device(psi), ctx.alloc(16 << L).zero()
    h = ctx.stream_pass(sp, np.zeros(2))
    h.run(src, dst)
    ctx.sync()
    got = dst.download(np.complex128, (2 ** L,))
    idx = np.arange(2 ** L)
    out_idx = np.zeros_like(idx)
    for p in range(L):
        out_idx |= ((idx >> p) & 1) << new_of_old[p]
    want = np.empty_like(psi)
    want[out_idx] = psi
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        h.run(src)                                        # a permuting store cannot run in place


@pytest.mark.parametrize("g", [1, 2, 3])
def test_peer_loads_fold_the_exchange_into_a_pass(g):
    """qsb_stream_run_peers on ONE device: 2^g shards stand in for the GPUs.  Loading the tiles through the peer maps
    must equal the all-to-all exchange (rank bits <-> top local bits) followed by the same pass."""
    from qsb import capi
    L, world = 17, 1 << g
    rng = np.random.default_rng(40 + g)
    shards = rng.normal(size=(world, 2 ** L)) + 1j * rng.normal(size=(world, 2 ** L))
    gl = ordered(L, layered_circuit(L, 2, 9))
    lw = lower(L, gl, layout="textbook")
    cdata = lw.pool.array()
    steps, _, _ = S.plan(lw.items, cdata, L, 0, list(range(L)))
    sp = steps[0].spass
    ctx = capi.get_context()
    bufs = [ctx.to_device(np.ascontiguousarray(shards[r])) for r in range(world)]
    h = ctx.stream_pass(sp, cdata)
    chunk = 2 ** (L - g)
    for r in range(world):
        exchanged = np.concatenate([shards[c][r * chunk:(r + 1) * chunk] for c in range(world)])
        want_buf = ctx.to_device(exchanged)
        h.run(want_buf)
        out = ctx.alloc(16 << L).zero()
        h.run_peers([b.ptr for b in bufs], L - g, r << (L - g), out)
        ctx.sync()
        assert np.array_equal(out.download(np.complex128, (2 ** L,)), want_buf.download(np.complex128, (2 ** L,))), (g, r)


@pytest.mark.parametrize("g", [1, 2, 3])
def test_scatter_stores_fold_the_exchange_into_a_pass(g):
    """qsb_stream_run_scatter on ONE device: 2^g shards stand in for the GPUs.  A pass (sweeps + position reorder) whose
    boxes are stored straight into the peers' shards must equal the same pass stored locally followed by the all-to-all
    exchange (rank bits <-> top local bits)."""
    from qsb import capi
    n, world = 17 + g, 1 << g
    L = n - g
    rng = np.random.default_rng(60 + g)
    gl = ordered(n, layered_circuit(n, 3, 11))
    lw = lower(n, gl, layout="textbook")
    cdata = lw.pool.array()
    steps, _, _ = S.plan(lw.items, cdata, n, g, list(range(n)))
    st = next(s for s in steps if s.scatter)
    sp = st.spass
    assert sp.positions != sp.positions_out or True
    ctx = capi.get_context()
    shards = rng.normal(size=(world, 2 ** L)) + 1j * rng.normal(size=(world, 2 ** L))
    srcs = [ctx.to_device(np.ascontiguousarray(shards[r])) for r in range(world)]
    dsts = [ctx.alloc(16 << L).zero() for _ in range(world)]
    h = ctx.stream_pass(sp, cdata)
    for r in range(world):
        h.run_scatter(srcs[r], [d.ptr for d in dsts], L - g, r << (L - g))
    ctx.sync()
    got = [d.download(np.complex128, (2 ** L,)) for d in dsts]
    # the same pass stored locally, then the exchange on the host
    local = []
    for r in range(world):
        out = ctx.alloc(16 << L).zero()
        h.run(srcs[r], out)
        ctx.sync()
        local.append(out.download(np.complex128, (2 ** L,)))
    chunk = 2 ** (L - g)
    for r in range(world):
        want = np.concatenate([local[c][r * chunk:(r + 1) * chunk] for c in range(world)])
        assert np.array_equal(got[r], want), (g, r)
    with pytest.raises(ValueError):
        h.run_scatter(srcs[0], [srcs[0].ptr] + [d.ptr for d in dsts[1:]], L - g, 0)     # would overwrite what it reads


def test_argument_checks():
    from qsb import capi
    ctx = capi.get_context()
    ok = S.StreamPass(18, 12, 4, 3, list(range(18)), list(range(18)), [])
    ctx.stream_pass(ok, np.zeros(2))
    for bad in (S.StreamPass(18, 13, 4, 3, list(range(18)), list(range(18)), []),            # tile too large
                S.StreamPass(18, 12, 2, 3, list(range(18)), list(range(18)), []),            # rows shorter than 128 B
                S.StreamPass(18, 12, 4, 3, [1, 0] + list(range(2, 18)), [1, 0] + list(range(2, 18)), []),   # row bits moved
                S.StreamPass(18, 12, 4, 3, list(range(18)), list(range(18)), [],
                             blocks=[S.Block([3, 12, 5, 6], [S.BlockOp(S.B_CX, [0, 1])])]),   # block bit outside the tile
                S.StreamPass(18, 12, 4, 3, list(range(18)), list(range(18)), [],
                             blocks=[S.Block([3, 4, 5, 6], [S.BlockOp(S.B_CCX, [2, 1, 0])])]),  # controls not ascending
                S.StreamPass(18, 12, 4, 3, list(range(18)), list(range(18)), [],
                             blocks=[S.Block([3, 4, 5, 6], [S.BlockOp(S.B_CZ, [1, 1])])])):     # repeated bit
        with pytest.raises(ValueError):
            ctx.stream_pass(bad, np.zeros(2))
    with pytest.raises(ValueError):
        ctx.stream_pass(S.StreamPass(18, 12, 4, 3, list(range(18)), list(range(18)), [],
                                     blocks=[S.Block([3, 4, 5, 6], [])] * (S.MAX_SWEEPS + 1)), np.zeros(2))


def test_every_matrix_mask_of_the_block_sweep():
    """csrc/qsb_stream.cuh instantiates the block sweep per mask of local bits that carry a 2x2 (16 variants); classes
    (real diagonal, complex diagonal, dense) are all applied as full matrices, signs by XOR.  One hand-made block per
    mask -- its matrices, then CX and CZ -- against NumPy."""
    from qsb import capi
    from test_stream_plan import block_op_matrix
    L, m, l, e = 17, 12, 5, 3
    rng = np.random.default_rng(11)
    psi = rng.normal(size=2 ** L) + 1j * rng.normal(size=2 ** L)
    psi /= np.linalg.norm(psi)
    ctx = capi.get_context()

    def matrix(kind):
        if kind == 0:                                             # dense unitary
            return np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]
        if kind == 1:                                             # complex diagonal (Rz-like)
            th = rng.uniform(-3, 3)
            return np.diag([np.exp(-0.5j * th), np.exp(0.5j * th)])
        return np.diag([1.0, rng.uniform(0.5, 1.0)]).astype(np.complex128)   # real diagonal (a damping K0)

    for mask in range(16):
        bits = [int(x) for x in rng.permutation(m)[:4]]
        ops = [S.BlockOp(S.B_MAT1, [t], matrix((mask + t) % 3)) for t in range(4) if (mask >> t) & 1]
        ops += [S.BlockOp(S.B_CX, [3, 0]), S.BlockOp(S.B_CZ, [1, 2])]
        blk = S.Block(bits, ops)
        sp = S.StreamPass(L, m, l, e, list(range(L)), list(range(L)), [], blocks=[blk])
        buf = ctx.to_device(psi)
        ctx.stream_pass(sp, np.zeros(2)).run(buf)
        ctx.sync()
        got = buf.download(np.complex128, (2 ** L,))
        want = psi
        for op in ops:
            want = O.apply_textbook(want, L, block_op_matrix(op, np.zeros(2), -1), [L - 1 - bits[x] for x in op.t])
        assert np.max(np.abs(got - want)) < 1e-12, mask
