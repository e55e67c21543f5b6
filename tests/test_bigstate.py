"""Streamed / sharded statevectors (qsb/bigstate.py, BASELINE.json config 5).

CPU: the planner alone (passes, reorders, exchanges) is checked by replaying its steps with NumPy.
GPU: single-device streamed states against the oracle, and size-independent properties at full size."""

import numpy as np
import pytest

from oracle import qsim_oracle as O
from qsb.bigstate import plan_distributed
from qsb.compiler import Lowering, OP_DTYPE
from qsb.workloads import layered_circuit
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.gate_registry import GateRegistry

REG = GateRegistry.instance()


def ordered(n, gates):
    qc = QuantumCircuit(n) if n <= 16 else QuantumCircuit(16)
    qc.num_qubits = n
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    return [(g.gate_name, list(g.target_qubits), list(g.params)) for col in qc.get_ordered_gates() for g in col]


def lower(n, gl, layout="reference"):
    lw = Lowering(n, layout=layout)
    for name, targets, params in gl:
        lw.gate(name, targets, params, REG.get(name).matrix_func)
    return lw


def replay_numpy(steps, n, g, psi, cdata):
    """Execute a plan with NumPy on the full 2^n vector viewed as 2^g shards (host model of the device path)."""
    L = n - g
    shards = [psi[r << L:(r + 1) << L].copy() for r in range(1 << g)]
    from qsb import compiler as CK
    for st in steps:
        if st.kind == "exchange":
            chunks = [s.reshape(1 << g, -1) for s in shards]
            shards = [np.concatenate([chunks[src][dst] for src in range(1 << g)]) for dst in range(1 << g)]
            continue
        prog = st.prog
        load, store = prog.idata[:L].tolist(), prog.idata[L:2 * L].tolist()
        new = []
        for s in shards:
            t = s.reshape([2] * L)                                   # numpy axis a <-> position L-1-a
            # bring load positions into slot order: slot j = position load[j]; slot 0 = least significant
            t = t.transpose([L - 1 - load[j] for j in reversed(range(L))]).reshape(-1)
            for op in prog.ops:
                kind = int(op["kind"])
                nb = {CK.U1: 1, CK.D1: 1, CK.PX: 1, CK.PY: 1, CK.PZ: 1, CK.U2: 2, CK.U3Q: 3, CK.CX: 2, CK.CZ: 2,
                      CK.SWAP: 2, CK.CCX: 3, CK.CSWAP: 3}[kind]
                bits = [int(op["b0"]), int(op["b1"]), int(op["b2"])][:nb]
                if kind == CK.U1:
                    mat = cdata[op["data"]:op["data"] + 8].view(np.complex128).reshape(2, 2)
                elif kind == CK.D1:
                    mat = np.diag(cdata[op["data"]:op["data"] + 4].view(np.complex128))
                else:
                    mat = O.gate_matrix({CK.PX: "X", CK.PY: "Y", CK.PZ: "Z", CK.CX: "CNOT", CK.CZ: "CZ", CK.SWAP: "SWAP",
                                         CK.CCX: "Toffoli", CK.CSWAP: "Fredkin"}[kind])
                t = O.apply_textbook(t, L, mat, [L - 1 - b for b in bits])     # slot bit b = textbook qubit L-1-b
            out = np.empty_like(t).reshape([2] * L)
            # slot j -> position store[j]
            src = t.reshape([2] * L)                                   # axis a <-> slot L-1-a
            perm = [0] * L
            for j in range(L):
                perm[L - 1 - store[j]] = L - 1 - j
            new.append(np.ascontiguousarray(src.transpose(perm)).reshape(-1))
        shards = new
    return np.concatenate(shards)


@pytest.mark.parametrize("n,g,m", [(7, 0, 4), (8, 1, 4), (9, 2, 4), (10, 3, 5)])
def test_plan_replay_matches_oracle(n, g, m):
    rng = np.random.default_rng(n)
    gl = ordered(n, layered_circuit(n, 5, 40 + n))
    lw = lower(n, gl)
    steps, pos_of = plan_distributed(lw, g, local_bits=m)
    kinds = [s.kind for s in steps]
    if g:
        assert "exchange" in kinds
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    got_mem = replay_numpy(steps, n, g, psi, lw.pool.array())
    ref = psi
    for name, targets, params in gl:
        ref = O.apply_gate(ref, n, O.gate_matrix(name, params), targets)
    # memory position of reference axis j
    pos = [pos_of[lw.bit_of_axis[j]] for j in range(n)]
    got = np.ascontiguousarray(got_mem.reshape([2] * n).transpose([n - 1 - pos[j] for j in range(n)])).reshape(-1)
    assert np.max(np.abs(got - ref)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("n", [17, 18, 20])
def test_streamed_state_vs_oracle(n):
    from qsb.bigstate import BigState
    gl = ordered(n, layered_circuit(n, 3, 7 + n))
    st = BigState(n)
    st.apply_gates(gl)
    got = st.to_reference_order([st.local_shard()])
    ref = np.zeros(2 ** n, dtype=np.complex128)
    ref[0] = 1.0
    for name, targets, params in gl:
        ref = O.apply_gate(ref, n, O.gate_matrix(name, params), targets)
    assert np.max(np.abs(got - ref)) < 1e-12
    assert abs(st.norm2() - 1.0) < 1e-12


@pytest.mark.gpu
def test_streamed_state_inverse_property_26q():
    """Full-size property (no 2^n oracle needed): circuit followed by its inverse returns |0...0>."""
    from qsb.bigstate import BigState
    n = 26
    gl = ordered(n, layered_circuit(n, 4, 2026))
    inv = []
    for name, targets, params in reversed(gl):
        if name in ("Rx", "Ry", "Rz"):
            inv.append((name, targets, [-params[0]]))
        elif name == "U3":
            th, ph, la = params
            inv.append(("U3", targets, [-th, -la, -ph]))
        else:
            inv.append((name, targets, params))           # H, CNOT, CZ, Toffoli are involutions
    st = BigState(n, layout="textbook")
    st.apply_gates(gl)
    assert abs(st.norm2() - 1.0) < 1e-10
    st.apply_gates(inv)
    psi = st.to_reference_order([st.local_shard()])
    assert abs(abs(psi[0]) - 1.0) < 1e-10
    assert np.max(np.abs(psi[1:])) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("g", [1, 2])
def test_peer_table_load_folds_the_exchange_into_a_pass(g):
    """`qsb_run_args.peer_table` on ONE device: 2^g shards stand in for the GPUs of a sharded state.  A streamed pass
    that loads through the table must equal exchange (rank bits <-> top local bits, the all-to-all of
    `exchange_rank_bits`) followed by the same pass on the exchanged shard."""
    from qsb import capi
    from qsb.lowering import lower_circuit
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.gate_registry import GateRegistry
    L, m, world = 11, 5, 1 << g
    rng = np.random.default_rng(40 + g)
    shards = rng.normal(size=(world, 2 ** L)) + 1j * rng.normal(size=(world, 2 ** L))
    gates = layered_circuit(L, 2, 9)
    qc = QuantumCircuit(L)
    for x in gates:
        qc.add_gate(GateInstance(x[0], list(x[1]), list(x[2]), x[3]))
    plan, _ = lower_circuit(L, qc.get_ordered_gates(), GateRegistry.instance(), local_bits=m, stream=True, layout="textbook")
    prog = plan.passes[0]
    ctx = capi.get_context()
    bufs = [ctx.to_device(np.ascontiguousarray(shards[r])) for r in range(world)]
    table = ctx.to_device(np.array([b.ptr for b in bufs], dtype=np.int64))
    dp = ctx.program(prog)
    chunk = 2 ** (L - g)
    for r in range(world):
        # what the all-to-all would leave on rank r: chunk c of the new shard = chunk r of rank c's shard
        exchanged = np.concatenate([shards[c][r * chunk:(r + 1) * chunk] for c in range(world)])
        want_buf = ctx.to_device(exchanged)
        ctx.run(dp, 1, states=want_buf, load=True, store=True)
        want = want_buf.download(np.complex128, (2 ** L,))
        out = ctx.alloc(16 << L).zero()
        ctx.run(dp, 1, states=bufs[r], load=True, store=True, states_out=out, peer_table=table, peer_shift=L - g,
                peer_rank_or=r << (L - g))
        got = out.download(np.complex128, (2 ** L,))
        assert np.array_equal(got, want), (g, r)                # same kernel arithmetic on the same inputs: bit-identical
    # argument checking: a resident program (no streaming) refuses a peer table
    from qsb.compiler import Lowering
    lw = Lowering(4)
    lw.matrix(np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2), [0])
    small = ctx.program(lw.finish())
    st = ctx.alloc(16 << 4).zero()
    with pytest.raises(ValueError):
        ctx.run(small, 1, states=st, load=True, store=True, states_out=ctx.alloc(16 << 4), peer_table=table, peer_shift=3)
