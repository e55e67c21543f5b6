"""Host compiler for streamed passes (states beyond the resident executor's 16 qubits; BASELINE config 5).

Input: the physical-bit op list of `compiler.Lowering` (the reference's `apply_gate` calls, state_vector.py:41-74,
with the axis scramble of :66-73 already turned into bookkeeping).  Output: a list of steps for `bigstate.BigState`

    pass      one launch of the TMA tile pipeline (csrc/qsb_stream.cuh): every tile of 2^m amplitudes is gathered into
              shared memory, takes the pass's sweeps, and is scattered back
    reorder   a pass without sweeps whose store permutes index positions (brings exchange victims to the top)
    exchange  rank positions <-> top local positions (NCCL all-to-all, or folded into the next pass's peer loads)
A pass / reorder step with `scatter` set carries the exchange in its own store: its output positions already have the
leaving qubits on top, and on the device every TMA box is written into the shard of the peer it belongs to
(qsb_stream_run_scatter) -- the compute pass, the reorder and the all-to-all are ONE kernel.

Gate fusion happens HERE, once per program, because in a streamed pass every tile sees the same ops:
  * a 1-qubit op (gate, Pauli Kraus branch) never costs a sweep: it is multiplied into the PENDING 2x2 of its
    qubit, and that matrix rides into the next multi-qubit sweep that touches the qubit -- in whatever pass that
    happens, so 1-qubit ops put no demand on tile residency either;
  * what is still pending when the program ends is flushed by sweeps of up to three qubits.
Only multi-qubit gates (and those final flushes) are packed into passes.  This module multiplies 2x2 matrices and
moves no amplitudes: it is the compile step of the path, like `matrix_func(*params)` in the reference
(simulator.py:110-114).
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import compiler as K

CLS_NONE, CLS_RDIAG, CLS_DIAG, CLS_DENSE = 0, 1, 2, 3
G_NONE, G_CX, G_CZ, G_SWAP, G_CCX, G_CSWAP, G_DENSE = 0, 1, 2, 3, 4, 5, 6
B_MAT1, B_CX, B_CZ, B_SWAP, B_CCX, B_CSWAP, B_DENSE2, B_DENSE3 = 1, 2, 3, 4, 5, 6, 7, 8     # include/qsb.h QSB_B_*
MAX_SWEEPS = 16               # gates (or final flush groups) per pass: at most that many block sweeps come out
MAX_BLOCK_OPS = 12
MAX_TILE_BITS = 12

_GATE_OF = {K.CX: G_CX, K.CZ: G_CZ, K.SWAP: G_SWAP, K.CCX: G_CCX, K.CSWAP: G_CSWAP, K.U2: G_DENSE, K.U3Q: G_DENSE}
_ONE_QUBIT = (K.U1, K.D1, K.PX, K.PY, K.PZ, K.RX, K.RY, K.RZ, K.PHASE, K.U3, K.KRAUS_PAULI)
FLUSH = -1            # pseudo op kind: apply the pending matrix of one bit

_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_I = np.eye(2, dtype=np.complex128)


class StreamOp(C.Structure):
    """include/qsb.h: qsb_stream_op."""
    _fields_ = [("kind", C.c_int32), ("t", C.c_int32 * 3), ("cls", C.c_int32), ("pad", C.c_int32 * 3), ("U", C.c_double * 8)]


class StreamBlock(C.Structure):
    """include/qsb.h: qsb_stream_block."""
    _fields_ = [("n_ops", C.c_int32), ("b", C.c_int32 * 4), ("mat", C.c_int32), ("pad", C.c_int32 * 2),
                ("ops", StreamOp * MAX_BLOCK_OPS)]


@dataclass
class BlockOp:
    kind: int                     # B_*
    t: list                       # local bits of the register block
    U: object = None              # B_MAT1: 2x2


@dataclass
class Block:
    b: list                       # tile slot bit of local bit 0..3
    ops: list
    mat: int = -1                 # cdata offset of the block's dense matrix (B_DENSE2 / B_DENSE3)


@dataclass
class Sweep:
    gate: int
    bits: list                    # index POSITIONS while planning; turned into slot bits when the pass is closed
    P: list                       # one 2x2 (or None) per bit
    mat: int = -1                 # cdata offset of a dense gate matrix


@dataclass
class StreamPass:
    n: int                        # index bits of the (local) shard
    m: int
    l: int
    e: int
    positions: list               # slot j -> index position (load side)
    positions_out: list           # slot j -> index position (store side); == positions for an in-place pass
    sweeps: list = field(default_factory=list)      # one per multi-qubit gate / flush group, program order
    blocks: list = None           # what the device runs: the sweeps grouped into register blocks (group_blocks)

    def __post_init__(self):
        if self.blocks is None:
            self.blocks = group_blocks(self.sweeps, self.m)

    @property
    def in_place(self):
        return self.positions == self.positions_out


@dataclass
class Step:
    kind: str                     # "pass" | "reorder" | "exchange"
    spass: StreamPass = None
    handle: object = None         # device object, filled in by the runner
    scatter: bool = False         # the qubit exchange follows this pass and is folded into its STORE (peer-mapped TMA stores)
    meta: object = None           # planner bookkeeping: (resident positions, chosen ops) of a pass that may be re-closed


def classify(P):
    """Structure class of a pending 2x2 (what the sweep has to do for it)."""
    if P is None:
        return CLS_NONE
    if P[0, 1] == 0 and P[1, 0] == 0:
        if P[0, 0] == 1 and P[1, 1] == 1:
            return CLS_NONE
        if P[0, 0] == 1 and P[1, 1].imag == 0:
            return CLS_RDIAG
        return CLS_DIAG
    return CLS_DENSE


def philox_uniform(seed, traj, draw):
    """Philox4x32-10 uniform of (trajectory, draw): the same stream as the kernels' qsb_philox_uniform."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [traj & 0xFFFFFFFF, (traj >> 32) & 0xFFFFFFFF, draw >> 1, 0]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    lo, hi = c[2 * (draw & 1)], c[2 * (draw & 1) + 1]
    return ((hi >> 5) * 67108864.0 + (lo >> 6)) * (1.0 / 9007199254740992.0)


def one_qubit_matrix(kind, data, param, draw, cdata, params, uniforms, seed, traj):
    """2x2 of a 1-qubit op (gates.py:37-94; Pauli channels: noise.py:39-83 with the branch choice() would take);
    None = identity (a Kraus draw that picked the identity branch)."""
    if kind == K.U1:
        v = cdata[data:data + 8]
        return (v[0::2] + 1j * v[1::2]).reshape(2, 2)
    if kind == K.D1:
        v = cdata[data:data + 4]
        return np.diag([v[0] + 1j * v[1], v[2] + 1j * v[3]])
    if kind == K.PX:
        return _X
    if kind == K.PY:
        return _Y
    if kind == K.PZ:
        return _Z
    if kind in (K.RX, K.RY, K.RZ, K.PHASE, K.U3):
        if params is None:
            raise ValueError("the program has parameterised gates but no parameter row was given")
        t = float(params[param])
        if kind == K.RX:
            c, s = math.cos(t / 2), math.sin(t / 2)
            return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
        if kind == K.RY:
            c, s = math.cos(t / 2), math.sin(t / 2)
            return np.array([[c, -s], [s, c]], dtype=np.complex128)
        if kind == K.RZ:
            return np.diag([np.exp(-1j * t / 2), np.exp(1j * t / 2)])
        if kind == K.PHASE:
            return np.diag([1.0, np.exp(1j * t)])
        phi, lam = float(params[param + 1]), float(params[param + 2])
        c, s = math.cos(t / 2), math.sin(t / 2)
        return np.array([[c, -np.exp(1j * lam) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]],
                        dtype=np.complex128)
    if kind == K.KRAUS_PAULI:
        u = float(uniforms[draw]) if uniforms is not None else philox_uniform(seed, traj, draw)
        thr = cdata[data:data + 3]
        code = int(cdata[data + 3 + int(np.sum(thr <= u))])
        return (None, _X, _Y, _Z)[code]
    raise NotImplementedError(f"op kind {kind} in a streamed pass")


def fuse(items, cdata, *, params=None, uniforms=None, seed=0, traj=0, pending=None):
    """Scan the op list once: fold every 1-qubit op into the pending matrix of its bit and attach the pending
    matrices to the multi-qubit op that consumes them.  Returns (multi, pending): multi = [(gate, bits, [P per bit],
    mat offset)] in program order; pending = {bit: 2x2} still to be flushed."""
    pending = dict(pending or {})
    multi = []
    for kind, bits, data, param, draw, _ in items:
        if kind in (K.SNAPSHOT, K.KRAUS_AD, K.KRAUS_GEN):
            raise NotImplementedError("snapshots and state-dependent Kraus draws need the whole state resident")
        if kind in _ONE_QUBIT:
            U = one_qubit_matrix(kind, data, param, draw, cdata, params, uniforms, seed, traj)
            if U is not None:
                b = bits[0]
                pending[b] = U @ pending[b] if b in pending else U
            continue
        if kind not in _GATE_OF:
            raise NotImplementedError(f"op kind {kind} in a streamed pass")
        Ps = [pending.pop(b, None) for b in bits]
        multi.append((_GATE_OF[kind], list(bits), Ps, data if _GATE_OF[kind] == G_DENSE else -1))
    return multi, pending


def choose_geometry(L, local_bits=None, low_bits=None, box_bits=None):
    """(m, l, e) for a shard of L index bits: 64 KiB tiles, 512-byte rows, 4 KiB per TMA op by default (measured on the
    26-qubit layered circuit: 12.2 ms against 12.9 ms with 256-byte rows, although those need two passes fewer)."""
    m = min(L, MAX_TILE_BITS) if local_bits is None else int(local_bits)
    if not 3 <= m <= min(L, MAX_TILE_BITS):
        raise ValueError(f"local_bits {m} invalid for {L} local index bits (3..{min(L, MAX_TILE_BITS)})")
    # the device wants m >= 6 and l >= 3 (qsb_stream_create checks); smaller geometries only exist in planner tests
    l = max(0, min(5 if low_bits is None else int(low_bits), m - 3))
    e = min(3 if box_bits is None else int(box_bits), m - l)
    while m - l - e > 5:                      # at most 32 TMA ops per tile
        e += 1
        if e > 3:
            e, l = 3, l + 1
    return m, l, e


def plan(items, cdata, n, g, pos_of, *, local_bits=None, low_bits=None, box_bits=None, params=None, uniforms=None,
         seed=0, traj=0, flush=True, fuse_store=True):
    """Op list (bits = VIRTUAL bits) -> (steps, pos_of, pending).

    n = all index bits, g = rank positions n-g..n-1 (0: one device), pos_of[v] = position of virtual bit v now.
    Positions 0..L-1 (L = n - g) address the local shard.  `flush` applies whatever is pending at the end.
    fuse_store: fold every exchange (and the reorder that precedes it) into the store of the pass before it
    (`Step.scatter`); otherwise emit separate "reorder" and "exchange" steps."""
    L = n - g
    m, l, e = choose_geometry(L, local_bits, low_bits, box_bits)
    if g > 0 and L - l < 2 * g:
        raise ValueError("need at least as many movable local bits as rank bits to exchange them")
    pos_of = list(pos_of)
    multi, pending = fuse(items, cdata, params=params, uniforms=uniforms, seed=seed, traj=traj)
    todo = [(gate, bits, Ps, mat) for gate, bits, Ps, mat in multi]
    if flush:
        todo += [(FLUSH, [b], [P], -1) for b, P in sorted(pending.items()) if classify(P) != CLS_NONE]
        pending = {}
    steps = []

    def close_pass(resident, chosen, store_map=None, victims=()):
        """Emit one pass.  store_map (position -> position) makes the store permute positions; `victims` are resident
        positions that must not ride in the TMA box (they become peer-selecting bits on the store side)."""
        low = list(range(l))
        rest = sorted(p for p in resident if p >= l and p not in victims)
        resident_order = low + rest + sorted(p for p in resident if p in victims)
        others = [p for p in range(L) if p not in resident]
        positions = resident_order + others
        slot = {p: j for j, p in enumerate(positions)}
        sweeps, flushes = [], []
        for gate, bits, Ps, mat in chosen:
            ps = [pos_of[v] for v in bits]
            if gate == FLUSH:
                flushes.append((slot[ps[0]], Ps[0]))
                continue
            sweeps.append(Sweep(gate, [slot[p] for p in ps], Ps, mat))
        for i in range(0, len(flushes), 3):
            grp = flushes[i:i + 3]
            sweeps.append(Sweep(G_NONE, [b for b, _ in grp], [P for _, P in grp]))
        out = list(positions) if store_map is None else [store_map[p] for p in positions]
        return Step("pass" if chosen else "reorder", StreamPass(L, m, l, e, positions, out, sweeps),
                    scatter=store_map is not None and fuse_store, meta=(set(resident), list(chosen)))

    while todo:
        resident = set(range(l))
        blocked = set()
        chosen, rest = [], []
        n_sweeps = n_flush = 0
        for op in todo:
            ps = [pos_of[v] for v in op[1]]
            if any(p in blocked for p in ps) or any(p >= L for p in ps):
                blocked.update(ps)
                rest.append(op)
                continue
            need = [p for p in ps if p not in resident]
            cost = n_sweeps + (0 if op[0] == FLUSH else 1) + (n_flush + (1 if op[0] == FLUSH else 0) + 2) // 3
            if len(resident) + len(need) <= m and cost <= MAX_SWEEPS:
                resident.update(need)
                chosen.append(op)
                if op[0] == FLUSH:
                    n_flush += 1
                else:
                    n_sweeps += 1
            else:
                blocked.update(ps)
                rest.append(op)
        if chosen:
            for p in range(L):                       # fill the tile with the lowest free positions
                if len(resident) >= m:
                    break
                resident.add(p)
            steps.append(close_pass(resident, chosen))
            todo = rest
            continue
        if g == 0:
            raise ValueError(f"an op needs more than {m} resident bits")
        # ---- every runnable op needs a qubit that lives in a rank position: bring all g of them in.  Victims = the g
        # local virtual bits (above the row bits, which never move) whose next use is farthest away; a reorder pass
        # takes them to the top local positions, then the exchange swaps those with the rank positions.
        next_use = {}
        for order, op in enumerate(todo):
            for v in op[1]:
                next_use.setdefault(v, order)
        movable = [v for v in range(n) if l <= pos_of[v] < L]
        victims = sorted(movable, key=lambda v: (-next_use.get(v, 1 << 60), -pos_of[v]))[:g]
        vic_pos = sorted(pos_of[v] for v in victims)
        keep = [p for p in range(l, L) if p not in vic_pos]
        new_of_old = {p: p for p in range(l)}
        for j, p in enumerate(keep):
            new_of_old[p] = l + j
        for j, p in enumerate(vic_pos):
            new_of_old[p] = L - g + j
        identity = vic_pos == list(range(L - g, L))
        if fuse_store and steps and steps[-1].kind == "pass" and not steps[-1].scatter:
            # the pass just before the exchange does the reorder AND the exchange in its store
            resident, chosen = steps.pop().meta
            steps.append(close_pass(resident, chosen, store_map=new_of_old, victims=vic_pos))
        elif fuse_store:
            # nothing to ride on: a pass without sweeps (its tile = the low positions, victims excluded from the box)
            resident = set(range(l)) | set(keep[:m - l])
            steps.append(close_pass(resident, [], store_map=new_of_old, victims=vic_pos))
        else:
            if not identity:
                positions = list(range(L))           # slot j = position j on the load side
                steps.append(Step("reorder", StreamPass(L, m, l, e, positions, [new_of_old[p] for p in positions], [])))
            steps.append(Step("exchange"))
        for v in range(n):
            if pos_of[v] < L:
                pos_of[v] = new_of_old[pos_of[v]]
        for v in range(n):
            p = pos_of[v]
            if p >= L:
                pos_of[v] = p - g
            elif p >= L - g:
                pos_of[v] = p + g
    return steps, pos_of, pending


def _sweep_ops(sw):
    """Primitive ops of one sweep on its SLOT bits: the pending 2x2s, then the gate."""
    ops = [("mat1", [b], P) for b, P in zip(sw.bits, sw.P) if classify(P) != CLS_NONE]
    if sw.gate != G_NONE:
        ops.append((sw.gate, list(sw.bits), sw.mat))
    return ops


def group_blocks(sweeps, m):
    """Sweeps (program order) -> register blocks of four slot bits in CANONICAL form: inside a block every 2x2 comes
    before any gate on its bit, so the device runs "2x2s, then one composed permutation / sign pattern" as straight-line
    code.  A sweep joins the open block when the union of bits stays within four and none of its pending 2x2s sits on a
    bit a gate of the block has already used; a sweep that does not fit blocks its bits (later sweeps on those bits
    wait, sweeps on other bits commute past it).  A dense 2- / 3-qubit gate gets a block of its own with its bits on
    the canonical local positions (3, 2[, 1])."""
    remaining = list(sweeps)
    blocks = []
    while remaining:
        bits, ops, dense = [], [], None
        blocked, touched, rest = set(), set(), []
        for sw in remaining:
            sb = set(sw.bits)
            mine = _sweep_ops(sw)
            if sb & blocked or dense is not None:
                blocked |= sb
                rest.append(sw)
                continue
            if sw.gate == G_DENSE:
                if ops:                                   # only at the head of a block
                    blocked |= sb
                    rest.append(sw)
                    continue
                bits, ops, dense = list(sw.bits), mine, sw
                continue
            union = bits + [b for b in sw.bits if b not in bits]
            pend = {b for b, P in zip(sw.bits, sw.P) if classify(P) != CLS_NONE}
            if len(union) <= 4 and len(ops) + len(mine) <= MAX_BLOCK_OPS and not (pend & touched):
                bits, ops = union, ops + mine
                touched |= sb
            else:
                blocked |= sb
                rest.append(sw)
        # filler bits: the highest free tile bits (positions 0..5 are the bank-swizzle pairs; leaving them to the group
        # numbers keeps the LDS.128 / STS.128 of the sweep conflict-free)
        free = [b for b in range(m - 1, -1, -1) if b not in bits]
        if dense is not None:
            local = list(reversed(bits))                  # targets[0] -> local bit 3, targets[1] -> 2, targets[2] -> 1
            b4 = free[:4 - len(bits)][::-1] + local
        else:
            b4 = bits + free[:4 - len(bits)]
        loc = {b: i for i, b in enumerate(b4)}
        bops, mat = [], -1
        for kind, obits, extra in ops:
            t = [loc[b] for b in obits]
            if kind == "mat1":
                bops.append(BlockOp(B_MAT1, t, extra))
            elif kind == G_CX:
                bops.append(BlockOp(B_CX, t))
            elif kind == G_CZ:
                bops.append(BlockOp(B_CZ, t))
            elif kind == G_SWAP:
                bops.append(BlockOp(B_SWAP, t))
            elif kind == G_CCX:
                bops.append(BlockOp(B_CCX, sorted(t[:2]) + [t[2]]))
            elif kind == G_CSWAP:
                bops.append(BlockOp(B_CSWAP, [t[0]] + sorted(t[1:])))
            elif kind == G_DENSE:
                bops.append(BlockOp(B_DENSE2 if len(obits) == 2 else B_DENSE3, t))
                mat = extra
        blocks.append(Block(b4, bops, mat))
        remaining = rest
    return blocks


def pack_blocks(spass):
    """StreamPass -> ctypes array of qsb_stream_block."""
    arr = (StreamBlock * max(len(spass.blocks), 1))()
    for i, bk in enumerate(spass.blocks):
        s = arr[i]
        s.n_ops, s.mat = len(bk.ops), bk.mat
        for j in range(4):
            s.b[j] = bk.b[j]
        for q, op in enumerate(bk.ops):
            o = s.ops[q]
            o.kind = op.kind
            for j, t in enumerate(op.t):
                o.t[j] = t
            if op.kind == B_MAT1:
                o.cls = classify(op.U)
                flat = np.ascontiguousarray(op.U, dtype=np.complex128).reshape(-1)
                for z in range(4):
                    o.U[2 * z], o.U[2 * z + 1] = flat[z].real, flat[z].imag
    return arr
