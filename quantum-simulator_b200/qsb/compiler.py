"""Host compiler: reference-level operations -> `qsb_op` programs (include/qsb.h).

Three jobs, all bookkeeping (no amplitudes are touched here):

1. Track the reference's axis scramble.  `StateVector.apply_gate`
   (state_vector.py:66-73) transposes its tensordot result with
   ``argsort(dest_order)`` instead of ``dest_order``, so after every call array
   axis ``i`` holds textbook qubit ``sigma[i]``.  The device never moves data
   for this: we keep ``bit_of_axis[j]`` = physical bit currently playing
   reference axis ``j``, aim each op at ``bit_of_axis[t]`` and update
   ``bit_of_axis = bit_of_axis[sigma]``.  One bit permutation on store puts
   amplitudes back in reference order.

2. Pick device op kinds: structured gates (CNOT, CZ, Toffoli, ...) by name,
   diagonal gates as QSB_OP_D1, everything else dense; Kraus channels
   (noise.py:27-103) as PAULI / AD / GEN with their `choice()` cdf.

3. Place physical bits into tile slots.  With a cluster of C = 2^(n-m) CTAs the
   top n-m slot bits are the CTA rank; an op whose target sits there is
   preceded by a QSB_OP_REMAP that swaps it with the resident bit whose next
   use is farthest away (Belady).
"""

from __future__ import annotations

import bisect
from dataclasses import dataclass, field

import numpy as np

OP_DTYPE = np.dtype([("kind", "<i4"), ("b0", "<i4"), ("b1", "<i4"), ("b2", "<i4"),
                     ("data", "<i4"), ("param", "<i4"), ("draw", "<i4"), ("aux", "<i4")])

# op kinds (include/qsb.h)
NOP, U1, U2, U3Q, D1 = 0, 1, 2, 3, 4
PX, PY, PZ, CX, CZ, SWAP, CCX, CSWAP = 10, 11, 12, 13, 14, 15, 16, 17
RX, RY, RZ, PHASE, U3 = 20, 21, 22, 23, 24
KRAUS_PAULI, KRAUS_AD, KRAUS_GEN = 30, 31, 32
REMAP, SNAPSHOT = 40, 50

MAX_QUBITS = 16
MAX_LOCAL_BITS = 13

_STRUCTURED = {"X": PX, "Y": PY, "Z": PZ, "CNOT": CX, "CZ": CZ, "SWAP": SWAP,
               "Toffoli": CCX, "Fredkin": CSWAP}
_STRUCT_ARITY = {PX: 1, PY: 1, PZ: 1, CX: 2, CZ: 2, SWAP: 2, CCX: 3, CSWAP: 3}
_PARAM_KIND = {"Rx": RX, "Ry": RY, "Rz": RZ, "Phase": PHASE, "U3": U3}
_PARAM_COUNT = {"Rx": 1, "Ry": 1, "Rz": 1, "Phase": 1, "U3": 3}
_PAULI_CODES = {"bit_flip": [0, 1], "phase_flip": [0, 3], "depolarizing": [0, 1, 2, 3]}


def sigma(n, targets):
    """Axis permutation left behind by the reference's apply_gate (state_vector.py:66-73)."""
    k = len(targets)
    tset = set(targets)
    dest = [0] * n
    for i, q in enumerate(targets):
        dest[q] = i
    i = k
    for q in range(n):
        if q not in tset:
            dest[q] = i
            i += 1
    inv = [0] * n
    for q, d in enumerate(dest):
        inv[d] = q
    return [inv[inv[i]] for i in range(n)]


def default_local_bits(n):
    return min(n, MAX_LOCAL_BITS)


@dataclass
class Program:
    """Lowered program, ready for qsb_program_create."""
    n: int
    m: int
    ops: np.ndarray
    cdata: np.ndarray
    idata: np.ndarray
    load_perm: int
    store_perm: int
    n_snapshots: int
    n_draws: int
    n_params: int
    normalize: bool
    n_gate_ops: int = 0          # executed reference gates (for gate-apps accounting)
    n_kraus_ops: int = 0
    n_remaps: int = 0
    ops_stride: int = 0
    n_programs: int = 1
    meta: dict = field(default_factory=dict)


class _Pool:
    """cdata pool with de-duplication; dense matrices start on even offsets (16-byte c128 loads)."""

    def __init__(self):
        self.buf = []
        self.index = {}

    def add(self, values, align2=False):
        arr = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        key = (arr.tobytes(), align2)
        if key in self.index:
            return self.index[key]
        if align2 and len(self.buf) % 2:
            self.buf.append(0.0)
        off = len(self.buf)
        self.buf.extend(arr.tolist())
        self.index[key] = off
        return off

    def array(self):
        return np.array(self.buf if self.buf else [0.0, 0.0], dtype=np.float64)


def _cplx(mat):
    m = np.ascontiguousarray(mat, dtype=np.complex128).reshape(-1)
    return np.stack([m.real, m.imag], axis=1).reshape(-1)


def choice_cdf(weights):
    """Thresholds numpy's Generator.choice compares its uniform against
    (noise.py:248-254): p / p.sum(), cumsum, / cdf[-1]."""
    p = np.asarray(weights, dtype=np.float64)
    tot = p.sum()
    if tot > 1e-15:
        p = p / tot
    cdf = np.cumsum(p)
    cdf = cdf / cdf[-1]
    return cdf


class Lowering:
    """Accumulates reference-level operations for one n-qubit register and lowers them.

    All qubit arguments are REFERENCE axes (qubit q of the reference API)."""

    def __init__(self, n, layout="reference"):
        if n < 1:
            raise ValueError(f"num_qubits must be >= 1, got {n}")
        self.n = n
        self.layout = layout
        self.bit_of_axis = [n - 1 - j for j in range(n)]
        self.items = []          # (kind, pbits, data spec...) in physical bits
        self.pool = _Pool()
        self.n_draws = 0
        self.n_params = 0
        self.n_snapshots = 0
        self.normalize = False
        self.n_gate_ops = 0
        self.n_kraus_ops = 0

    # -- bookkeeping ---------------------------------------------------------------
    def _check(self, targets):
        n = self.n
        for q in targets:
            if q < 0 or q >= n:
                raise ValueError(f"Qubit index {q} out of range [0, {n-1}]")   # state_vector.py:50-52
        if len(set(targets)) != len(targets):
            raise ValueError(f"repeated qubit in {list(targets)}")

    def _advance(self, targets):
        if self.layout == "reference":
            s = sigma(self.n, list(targets))
            self.bit_of_axis = [self.bit_of_axis[s[i]] for i in range(self.n)]

    def _emit(self, kind, targets, data=-1, param=-1, draw=-1):
        pbits = [self.bit_of_axis[t] for t in targets]
        self.items.append([kind, pbits, data, param, draw, None])
        self._advance(targets)

    # -- gates ----------------------------------------------------------------------
    def matrix(self, mat, targets):
        """apply_gate(matrix, targets) (state_vector.py:41-74)."""
        targets = list(targets)
        self._check(targets)
        k = len(targets)
        mat = np.asarray(mat, dtype=np.complex128)
        if mat.size != 4 ** k:
            raise ValueError(f"cannot reshape array of size {mat.size} into shape {tuple([2] * (2 * k))}")
        mat = mat.reshape(2 ** k, 2 ** k)
        self.n_gate_ops += 1
        if k == 1:
            if mat[0, 1] == 0 and mat[1, 0] == 0:
                self._emit(D1, targets, self.pool.add(_cplx([mat[0, 0], mat[1, 1]])))
            else:
                self._emit(U1, targets, self.pool.add(_cplx(mat)))
        elif k == 2:
            self._emit(U2, targets, self.pool.add(_cplx(mat), align2=True))
        elif k == 3:
            self._emit(U3Q, targets, self.pool.add(_cplx(mat), align2=True))
        else:
            self._dense_big(mat, targets)

    def _dense_big(self, mat, targets):
        """k > 3 (only Pauli-string observables reach this, optimizer.py:155-161): factor a Kronecker
        product of 2x2 blocks into 1-qubit ops; a general dense k > 3 operator is not supported yet."""
        k = len(targets)
        factors = _kron_factors(mat, k)
        if factors is None:
            raise NotImplementedError(f"dense {k}-qubit operators that are not Kronecker products of "
                                      "1-qubit operators are not supported by the device executor")
        # textbook action factorises per qubit; the scramble is taken once for the whole target list
        pb = [self.bit_of_axis[t] for t in targets]
        for f, b in zip(factors, pb):
            if f[0, 1] == 0 and f[1, 0] == 0:
                self.items.append([D1, [b], self.pool.add(_cplx([f[0, 0], f[1, 1]])), -1, -1, None])
            else:
                self.items.append([U1, [b], self.pool.add(_cplx(f)), -1, -1, None])
        self._advance(targets)

    def gate(self, name, targets, params=(), matrix_func=None, builtin=True):
        """One executed GateInstance (simulator.py:110-114).  `matrix_func` is the registry's factory
        for names this module has no structured form for; `builtin` = the registry still maps `name` to its
        built-in definition (a re-registered name must run its own matrix, never the structured kernel)."""
        targets = list(targets)
        if not builtin:
            if matrix_func is None:
                raise KeyError(f"Gate '{name}' not found in registry")
            self.matrix(matrix_func(*params), targets)
            return
        if name in _STRUCTURED and len(targets) == _STRUCT_ARITY[_STRUCTURED[name]]:
            self._check(targets)
            self.n_gate_ops += 1
            self._emit(_STRUCTURED[name], targets)
            return
        if name == "I" and len(targets) == 1:
            self._check(targets)
            self.n_gate_ops += 1
            self._advance(targets)          # identity still leaves the scramble behind
            return
        if matrix_func is None:
            raise KeyError(f"Gate '{name}' not found in registry")
        self.matrix(matrix_func(*params), targets)

    def param_gate(self, name, targets, param_offset):
        """Rx/Ry/Rz/Phase/U3 whose angle(s) come from the per-state parameter row."""
        targets = list(targets)
        self._check(targets)
        if len(targets) != 1:
            raise ValueError("parameterised gates act on one qubit")
        self.n_gate_ops += 1
        self._emit(_PARAM_KIND[name], targets, param=param_offset)
        self.n_params = max(self.n_params, param_offset + _PARAM_COUNT[name])

    # -- noise -----------------------------------------------------------------------
    def kraus(self, kind, p, qubit, kraus_ops=None):
        """One (channel, qubit) draw of NoiseModel._apply_channel (noise.py:235-260)."""
        if qubit >= self.n:
            return                                   # noise.py:235-236: skipped, no draw
        self._check([qubit])
        draw = self.n_draws
        self.n_draws += 1
        self.n_kraus_ops += 1
        self.normalize = True
        if kind in _PAULI_CODES:
            codes = _PAULI_CODES[kind]
            if kind == "depolarizing":
                w = [np.sqrt(1 - p) ** 2] + [np.sqrt(p / 3) ** 2] * 3
            else:
                w = [np.sqrt(1 - p) ** 2, np.sqrt(p) ** 2]
            cdf = choice_cdf(w)
            thr = list(cdf[:-1]) + [2.0] * (3 - (len(cdf) - 1))
            codes4 = codes + [0] * (4 - len(codes))
            data = self.pool.add(thr + [float(c) for c in codes4])
            self._emit(KRAUS_PAULI, [qubit], data, draw=draw)
        elif kind == "amplitude_damping":
            data = self.pool.add([p, np.sqrt(1 - p), np.sqrt(p)])
            self._emit(KRAUS_AD, [qubit], data, draw=draw)
        else:
            ks = [np.asarray(k, dtype=np.complex128).reshape(2, 2) for k in kraus_ops]
            if not 1 <= len(ks) <= 8:
                raise ValueError("a Kraus set needs 1..8 operators")
            vals = [float(len(ks))]
            for k in ks:
                e = k.conj().T @ k
                vals += _cplx(k).tolist() + [e[0, 0].real, e[1, 1].real, e[0, 1].real, e[0, 1].imag]
            data = self.pool.add(vals)
            self._emit(KRAUS_GEN, [qubit], data, draw=draw)

    # -- snapshots (record_steps, simulator.py:70-71) -----------------------------------
    def snapshot(self):
        slot = self.n_snapshots
        self.n_snapshots += 1
        self.items.append([SNAPSHOT, [], slot, -1, -1, list(self.bit_of_axis)])
        return slot

    # -- streaming (state in HBM): passes of resident tiles -------------------------------
    def finish_stream(self, local_bits=None, low_bits=5, reorder=True):
        """Lower to a StreamPlan.  Between passes the state rests in PHYSICAL order (the order it was loaded
        in); with `reorder` the last pass stores it back in reference order (out of place when the
        reference's axis scramble left a non-trivial permutation)."""
        n = self.n
        m = min(n, MAX_LOCAL_BITS) if local_bits is None else int(local_bits)
        if not 1 <= m <= min(n, MAX_LOCAL_BITS):
            raise ValueError(f"local_bits {m} invalid for n = {n}")
        low = max(0, min(low_bits, m - 3))      # keep room for a 3-qubit gate on arbitrary bits
        for it in self.items:
            if it[0] in (SNAPSHOT, KRAUS_AD, KRAUS_GEN):
                raise NotImplementedError("snapshots and state-dependent Kraus draws need the whole state "
                                          "resident (n <= 16)")
        packed = plan_passes(self.items, n, m, low)
        final_perm_phys = [0] * n                  # physical bit -> reference-order bit
        for axis in range(n):
            final_perm_phys[self.bit_of_axis[axis]] = n - 1 - axis
        needs_reorder = reorder and any(final_perm_phys[b] != b for b in range(n))
        if not packed and needs_reorder:
            packed = [(list(range(m)), [])]
        if not packed:
            packed = [(list(range(m)), [])]
        cdata = self.pool.array()
        progs = []
        for k, (resident, chosen) in enumerate(packed):
            others = [b for b in range(n) if b not in resident]
            phys_of_slot = resident + others       # slot j (tile-local first, then the tile-number bits)
            slot_of = {b: j for j, b in enumerate(phys_of_slot)}
            ops = []
            for idx in chosen:
                kind, pbits, data, param, draw, _ = self.items[idx]
                sb = [slot_of[b] for b in pbits] + [0, 0, 0]
                ops.append((kind, sb[0], sb[1], sb[2], data, param, draw, 0))
            last = k == len(packed) - 1
            store = [final_perm_phys[b] for b in phys_of_slot] if (last and needs_reorder) else list(phys_of_slot)
            idata = list(phys_of_slot) + store
            arr = np.array(ops, dtype=OP_DTYPE) if ops else np.zeros(0, dtype=OP_DTYPE)
            progs.append(Program(n=n, m=m, ops=arr, cdata=cdata, idata=np.array(idata, dtype=np.int32),
                                 load_perm=0, store_perm=n, n_snapshots=0, n_draws=self.n_draws,
                                 n_params=self.n_params, normalize=False, n_gate_ops=0, n_kraus_ops=0,
                                 meta={"resident": resident}))
        return StreamPlan(n=n, m=m, passes=progs, final_out_of_place=needs_reorder, n_gate_ops=self.n_gate_ops,
                          n_draws=self.n_draws, n_params=self.n_params)

    # -- slot placement + emission ------------------------------------------------------
    def finish(self, local_bits=None, max_local_bits=MAX_LOCAL_BITS, multi_remap=True):
        """max_local_bits: 13 for complex128 tiles, 14 in the context's complex64 mode (twice the amplitudes per CTA).
        multi_remap: let one REMAP op exchange up to three (rank bit, local bit) pairs."""
        n = self.n
        if n > MAX_QUBITS:
            raise NotImplementedError(f"the resident executor holds at most {MAX_QUBITS} qubits, got {n}")
        m = min(n, max_local_bits) if local_bits is None else int(local_bits)
        if not (1 <= m <= min(n, max_local_bits)) or n - m > 3:
            raise ValueError(f"local_bits {m} invalid for n = {n}")
        g = n - m
        items = self.items
        # 1-qubit gates and Pauli / amplitude-damping draws only update the pending 2x2 of their slot,
        # which may be a cluster-rank slot; only ops that sweep the tile need their bits resident
        def needs_local(kind, pbits):
            return len(pbits) > 1 or kind == KRAUS_GEN
        uses = [[] for _ in range(n)]          # indices of the sweeping ops touching each physical bit
        for i, it in enumerate(items):
            if it[0] != SNAPSHOT and needs_local(it[0], it[1]):
                for b in it[1]:
                    uses[b].append(i)
        for it in items:
            if len(it[1]) > m:
                raise ValueError(f"a {len(it[1])}-qubit op does not fit {m} resident bits")
        # initial placement: the g bits first needed last live in the rank bits
        first_use = [(uses[b][0] if uses[b] else len(items) + 1 + b, b) for b in range(n)]
        glob = sorted(first_use, reverse=True)[:g]
        glob_bits = sorted(b for _, b in glob)
        slot_of = [0] * n
        s = 0
        for b in range(n):
            if b not in glob_bits:
                slot_of[b] = s
                s += 1
        for i, b in enumerate(glob_bits):
            slot_of[b] = m + i
        load_perm = [0] * n
        for b in range(n):
            load_perm[slot_of[b]] = b
        idata = list(load_perm)
        ops = []
        n_remaps = 0

        def next_use(b, i):
            u = uses[b]
            k = bisect.bisect_left(u, i)
            return u[k] if k < len(u) else 1 << 60

        def perm_for(bit_of_axis):
            perm = [0] * n
            for axis in range(n):
                perm[slot_of[bit_of_axis[axis]]] = n - 1 - axis
            return perm

        for i, (kind, pbits, data, param, draw, snap_axes) in enumerate(items):
            if kind == SNAPSHOT:
                off = len(idata)
                idata += perm_for(snap_axes)
                ops.append((SNAPSHOT, data, 0, 0, -1, -1, -1, off))
                continue
            if needs_local(kind, pbits) and any(slot_of[b] >= m for b in pbits):
                # One exchange op moves every rank-bit qubit this op needs into the tile and, while the data is moving
                # anyway, every other rank-bit qubit that is needed sooner than the local qubit it would displace
                # (k bits in one pass cost 1 - 2^-k tile volumes instead of k / 2).  Victims: farthest next use.
                pairs = []                                   # (incoming qubit, outgoing qubit)
                taken = set(pbits)
                def farthest():
                    cand = [c for c in range(n) if slot_of[c] < m and c not in taken]
                    return max(cand, key=lambda c: (next_use(c, i), -slot_of[c]))
                for b in pbits:
                    if slot_of[b] >= m:
                        v = farthest()
                        pairs.append((b, v))
                        taken.add(v)
                if multi_remap:
                    others = sorted((c for c in range(n) if slot_of[c] >= m and c not in taken),
                                    key=lambda c: next_use(c, i))
                    for x in others:
                        if len(pairs) >= 3 or not any(slot_of[c] < m and c not in taken for c in range(n)):
                            break
                        v = farthest()
                        if next_use(x, i) >= next_use(v, i):
                            break
                        pairs.append((x, v))
                        taken.update((x, v))
                aux = 0
                for j, (b, v) in enumerate(pairs[1:]):
                    aux |= ((slot_of[b] - m) | (slot_of[v] << 8)) << (16 * j)
                ops.append((REMAP, slot_of[pairs[0][0]] - m, slot_of[pairs[0][1]], len(pairs) - 1, -1, -1, -1, aux))
                for b, v in pairs:
                    slot_of[b], slot_of[v] = slot_of[v], slot_of[b]
                n_remaps += 1
            sb = [slot_of[b] for b in pbits] + [0, 0, 0]
            ops.append((kind, sb[0], sb[1], sb[2], data, param, draw, 0))
        store_off = len(idata)
        idata += perm_for(self.bit_of_axis)
        arr = np.array(ops, dtype=OP_DTYPE) if ops else np.zeros(0, dtype=OP_DTYPE)
        return Program(n=n, m=m, ops=arr, cdata=self.pool.array(), idata=np.array(idata, dtype=np.int32),
                       load_perm=0, store_perm=store_off, n_snapshots=self.n_snapshots, n_draws=self.n_draws,
                       n_params=self.n_params, normalize=self.normalize, n_gate_ops=self.n_gate_ops,
                       n_kraus_ops=self.n_kraus_ops, n_remaps=n_remaps)


@dataclass
class StreamPlan:
    """A program for a state that lives in HBM (n > 16, or any n with n - m > 3): a list of passes.
    Every pass but (possibly) the last runs in place; `final_out_of_place` says the last one stores through a
    different bit permutation (back to reference order) and therefore needs a second buffer."""
    n: int
    m: int
    passes: list                 # [Program]
    final_out_of_place: bool
    n_gate_ops: int = 0
    n_draws: int = 0
    n_params: int = 0


def plan_passes(items, n, m, low, frozen=()):
    """Greedy packing of ops into passes.  A pass keeps `m` physical bits resident: always the `low`
    lowest ones (contiguous 16 * 2^low-byte rows in HBM), never those in `frozen` (bits owned by another
    device), the rest by demand.  An op that does not fit blocks its qubits; later ops on blocked qubits wait
    for the next pass (ops on other qubits commute past them and are packed now).
    Returns [(sorted resident bits, [item indices])]."""
    remaining = list(range(len(items)))
    passes = []
    frozen = set(frozen)
    while remaining:
        resident = set(range(low))
        blocked = set()
        chosen, rest = [], []
        for idx in remaining:
            pbits = items[idx][1]
            if any(b in frozen for b in pbits):
                raise ValueError("op on a bit that is not addressable on this device")
            if any(b in blocked for b in pbits):
                blocked.update(pbits)
                rest.append(idx)
                continue
            need = [b for b in pbits if b not in resident]
            if len(resident) + len(need) <= m:
                resident.update(need)
                chosen.append(idx)
            else:
                blocked.update(pbits)
                rest.append(idx)
        if not chosen:
            raise ValueError(f"an op needs more than {m} resident bits")
        for b in range(n):                       # fill the tile with the lowest free bits
            if len(resident) >= m:
                break
            if b not in resident and b not in frozen:
                resident.add(b)
        passes.append((sorted(resident), chosen))
        remaining = rest
    return passes


def _kron_factors(mat, k):
    """Split a 2^k x 2^k Kronecker product of 2x2 blocks into its factors (first factor = MSB qubit);
    None if `mat` is not such a product (checked to 1e-13)."""
    factors = []
    rest = mat
    for _ in range(k - 1):
        d = rest.shape[0] // 2
        blocks = rest.reshape(2, d, 2, d).transpose(0, 2, 1, 3)          # [a, b] -> d x d block
        norms = np.array([[np.linalg.norm(blocks[a, b]) for b in range(2)] for a in range(2)])
        a0, b0 = np.unravel_index(np.argmax(norms), (2, 2))
        if norms[a0, b0] == 0:
            return None
        base = blocks[a0, b0]
        piv = np.unravel_index(np.argmax(np.abs(base)), base.shape)
        f = np.array([[blocks[a, b][piv] / base[piv] for b in range(2)] for a in range(2)])
        if np.max(np.abs(np.kron(f, base) - rest)) > 1e-13 * max(1.0, np.max(np.abs(rest))):
            return None
        factors.append(f)
        rest = base
    factors.append(rest)
    return factors
