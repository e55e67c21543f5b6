"""Statevectors beyond the resident executor's 16 qubits: one 2^n state in HBM, streamed tile by tile,
optionally sharded over the GPUs of one box (BASELINE.json config 5).

The reference caps StateVector at 16 qubits in its constructor only (state_vector.py:15-20);
`apply_gate` itself works for any n (verified to n = 26).  This module is that
"headless engine path": the same gate semantics (including the axis scramble of
state_vector.py:66-73, tracked as bookkeeping), executed as streamed passes of the tile executor
(csrc/qsb_exec.cuh, streaming mode).

Layout.  Memory index bit p of the (global) amplitude index is a POSITION.  Positions n-g .. n-1 are the
rank of the owning GPU (g = log2(world size)); the others address the local shard.  The circuit is
lowered against VIRTUAL bits (compiler.Lowering's "physical bits"); `pos_of[v]` says where virtual bit v
currently lives.  Three kinds of step:
  * pass      -- one launch over all tiles of the local shard; ops only touch local positions
  * reorder   -- an (empty) out-of-place pass that permutes local positions (used to bring the qubits that
                 are not needed for the longest time to the top local positions)
  * exchange  -- swap the g rank positions with the top g local positions: one all_to_all_single over
                 NCCL/NVLink of contiguous chunks (each GPU sends (1 - 2^-g) of its shard)
`torch` owns the device memory and the process group; every amplitude operation is a libqsb kernel.
"""

from __future__ import annotations

import os
import sys

import numpy as np

from . import capi
from . import stream
from .compiler import Lowering, Program, OP_DTYPE, MAX_LOCAL_BITS, SNAPSHOT, KRAUS_AD, KRAUS_GEN


class Step:
    __slots__ = ("kind", "prog", "out_of_place", "dprog")

    def __init__(self, kind, prog=None, out_of_place=False):
        self.kind, self.prog, self.out_of_place, self.dprog = kind, prog, out_of_place, None


def plan_distributed(lw: Lowering, g: int, local_bits=None, low_bits=4):
    """Lowering -> (steps, pos_of).  g = number of rank (global) positions; 0 = single device.

    Returns the step list and the final `pos_of` (virtual bit -> position)."""
    n = lw.n
    L = n - g                                   # local positions 0 .. L-1
    m = min(L, MAX_LOCAL_BITS) if local_bits is None else int(local_bits)
    if not 1 <= m <= min(L, MAX_LOCAL_BITS):
        raise ValueError(f"local_bits {m} invalid for {L} local bits")
    if g > 0 and L < 2 * g:
        raise ValueError("need at least as many local bits as rank bits to exchange them")
    low = max(0, min(low_bits, m - 3))
    items = lw.items
    for it in items:
        if it[0] in (SNAPSHOT, KRAUS_AD, KRAUS_GEN):
            raise NotImplementedError("snapshots and state-dependent Kraus draws need the whole state resident")
    cdata = lw.pool.array()
    pos_of = list(range(n))                     # virtual bit -> position
    remaining = list(range(len(items)))
    steps = []

    def make_pass(resident_pos, chosen, store_pos=None):
        """resident_pos: sorted local positions held in the tile; chosen: item indices (ops in positions).
        store_pos[j]: position the data at load slot j is stored to (None = in place)."""
        others = [p for p in range(L) if p not in resident_pos]
        load = list(resident_pos) + others       # slot j -> local position
        slot_of_pos = {p: j for j, p in enumerate(load)}
        ops = []
        for idx in chosen:
            kind, vbits, data, param, draw, _ = items[idx]
            sb = [slot_of_pos[pos_of[v]] for v in vbits] + [0, 0, 0]
            ops.append((kind, sb[0], sb[1], sb[2], data, param, draw, 0))
        store = list(load) if store_pos is None else [store_pos[p] for p in load]
        arr = np.array(ops, dtype=OP_DTYPE) if ops else np.zeros(0, dtype=OP_DTYPE)
        return Program(n=L, m=m, ops=arr, cdata=cdata, idata=np.array(load + store, dtype=np.int32), load_perm=0,
                       store_perm=L, n_snapshots=0, n_draws=lw.n_draws, n_params=lw.n_params, normalize=False)

    while remaining:
        # ---- pack one pass out of the ops whose qubits are all local right now
        resident = set(range(low))
        blocked = set()
        chosen, rest = [], []
        for idx in remaining:
            ps = [pos_of[v] for v in items[idx][1]]
            if any(p in blocked for p in ps) or any(p >= L for p in ps):
                blocked.update(ps)
                rest.append(idx)
                continue
            need = [p for p in ps if p not in resident]
            if len(resident) + len(need) <= m:
                resident.update(need)
                chosen.append(idx)
            else:
                blocked.update(ps)
                rest.append(idx)
        if chosen:
            for p in range(L):
                if len(resident) >= m:
                    break
                resident.add(p)
            steps.append(Step("pass", make_pass(sorted(resident), chosen)))
            remaining = rest
            continue
        if g == 0:
            raise ValueError(f"an op needs more than {m} resident bits")
        # ---- nothing runs without a qubit that lives in the rank bits: bring all g of them in.
        # Victims = the g local virtual bits whose next use is farthest away (Belady); a reorder pass moves them
        # to the top local positions, then one all-to-all swaps them with the rank positions.
        next_use = {}
        for order, idx in enumerate(remaining):
            for v in items[idx][1]:
                next_use.setdefault(v, order)
        local_v = [v for v in range(n) if pos_of[v] < L]
        victims = sorted(local_v, key=lambda v: (-next_use.get(v, 1 << 60), -pos_of[v]))[:g]
        top = list(range(L - g, L))
        if sorted(pos_of[v] for v in victims) != top:
            # permutation of local positions: victims -> top (ascending by current position), the rest keep order
            vic_pos = sorted(pos_of[v] for v in victims)
            keep = [p for p in range(L) if p not in vic_pos]
            new_of_old = {}
            for j, p in enumerate(keep):
                new_of_old[p] = j
            for j, p in enumerate(vic_pos):
                new_of_old[p] = L - g + j
            res = list(range(m))
            steps.append(Step("reorder", make_pass(res, [], store_pos=new_of_old), out_of_place=True))
            for v in range(n):
                if pos_of[v] < L:
                    pos_of[v] = new_of_old[pos_of[v]]
        steps.append(Step("exchange"))
        for v in range(n):
            p = pos_of[v]
            if p >= L:
                pos_of[v] = p - g
            elif p >= L - g:
                pos_of[v] = p + g
    return steps, pos_of


def exchange_rank_bits(src, dst, group=None):
    """Swap the g rank positions with the top g local positions of a sharded state: rank r's chunk c (the
    shard cut into world-size contiguous chunks) becomes rank c's chunk r.  One all_to_all_single; works on
    CUDA tensors over NCCL/NVLink and on CPU tensors over gloo (tests)."""
    import torch.distributed as dist
    dist.all_to_all_single(dst, src, group=group)
    return dst


class BigState:
    """One n-qubit complex128 state on this rank's GPU (a 2^(n-g) shard when the process group has 2^g ranks)."""

    def __init__(self, n, *, group=None, device=None, layout="reference", local_bits=None, distributed=True,
                 fuse_exchange=True, engine="tma", fuse_where="store"):
        """engine: "tma" = the TMA tile pipeline with host-fused sweeps (csrc/qsb_stream.cuh, qsb/stream.py);
        "executor" = round 1's path, one tile per CTA of the resident executor in its streaming mode (kept for A/B
        and as a second implementation the tests compare against).
        distributed=False keeps the whole state on this device even when a process group is initialised.
        fuse_exchange: fold every qubit exchange into the pass that follows it -- that pass LOADs its tiles straight
        from the peers' shards over NVLink peer mappings (torch symmetric memory) instead of waiting for an NCCL
        all-to-all into a second buffer; falls back to the all-to-all when the mappings cannot be set up."""
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.n = int(n)
        self.layout = layout
        self.group = group
        self.world = dist.get_world_size(group) if (distributed and dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.g = self.world.bit_length() - 1
        if 1 << self.g != self.world:
            raise ValueError("world size must be a power of two")
        self.L = self.n - self.g
        if self.n > 30 + self.g or self.L < 1:
            raise ValueError(f"num_qubits {n} not supported on {self.world} device(s)")
        self.local_bits = local_bits
        if engine not in ("tma", "executor"):
            raise ValueError("engine must be 'tma' or 'executor'")
        self.engine = engine
        # where a fused exchange happens (engine "tma"): "store" = in the store of the pass BEFORE it (peer-mapped TMA
        # stores; the reorder rides along, so an exchange costs no pass of its own), "load" = in the load of the pass
        # AFTER it (peer-mapped TMA loads behind a separate reorder pass)
        if fuse_where not in ("store", "load"):
            raise ValueError("fuse_where must be 'store' or 'load'")
        self.fuse_where = fuse_where
        self.fuse_store = fuse_exchange and fuse_where == "store"
        self.fused_exchanges = 0
        dev = capi.default_device() if device is None else device
        self.ctx = capi.get_context(dev)
        self.tdev = torch.device("cuda", dev)
        torch.cuda.set_device(self.tdev)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        dimL = 1 << self.L
        self._peer_tables = None
        if self.world > 1 and fuse_exchange:
            try:
                self._setup_symmetric(dimL, dist)
            except Exception as e:                      # no peer access on this box: keep the NCCL exchange
                self._peer_tables = None
                self._symm_error = repr(e)
        if self._peer_tables is None:
            self.buf = [torch.zeros(2 * dimL, dtype=torch.float64, device=self.tdev), None]
        self.cur = 0
        self._wrapped = [self.ctx.wrap(self.buf[0].data_ptr(), dimL * 16),
                         self.ctx.wrap(self.buf[1].data_ptr(), dimL * 16) if self.buf[1] is not None else None]
        self.pos_of = list(range(self.n))        # virtual bit v (= reference-order bit at rest) -> position
        self.bit_of_axis = [self.n - 1 - j for j in range(self.n)]
        if self.rank == 0:
            self.buf[0][0] = 1.0                 # |0...0>
        self.launches = 0

    def _setup_symmetric(self, dimL, dist):
        """Both shard buffers in symmetric memory; device tables of the peers' base pointers for each."""
        torch = self.torch
        import torch.distributed._symmetric_memory as symm
        grp = self.group if self.group is not None else dist.group.WORLD
        self.buf, tables = [], []
        for _ in range(2):
            t = symm.empty(2 * dimL, dtype=torch.float64, device=self.tdev)
            h = symm.rendezvous(t, grp.group_name)
            t.zero_()
            self.buf.append(t)
            ptrs = np.array([int(p) for p in h.buffer_ptrs], dtype=np.int64)
            assert len(ptrs) == self.world and int(ptrs[self.rank]) == t.data_ptr()
            tables.append(self.ctx.to_device(ptrs))
            self._peer_ptrs = getattr(self, "_peer_ptrs", []) + [ptrs.tolist()]
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._peer_tables = tables
        self._fence = torch.zeros(1, dtype=torch.float32, device=self.tdev)

    def _rank_fence(self):
        """Stream-ordered barrier over the ranks (a 4-byte all-reduce on the stream the passes run on)."""
        import torch.distributed as dist
        dist.all_reduce(self._fence, group=self.group)

    # -- buffers ------------------------------------------------------------------------------
    def _other(self):
        o = 1 - self.cur
        if self.buf[o] is None:
            dimL = 1 << self.L
            self.buf[o] = self.torch.empty(2 * dimL, dtype=self.torch.float64, device=self.tdev)
            self._wrapped[o] = self.ctx.wrap(self.buf[o].data_ptr(), dimL * 16)
        return o

    # -- execution ------------------------------------------------------------------------------
    def lowering(self):
        """A Lowering whose virtual bits continue this state's axis bookkeeping."""
        lw = Lowering(self.n, layout=self.layout)
        lw.bit_of_axis = list(self.bit_of_axis)
        return lw

    # -- the TMA tile pipeline ------------------------------------------------------------------------------------
    def compile(self, lw: Lowering, *, params=None, uniforms=None, seed=0):
        """Plan `lw` against the state's CURRENT layout and upload every pass: (steps, moved, bit_of_axis).  The result
        can be executed any number of times while the layout is the one it was compiled for (always true on one
        device in the textbook layout; after an exchange the same steps still run, on relabelled qubits)."""
        rel = _Relabel(lw, self.pos_of)
        cdata = lw.pool.array()
        steps, moved, _ = stream.plan(rel.items, cdata, self.n, self.g, list(range(self.n)), local_bits=self.local_bits,
                                      params=params, uniforms=uniforms, seed=seed, fuse_store=self.fuse_where != "load")
        for st in steps:
            if st.spass is not None:
                st.handle = self.ctx.stream_pass(st.spass, cdata)
        return steps, moved, list(lw.bit_of_axis)

    def execute(self, compiled, sync=True):
        """Run compiled steps: launches are asynchronous on the context's stream, one host sync at the end."""
        steps, moved, bit_of_axis = compiled
        pending_exchange = False
        trace = os.environ.get("QSB_TRACE")
        for k, st in enumerate(steps):
            if trace:                                # developer aid: which step a rank is in when something goes wrong
                self.ctx.sync()
                print(f"[rank {self.rank}] step {k}/{len(steps)} {st.kind} cur={self.cur}", file=sys.stderr, flush=True)
            if st.kind == "exchange":
                if self._peer_tables is not None and not pending_exchange:
                    pending_exchange = True          # folded into the LOAD of the next pass
                else:
                    self._exchange()
                continue
            if st.scatter:
                # the exchange (and the reorder that precedes it) rides in this pass's STORE: every TMA box goes
                # straight into the shard of the peer it belongs to, posted writes over NVLink under the sweeps of the
                # following tiles.  Destination = everybody's idle buffer (the fence after the previous exchange made sure
                # nobody still reads it); one fence afterwards: all boxes have landed before anyone sweeps them.
                shift = self.L - self.g
                o = self._other()
                if self._peer_tables is not None and self.fuse_store:
                    st.handle.run_scatter(self._wrapped[self.cur], self._peer_ptrs[o], shift, self.rank << shift)
                    self._rank_fence()
                    self.cur = o
                    self.fused_exchanges += 1
                else:                                # no peer mappings: the same pass out of place, then the all-to-all
                    st.handle.run(self._wrapped[self.cur], self._wrapped[o])
                    self.cur = o
                    self._exchange()
                self.launches += 1
                continue
            if pending_exchange:
                pending_exchange = False
                shift = self.L - self.g
                if all(p < shift for p in st.spass.positions[st.spass.l:st.spass.l + st.spass.e]):
                    # every rank has finished writing its shard (fence), this pass then reads its tiles straight from
                    # the peers' shards -- element s of my post-exchange shard = peer s >> shift, offset
                    # (s & mask) | rank << shift -- and stores into my other buffer; the second fence keeps anyone
                    # from overwriting a shard a peer is still reading
                    o = self._other()
                    self._rank_fence()
                    st.handle.run_peers(self._peer_ptrs[self.cur], shift, self.rank << shift, self._wrapped[o])
                    self._rank_fence()
                    self.cur = o
                    self.launches += 1
                    self.fused_exchanges += 1
                    continue
                self._exchange()                     # a TMA box dimension sits on a peer-selecting bit: plain all-to-all
            if st.spass.in_place:
                st.handle.run(self._wrapped[self.cur])
            else:
                o = self._other()
                st.handle.run(self._wrapped[self.cur], self._wrapped[o])
                self.cur = o
            self.launches += 1
        if pending_exchange:                         # an exchange with no pass after it
            self._exchange()
        self.pos_of = [moved[p] for p in self.pos_of]
        self.bit_of_axis = list(bit_of_axis)
        if sync:
            self.ctx.sync()

    def run(self, lw: Lowering, *, params=None, uniforms=None, seed=0):
        """Apply everything recorded in `lw` (created by self.lowering())."""
        if self.engine == "tma":
            self.execute(self.compile(lw, params=params, uniforms=uniforms, seed=seed))
            return
        steps, moved = plan_distributed(_Relabel(lw, self.pos_of), self.g, self.local_bits)
        kw = {}
        if params is not None:
            p = np.ascontiguousarray(params, dtype=np.float64).reshape(1, -1)
            kw.update(params=self.ctx.to_device(p), params_stride=p.shape[1])
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(1, -1)
            kw.update(uniforms=self.ctx.to_device(u), uniforms_stride=u.shape[1])
        pending_exchange = False
        for st in steps:
            if st.kind == "exchange":
                if self._peer_tables is not None and not pending_exchange:
                    pending_exchange = True          # folded into the LOAD of the next pass
                else:
                    self._exchange()
                continue
            dp = self.ctx.program(st.prog)
            if pending_exchange:
                # every rank has finished writing its shard; then this pass reads the peers' shards directly
                # (element s of my post-exchange shard = peer s >> (L - g), offset (s & mask) | rank << (L - g))
                # and stores into my other buffer; nobody may overwrite a shard that a peer is still reading
                pending_exchange = False
                o = self._other()
                self._rank_fence()
                self.ctx.run(dp, 1, states=self._wrapped[self.cur], load=True, store=True, states_out=self._wrapped[o],
                             peer_table=self._peer_tables[self.cur], peer_shift=self.L - self.g,
                             peer_rank_or=self.rank << (self.L - self.g), seed=seed, async_=True, **kw)
                self._rank_fence()
                self.cur = o
                self.launches += 1
                self.fused_exchanges += 1
                self.ctx.sync()
                continue
            if st.out_of_place:
                o = self._other()
                self.ctx.run(dp, 1, states=self._wrapped[self.cur], load=True, store=True, states_out=self._wrapped[o],
                             seed=seed, async_=True, **kw)
                self.cur = o
            else:
                self.ctx.run(dp, 1, states=self._wrapped[self.cur], load=True, store=True, seed=seed, async_=True, **kw)
            self.launches += 1
            self.ctx.sync()                        # the program object is freed when `dp` goes out of scope
        if pending_exchange:                         # an exchange with no pass after it
            self._exchange()
        # bookkeeping: `moved[p]` = where the data that sat at position p when the program started is now
        self.pos_of = [moved[p] for p in self.pos_of]
        self.bit_of_axis = list(lw.bit_of_axis)

    def _exchange(self):
        """Swap the g rank positions with the top g local positions (one all-to-all of contiguous chunks)."""
        o = self._other()
        exchange_rank_bits(self.buf[self.cur], self.buf[o], self.group)
        self.cur = o

    # -- circuits -------------------------------------------------------------------------------
    def apply_gates(self, gates, registry=None):
        """gates: iterable of (name, targets, params[, column]) in execution order (already ordered the way
        QuantumCircuit.get_ordered_gates orders them)."""
        if registry is None:
            from quantum_sim.engine.gate_registry import GateRegistry
            registry = GateRegistry.instance()
        lw = self.lowering()
        for g in gates:
            gd = registry.get(g[0])
            if gd.gate_type.value in ("measurement", "barrier"):
                continue
            lw.gate(g[0], list(g[1]), list(g[2]), gd.matrix_func,
                    builtin=registry.is_builtin(g[0]) if hasattr(registry, "is_builtin") else True)
        self.run(lw)
        return lw.n_gate_ops

    def apply_circuit(self, circuit, registry=None):
        gates = [(g.gate_name, g.target_qubits, g.params) for col in circuit.get_ordered_gates() for g in col]
        return self.apply_gates(gates, registry)

    # -- reductions -----------------------------------------------------------------------------
    def norm2(self):
        """sum |a|^2 over the whole (global) state."""
        torch = self.torch
        out = self.ctx.alloc(16)
        self.ctx.overlap(self.L, self._wrapped[self.cur], 0, self._wrapped[self.cur], 0, 1, 1, out)
        v = out.download(np.complex128, (1,))[0].real
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([v], dtype=torch.float64, device=self.tdev)
            dist.all_reduce(t, group=self.group)
            v = float(t.item())
        return float(v)

    def local_shard(self):
        """This rank's shard as a NumPy array in MEMORY order (see index_map for the meaning of the index bits)."""
        self.ctx.sync()
        return self.buf[self.cur].cpu().numpy().view(np.complex128)

    def amplitude_position_of_axis(self):
        """position (bit of the global memory index) that holds reference axis j, for j = 0..n-1."""
        return [self.pos_of[self.bit_of_axis[j]] for j in range(self.n)]

    def to_reference_order(self, shards):
        """Host-side assembly (tests / inspection): shards[r] = rank r's local_shard() -> the state in the
        reference's index order (qubit 0 = most significant bit)."""
        n = self.n
        full = np.concatenate(shards) if len(shards) > 1 else shards[0]
        # memory index bit p = numpy axis n-1-p; reference axis j lives at position pos[j]
        pos = self.amplitude_position_of_axis()
        t = full.reshape([2] * n).transpose([n - 1 - pos[j] for j in range(n)])
        return np.ascontiguousarray(t).reshape(-1)


class _Relabel:
    """View of a Lowering whose virtual bits are mapped through a starting `pos_of` (the state's layout when
    the program starts); plan_distributed then tracks positions from there."""

    def __init__(self, lw, pos_of):
        self.n = lw.n
        self.items = [[it[0], [pos_of[v] for v in it[1]], it[2], it[3], it[4], it[5]] for it in lw.items]
        self.pool = lw.pool
        self.n_draws, self.n_params = lw.n_draws, lw.n_params
