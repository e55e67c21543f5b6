"""Host-side sharding logic on CPU: world_size-2 gloo processes reproduce the single-process streams,
child seeds and count merge.  The per-shot work is done by the oracle here (the GPU path is covered by
tests/test_gpu_*); what is under test is the partitioning / RNG positioning / collectives."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qsim_oracle as O
from qsb import distributed as D
from qsb.workloads import ghz


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 500, 4096):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_positioned_rng_equals_sequential_stream():
    full = np.random.default_rng(99).random(1000)
    for skip in (0, 1, 333, 999):
        assert np.array_equal(D.positioned_rng(99, skip).random(1000 - skip), full[skip:])
    g = np.random.default_rng(5)
    g.random(10)
    want = np.random.default_rng(5).random(40)[10:]
    assert np.array_equal(D.positioned_rng(g, 7).random(23), want[7:])
    assert np.array_equal(g.random(30), want)          # the caller's generator was not advanced


def test_child_seeds_match_reference_chain():
    assert D.child_seeds(42, 5) == O.trial_seeds(42, 5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shots, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, gates = 3, ghz(3)
    noise = {"global": [("depolarizing", 0.1)], "gate": {}}
    d = O.draw_count(n, gates, noise)
    lo, hi = D.shard_bounds(shots, world, rank)
    nrng = D.positioned_rng(7, lo * d)          # noise generator: d doubles per shot
    mrng = D.positioned_rng(42, lo)             # measurement generator: one double per shot
    idx = np.empty(hi - lo, dtype=np.int64)
    hist = torch.zeros(2 ** n, dtype=torch.float64)
    for i in range(hi - lo):
        psi = O.run_state(n, gates, None, noise, nrng.random(d))[0]
        idx[i] = O.measure_all_index(psi, mrng.random())
        hist += torch.from_numpy(np.abs(psi) ** 2)
    D.allreduce_sum_(hist)
    allidx = D.gather_concat(idx)
    if rank == 0:
        out_q.put((D.merge_counts_in_shot_order(allidx, n), hist.numpy().copy()))
    dist.destroy_process_group()


def test_world2_gloo_matches_single_process(golden):
    j, _ = golden
    shots = 200
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shots, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, hist = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # same counts, same insertion order as the reference's single loop (golden from the real reference)
    assert counts == j["ghz3"]["run_with_noise"]
    assert list(counts) == list(j["ghz3"]["run_with_noise"])
    assert abs(hist.sum() - shots) < 1e-9


def _exchange_worker(rank, world, port, n, out_q):
    """Sharded plan executed with NumPy per rank and a real gloo all_to_all_single for the exchange steps."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from qsb.bigstate import plan_distributed, exchange_rank_bits, Step
    from qsb.workloads import layered_circuit
    from test_bigstate import ordered, lower, replay_numpy
    g = world.bit_length() - 1
    L = n - g
    gl = ordered(n, layered_circuit(n, 5, 77))
    lw = lower(n, gl)
    steps, pos_of = plan_distributed(lw, g, local_bits=4)
    rng = np.random.default_rng(3)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    shard = psi[rank << L:(rank + 1) << L].copy()
    n_ex = 0
    for st in steps:
        if st.kind == "exchange":
            src = torch.from_numpy(shard.view(np.float64).copy())
            dst = torch.empty_like(src)
            exchange_rank_bits(src, dst)
            shard = dst.numpy().view(np.complex128).copy()
            n_ex += 1
        else:
            # one local pass: reuse the single-shard NumPy model (g = 0 view of this rank's shard)
            shard = replay_numpy([st], L, 0, shard, lw.pool.array())
    parts = [torch.zeros(2 * (1 << L), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(shard.view(np.float64).copy()))
    if rank == 0:
        full = np.concatenate([p.numpy().view(np.complex128) for p in parts])
        pos = [pos_of[lw.bit_of_axis[j]] for j in range(n)]
        got = np.ascontiguousarray(full.reshape([2] * n).transpose([n - 1 - pos[j] for j in range(n)])).reshape(-1)
        ref = psi
        for name, targets, params in gl:
            ref = O.apply_gate(ref, n, O.gate_matrix(name, params), targets)
        out_q.put((float(np.max(np.abs(got - ref))), n_ex))
    dist.destroy_process_group()


def test_world2_gloo_sharded_state_with_qubit_exchanges():
    """BASELINE config 5's exchange step on CPU: global qubit <-> local qubit swaps as all_to_all_single."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, n_ex = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert n_ex >= 1
    assert err < 1e-12


def _stream_exchange_worker(rank, world, port, n, out_q):
    """The TMA pipeline's plan (qsb/stream.py: host-fused block sweeps, reorder passes, exchanges) executed per rank with
    the NumPy model of a pass and a real gloo all_to_all_single for the exchange steps."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from qsb import stream as S
    from qsb.bigstate import exchange_rank_bits
    from qsb.workloads import layered_circuit
    from test_bigstate import ordered, lower
    from test_stream_plan import replay
    g = world.bit_length() - 1
    L = n - g
    gl = ordered(n, layered_circuit(n, 6, 78))
    lw = lower(n, gl)
    steps, pos_of, _ = S.plan(lw.items, lw.pool.array(), n, g, list(range(n)), local_bits=5, low_bits=2, box_bits=1)
    rng = np.random.default_rng(4)
    psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    psi /= np.linalg.norm(psi)
    shard = psi[rank << L:(rank + 1) << L].copy()
    kinds = ["exchange" if st.scatter else st.kind for st in steps]

    def a2a(shard):
        src = torch.from_numpy(shard.view(np.float64).copy())
        dst = torch.empty_like(src)
        exchange_rank_bits(src, dst)
        return dst.numpy().view(np.complex128).copy()

    for st in steps:
        if st.kind == "exchange":
            shard = a2a(shard)
        else:
            local = S.Step(st.kind, st.spass)                     # this rank's pass; the folded exchange is the all-to-all
            shard = replay([local], L, 0, shard, lw.pool.array())
            if st.scatter:
                shard = a2a(shard)
    parts = [torch.zeros(2 * (1 << L), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(shard.view(np.float64).copy()))
    if rank == 0:
        full = np.concatenate([p.numpy().view(np.complex128) for p in parts])
        pos = [pos_of[lw.bit_of_axis[j]] for j in range(n)]
        got = np.ascontiguousarray(full.reshape([2] * n).transpose([n - 1 - pos[j] for j in range(n)])).reshape(-1)
        ref = psi
        for name, targets, params in gl:
            ref = O.apply_gate(ref, n, O.gate_matrix(name, params), targets)
        out_q.put((float(np.max(np.abs(got - ref))), kinds.count("exchange"), kinds.count("reorder")))
    dist.destroy_process_group()


def test_world2_gloo_stream_plan_with_qubit_exchanges():
    """Config 5 on the TMA pipeline's planner, N = 2 on CPU: passes of block sweeps, reorder passes and exchanges."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_exchange_worker, args=(r, 2, port, 10, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, n_ex, n_re = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert n_ex >= 1
    assert err < 1e-12


def _qec_worker(rank, world, port, out_q):
    """threshold_sweep_sharded with the device stood in for by the oracle's cycle (the sharding, the seed chain and
    the gather are what is under test)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantum_sim.engine.qec import QECSimulator, SteaneCode

    def run_cycles(logicals, noise_type, p, seeds):
        rs = [O.qec_cycle("steane", l, noise_type, p, s) for l, s in zip(logicals, seeds)]
        return {"fidelity_after": np.array([r["fidelity_after"] for r in rs]), "z_exp": np.array([r["z_exp"] for r in rs]),
                "logical_error": np.array([r["logical_error"] for r in rs], dtype=bool)}

    sim = QECSimulator.__new__(QECSimulator)          # no device: only the sharded driver is exercised
    sim._code = None
    pts = sim.threshold_sweep_sharded([0.02, 0.1], n_trials=9, noise_type="depolarizing", seed=42, _run_cycles=run_cycles)
    if rank == 0:
        out_q.put([(pt.physical_rate, pt.logical_rate, pt.success_rate, pt.avg_fidelity, pt.logical_z_fidelity,
                    pt.decoder_success_rate, pt.projection_logical_rate) for pt in pts])
    dist.destroy_process_group()


def test_world2_gloo_sharded_threshold_sweep_is_bit_identical():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_qec_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = O.threshold_sweep("steane", [0.02, 0.1], 9, "depolarizing", 42)       # single loop, odd trial count
    for g, w in zip(got, want):
        assert g == (w["physical_rate"], w["logical_rate"], w["success_rate"], w["avg_fidelity"], w["logical_z_fidelity"],
                     w["decoder_success_rate"], w["projection_logical_rate"])


def _rwn_worker(rank, world, port, shots, out_q):
    """Simulator.run_with_noise_sharded (the product method) with the device leg stood in for by the oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise
    from quantum_sim.engine.simulator import Simulator
    n, gates = 3, ghz(3)
    noise = {"global": [("depolarizing", 0.1)], "gate": {}}
    qc = QuantumCircuit(n)
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    nm = NoiseModel()
    nm.add_global_noise(DepolarizingNoise(0.1))
    nm.set_seed(7)

    def indices(circuit, uniforms, measure_u):
        return [O.measure_all_index(O.run_state(n, gates, None, noise, uniforms[i])[0], measure_u[i])
                for i in range(len(measure_u))]
    indices.n_draws = O.draw_count(n, gates, noise)

    sim = Simulator(nm)
    res = sim.run_with_noise_sharded(qc, shots=shots, seed=42, _indices_fn=indices)
    # the model's generator ends where the single loop leaves it
    nxt = float(nm._rng.random())
    ref = np.random.default_rng(7)
    ref.random(shots * indices.n_draws)
    if rank == 0:
        out_q.put((res.measurement_counts, res.num_shots, nxt == float(ref.random())))
    dist.destroy_process_group()


def test_world2_gloo_run_with_noise_sharded_product_method(golden):
    j, _ = golden
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_rwn_worker, args=(r, 2, port, 200, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, num, rng_ok = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert counts == j["ghz3"]["run_with_noise"] and list(counts) == list(j["ghz3"]["run_with_noise"])
    assert num == 200 and rng_ok


def _costs_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantum_sim.engine.optimizer import batch_costs_sharded

    class Cfg:                                   # only what the sharded driver touches
        num_params = 3

    def fake(config, cost_fn, rows):             # stands in for the device batch: any row-wise function will do
        return np.array([cost_fn(r) for r in rows])

    vals = np.random.default_rng(5).uniform(-1, 1, (11, 3))
    got = batch_costs_sharded(Cfg(), lambda r: float(np.sum(np.cos(r))), vals, _batch_costs=fake)
    if rank == 1:
        out_q.put(got)
    dist.destroy_process_group()


def test_world2_gloo_batch_costs_sharded():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_costs_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    vals = np.random.default_rng(5).uniform(-1, 1, (11, 3))
    assert np.array_equal(got, np.array([float(np.sum(np.cos(r))) for r in vals]))      # every rank holds all costs


def _philox_sweep_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantum_sim.engine.qec import QECSimulator, SteaneCode
    out_q.put((rank, _philox_sweep(world)))
    dist.destroy_process_group()


def _philox_sweep(world):
    """threshold_sweep_philox with the device stood in for by a cheap deterministic function of the uniforms."""
    from quantum_sim.engine.qec import QECSimulator, SteaneCode

    def run_cycles(logicals, noise_type, p, seeds, uniforms=None):
        fired = (uniforms < p).sum(axis=1)
        return {"fidelity_after": np.where(fired <= 1, 1.0, 0.25), "z_exp": np.where(fired % 2 == 0, 1.0, -0.5),
                "logical_error": fired >= 2}

    sim = QECSimulator.__new__(QECSimulator)
    sim._code = SteaneCode.__new__(SteaneCode)
    pts = sim.threshold_sweep_philox([0.02, 0.2], n_trials=1000, noise_type="depolarizing", seed=9, batch=128,
                                     _run_cycles=run_cycles)
    return [(pt.physical_rate, pt.logical_rate, pt.avg_fidelity, pt.logical_z_fidelity, pt.decoder_success_rate) for pt in pts]


def test_world2_gloo_philox_threshold_sweep_is_independent_of_the_world_size():
    """Counter-based draws keyed by (seed, point, batch): two ranks sharing the batches give the single-process sums."""
    single = _philox_sweep(1)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_philox_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get() for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(2):
        for a, b in zip(got[r], single):
            assert a[0] == b[0] and all(abs(x - y) < 1e-12 for x, y in zip(a[1:], b[1:]))
    assert 0.0 < single[1][1] < 1.0
