"""bench.py -- headline benchmark of the B200 statevector backend.

Metric (BASELINE.json): 16-qubit noisy trajectories/s (and gate-applications/s for the noiseless
parameter batch of configs[1] as a secondary figure), reported with the fraction of the measured HBM
roofline for the path's ALGORITHMIC bytes and the CPU oracle timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--traj T] [--impl ours|reference]

One step = one pass of the hot path over one batch: T noisy trajectories of
layered_circuit(16, 64, 2026) (683 gates, 2048 Kraus draws each: depolarizing 0.01 + amplitude damping
0.02 after every gate) -> final states in HBM -> one measure_all sample per trajectory -> probability
histogram (summed over ranks with one NCCL all-reduce when N > 1).  Trajectories shard across ranks
(weak scaling: T per GPU).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "quantum-simulator_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_QUBITS, DEPTH, CIRCUIT_SEED = 16, 64, 2026
METRIC = "noisy_trajectories_per_sec_16q"
WORKLOAD = "layered_circuit(16,64,2026) + depolarizing(0.01) + amplitude_damping(0.02), reference-draw mode"


def algorithmic_bytes_per_trajectory(n, gates, k_pauli, k_ad):
    """(G + K_pauli + 1.5 K_ad) * B_sweep + final read, B_sweep = 2 * 2^n * 16 B."""
    sweep = 2 * (2 ** n) * 16
    return (gates + k_pauli + 1.5 * k_ad) * sweep + 16 * 2 ** n


# DRAM bytes one trajectory really moves (ncu dram__bytes_read.sum + dram__bytes_write.sum of the trajectory
# kernel divided by its trajectories: profiles/r01g_traj_kernel_ncu_full_traj480.csv, 468.3 MB / 480)
MEASURED_DRAM_BYTES_PER_TRAJECTORY = (11614464 + 456644096) / 480
# shared-memory wavefronts (128 B each) one trajectory costs, same capture
# (l1tex__data_pipe_lsu_wavefronts_mem_shared.sum = 3 367 032 907 per 480 trajectories)
MEASURED_SMEM_WAVEFRONTS_PER_TRAJECTORY = 3367032907 / 480


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------ reference (CPU) arm
# The reference is pure Python + NumPy.  tools/make_baseline_ref.py copies it UNMODIFIED into the git-ignored
# baseline/_ref/ (it travels to the GPU box with the snapshot); these legs import THAT package -- the reference's
# own Simulator.run_with_noise (simulator.py:116-153 -> state_vector.py:41-74, noise.py:224-260) -- in worker
# processes that never import this repo's engine.  If baseline/_ref is missing the oracle port stands in (kind "port").
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "quantum_sim", "engine", "simulator.py"))


def _limit_blas(threads):
    if threads:
        os.environ["OPENBLAS_NUM_THREADS"] = str(threads)
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=threads)
        except Exception:
            pass


def _reference_prefix_worker(args):
    """One noisy trajectory of the first `cols` columns of the headline circuit on the REFERENCE engine:
    Simulator(noise_model).run_with_noise(circuit, shots=1).  Returns seconds."""
    cols, seed, blas_threads = args
    _limit_blas(blas_threads)
    if REF_ROOT not in sys.path or sys.path.index(REF_ROOT) != 0:
        sys.path.insert(0, REF_ROOT)
    import quantum_sim
    assert os.path.abspath(quantum_sim.__file__).startswith(REF_ROOT), quantum_sim.__file__
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
    from quantum_sim.engine.simulator import Simulator
    from qsb.workloads import layered_circuit
    qc = QuantumCircuit(N_QUBITS)
    for name, targets, params, col in layered_circuit(N_QUBITS, DEPTH, CIRCUIT_SEED):
        if col < cols:
            qc.add_gate(GateInstance(name, list(targets), list(params), col))
    nm = NoiseModel()
    nm.add_global_noise(DepolarizingNoise(0.01))
    nm.add_global_noise(AmplitudeDampingNoise(0.02))
    nm.set_seed(seed)
    sim = Simulator(nm)
    t0 = time.perf_counter()
    res = sim.run_with_noise(qc, shots=1, seed=seed)
    dt = time.perf_counter() - t0
    assert sum(res.measurement_counts.values()) == 1
    return dt


def _oracle_prefix_worker(args):
    cols, seed, blas_threads = args
    _limit_blas(blas_threads)
    from oracle import qsim_oracle as O
    from qsb.workloads import layered_circuit, config3_noise
    gates = [g for g in layered_circuit(N_QUBITS, DEPTH, CIRCUIT_SEED) if g[3] < cols]
    noise = config3_noise()
    d = O.draw_count(N_QUBITS, gates, noise)
    t0 = time.perf_counter()
    O.run_state(N_QUBITS, gates, None, noise, np.random.default_rng(seed).random(d))
    return time.perf_counter() - t0


def _cpu_worker():
    return (_reference_prefix_worker, "reference") if reference_available() else (_oracle_prefix_worker, "port")


def cpu_baseline_leg():
    """`bench.py --impl reference --cpu-sample` (a fresh process, started by the GPU arm on rank 0 at N = 1):
    ONE FULL 16-qubit noisy trajectory (all 64 columns) on one core with one BLAS thread, plus a quarter
    trajectory with NumPy's default BLAS threading.  Prints one JSON object."""
    worker, kind = _cpu_worker()
    worker((2, 1, 1))                                     # warm-up: imports, first-call costs
    full = worker((DEPTH, 1001, 1))
    try:                                                  # default threading: lift the limit again
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    os.environ.pop("OPENBLAS_NUM_THREADS", None)
    quarter = worker((DEPTH // 4, 1002, 0)) * 4
    what = ("the UNMODIFIED reference (baseline/_ref): Simulator.run_with_noise(circuit, shots=1)" if kind == "reference"
            else "oracle/qsim_oracle.py (baseline/_ref missing)")
    print(json.dumps({"value": 1.0 / full, "unit": "trajectories/s", "cores": 1, "kind": kind,
                      "sample": f"1 full trajectory of the headline workload (64 columns, 683 gates, 2048 Kraus draws) "
                                f"on {what}, 1 process, OPENBLAS_NUM_THREADS=1: {full:.2f} s",
                      "value_default_blas_threads": 1.0 / quarter,
                      "sample_default_blas_threads": f"first 16/64 columns x 4, 1 process, default BLAS threading "
                                                     f"({os.cpu_count()} host cores): {quarter:.2f} s per trajectory",
                      "host_cores": os.cpu_count()}), flush=True)


def cpu_baseline_sample():
    """Run cpu_baseline_leg() in a process of its own (this one has the CUDA engine imported as `quantum_sim`)."""
    env = {k: v for k, v in os.environ.items() if k not in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS")}
    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-sample"],
                         capture_output=True, text=True, env=env, timeout=600)
    if res.returncode != 0:
        return {"error": res.stderr[-400:]}
    return json.loads(res.stdout.strip().splitlines()[-1])


def bench_config(T, world):
    """The `config` object -- identical in both arms."""
    return {"workload": WORKLOAD, "n_qubits": N_QUBITS, "gates": 683, "kraus_draws": 2048,
            "trajectories_per_gpu_per_step": T, "parallelism": f"trajectory shards x{world}",
            "l2": "256 MiB flush between steps; per-step working set 4 GiB > L2"}


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores -- one worker process
    per core (the engine's Python layer is single-threaded), one BLAS thread each.  A step = every worker runs the
    first 16 of the 64 columns of one trajectory through Simulator.run_with_noise; trajectories/s scaled by 4
    (every column has the same op mix; cpu_baseline times one full trajectory to back that)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.cpu_sample:
        cpu_baseline_leg()
        return
    import multiprocessing as mp
    worker, kind = _cpu_worker()
    cols = 16
    workers = max(1, min(os.cpu_count() or 1, 64))
    pool = mp.get_context("spawn").Pool(workers)
    pool.map(worker, [(1, w, 1) for w in range(workers)])            # imports in every worker, untimed
    step_times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        pool.map(worker, [(cols, 10_000 + s * workers + w, 1) for w in range(workers)], chunksize=1)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            step_times.append(dt)
    pool.close()
    ms = 1e3 * sum(step_times) / len(step_times)
    value = workers / (ms * 1e-3 * DEPTH / cols)
    what = ("unmodified reference engine from baseline/_ref (Simulator.run_with_noise, shots=1)" if kind == "reference"
            else "oracle port (baseline/_ref missing)")
    sample = (f"each step: {workers} worker processes (1 BLAS thread each) x first {cols}/{DEPTH} columns of one "
              f"trajectory on the {what}, scaled by {DEPTH // cols}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": bench_config(args.traj, args.gpus),
            "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": workers, "kind": kind,
                             "sample": sample, "host_cores": os.cpu_count()},
            "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from qsb import capi
    from qsb.lowering import lower_circuit
    from qsb.workloads import layered_circuit, config3_noise, to_gate_instances
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.gate_registry import GateRegistry
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
    from quantum_sim.engine.simulator import Simulator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["QSB_DEVICE"] = str(local)
    ctx = capi.get_context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)          # our kernels, torch events and NCCL order on one stream

    n, T = N_QUBITS, args.traj
    dim = 2 ** n
    gates = layered_circuit(n, DEPTH, CIRCUIT_SEED)
    qc = QuantumCircuit(n)
    for g in to_gate_instances(gates, GateInstance):
        qc.add_gate(g)
    nm = NoiseModel()
    nm.add_global_noise(DepolarizingNoise(0.01))
    nm.add_global_noise(AmplitudeDampingNoise(0.02))
    nm.set_seed(12345 + rank)
    sim = Simulator(nm)
    dp, _ = sim._program(qc)
    prog = dp.prog
    D = prog.n_draws
    k_ad = D // 2
    alg_bytes = algorithmic_bytes_per_trajectory(n, prog.n_gate_ops, D - k_ad, k_ad)

    # ---- resident inputs / outputs (torch owns the memory, libqsb wraps the pointers)
    t_states = torch.empty(T * dim * 2, dtype=torch.float64, device="cuda")
    t_uniforms = torch.from_numpy(np.random.default_rng(777 + rank).random((T, D))).cuda()
    t_sample_u = torch.from_numpy(np.random.default_rng(888 + rank).random(T)).cuda()
    t_idx = torch.empty(T, dtype=torch.int64, device="cuda")
    t_hist = torch.zeros(dim, dtype=torch.float64, device="cuda")
    t_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    b_states = ctx.wrap(t_states.data_ptr(), T * dim * 16)
    b_uniforms = ctx.wrap(t_uniforms.data_ptr(), T * D * 8)
    b_sample_u = ctx.wrap(t_sample_u.data_ptr(), T * 8)
    b_idx = ctx.wrap(t_idx.data_ptr(), T * 8)
    b_hist = ctx.wrap(t_hist.data_ptr(), dim * 8)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(timers=None):
        t_flush.zero_()                          # evict L2 between steps (256 MiB > 126 MB), not timed
        t_hist.zero_()
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(stream)
        ctx.run(dp, T, states=b_states, uniforms=b_uniforms, uniforms_stride=D, traj_offset=rank * T, async_=True)
        e1.record(stream)
        ctx.lib.qsb_sample_index(ctx.handle, n, b_states.handle, 0, T, b_sample_u.handle, b_idx.handle)
        ctx.lib.qsb_probabilities_sum(ctx.handle, n, b_states.handle, 0, T, b_hist.handle)
        if world > 1:
            dist.all_reduce(t_hist)
        e2.record(stream)
        if timers is not None:
            timers.append((e0, e1, e2))

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    fence()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    timers = []
    fence()
    for _ in range(args.steps):
        step(timers)
    fence()
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    step_ms = [e0.elapsed_time(e2) for e0, e1, e2 in timers]
    kern_ms = [e0.elapsed_time(e1) for e0, e1, e2 in timers]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * T * args.steps / (total_ms * 1e-3)
    hist_total = float(t_hist.sum().item())

    # ---- end to end through the engine API with host buffers: Simulator.run_with_noise(circuit, shots)
    e2e_T = args.e2e_traj
    sim.run_with_noise(qc, shots=e2e_T, seed=1)      # warm-up: program cached, staging pinned, pool blocks mapped
    fence()
    t0 = time.perf_counter()
    for s in range(args.e2e_steps):
        res = sim.run_with_noise(qc, shots=e2e_T, seed=100 + s)
    torch.cuda.synchronize()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_T * args.e2e_steps / float(e2e_dt.item())
    assert sum(res.measurement_counts.values()) == e2e_T

    extras = None
    if rank == 0 and not args.no_extras:
        extras = measure_other_paths(ctx, sim, qc)
    if rank == 0:
        peak, peak_src = measured_peak()
        kms = float(np.mean(kern_ms))
        achieved = alg_bytes * T / (kms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": bench_config(T, world),
            "check": {"histogram_total": hist_total, "gates": prog.n_gate_ops, "kraus_draws": D},
            "roofline": {"bound": "hbm", "kernel": f"qsb_traj_kernel<{1 << (prog.n - prog.m)}>",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": MEASURED_DRAM_BYTES_PER_TRAJECTORY * T,
                         "traffic_source": "ncu dram bytes per trajectory (profiles/r01g_traj_kernel_ncu_full_traj480.csv) x trajectories per launch",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes * T, "kernel_ms": kms,
                         "note": "trajectories are resident in cluster shared memory; algorithmic bytes count "
                                 "every gate/Kraus sweep (SURVEY 8d), real DRAM traffic is ~1 MiB/trajectory "
                                 "(see profiles/), so frac > 1 is expected; `onchip` is the physical bound",
                         # where the kernel really stands: shared-memory bytes per SM-clock against 128 B/clk/SM
                         "onchip": (lambda clk_hz, sms: {
                             "bound": "shared memory", "unit": "B/clk/SM", "peak": 128.0,
                             "achieved": MEASURED_SMEM_WAVEFRONTS_PER_TRAJECTORY * 128.0 * T / (kms * 1e-3 * clk_hz * sms),
                             "frac": MEASURED_SMEM_WAVEFRONTS_PER_TRAJECTORY * T / (kms * 1e-3 * clk_hz * sms),
                             "sms_holding_clusters": sms, "sm_mhz": clk_hz / 1e6,
                             "source": "ncu shared-memory wavefronts per trajectory (profiles/r01g_*) x trajectories / "
                                       "(kernel time x SM clock x SMs that can hold clusters of 8)"})(
                             float(clocks.get("sm_mhz") or 1965.0) * 1e6, 120)},
            "e2e": {"value": e2e_value, "unit": "trajectories/s",
                    "h2d_bytes_per_step": e2e_T * (D + 1) * 8, "d2h_bytes_per_step": e2e_T * 8,
                    "api": "quantum_sim.engine.simulator.Simulator.run_with_noise", "shots": e2e_T,
                    "steps": args.e2e_steps},
            "gpu_launches": launches, "clocks": clocks,
        }
        if extras is not None:
            line["other_paths"] = extras
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_other_paths(ctx, sim, qc16):
    """Short, separately timed runs of the other BASELINE configs (not part of the headline timing):
    device time by CUDA events on the ctx stream, inputs resident."""
    import torch
    from qsb.workloads import layered_circuit, to_gate_instances
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
    from quantum_sim.engine.optimizer import ParameterizedCircuitConfig
    from quantum_sim.engine.qec import QECSimulator, SteaneCode
    from quantum_sim.engine.simulator import Simulator
    out = {}
    # config 2: 16-qubit layered circuit, 4096 parameter sets in one launch (noiseless) -> gate-apps/s
    cfg = ParameterizedCircuitConfig.auto_detect(qc16)
    vals = np.random.default_rng(2027).uniform(-np.pi, np.pi, (4096, cfg.num_params))
    cfg.run_batch(vals)                      # warm-up at the measured batch size (program cached, pool blocks mapped)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cfg.run_batch(vals)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gates = len(qc16.gates)
    out["config2_param_batch"] = {"param_sets": 4096, "gates": gates, "seconds": dt, "gate_apps_per_s": 4096 * gates / dt,
                                  "algorithmic_GBps": 4096 * gates * 2 * 16 * 2 ** 16 / dt / 1e9,
                                  "includes": "h2d of the 4096x473 parameter matrix"}
    # config 3: 12 qubits, depolarizing + amplitude damping, 500 trajectories -> rho (DMMA) + all-pairs MI per layer
    n = 12
    qc = QuantumCircuit(n)
    for g in to_gate_instances(layered_circuit(n, 16, 2026), GateInstance):
        qc.add_gate(g)
    nm = NoiseModel()
    nm.add_global_noise(DepolarizingNoise(0.01))
    nm.add_global_noise(AmplitudeDampingNoise(0.02))
    s12 = Simulator(nm)
    s12.ensemble_density_matrix(qc, 8, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rho = s12.ensemble_density_matrix(qc, 500, seed=42)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["config3_ensemble_rho"] = {"qubits": n, "trials": 500, "seconds": dt, "trace": float(np.real(np.trace(rho))),
                                   "purity": float(np.real(np.sum(rho * rho.T))),
                                   "includes": "host draws, 500 trajectories, DMMA rho, 256 MiB d2h"}
    # the rho accumulation alone (FP64 tensor pipe): rho += (1/N) sum_t psi_t psi_t^dagger over N resident states
    try:
        N3 = 2048
        st3 = ctx.to_device((np.random.default_rng(9).normal(size=(N3, 2 ** n, 2)).view(np.complex128).reshape(N3, 2 ** n)
                             / np.sqrt(2.0 ** (n + 1))))
        rho3 = ctx.alloc(16 << (2 * n)).zero()
        ctx.rho_accumulate(n, st3, 0, N3, 1.0 / N3, rho3)
        ctx.sync()
        ctx.timer_start()
        ctx.rho_accumulate(n, st3, 0, N3, 1.0 / N3, rho3)
        ms3 = ctx.timer_stop()
        full = 8.0 * 4.0 ** n * N3 / (ms3 * 1e-3) / 1e12
        out["config3_rho_kernel"] = {"qubits": n, "states": N3, "ms": ms3, "tflops_full_matrix_convention": full,
                                     "tflops_executed": full / 2, "fp64_spec_tflops": 37.0, "frac_of_spec": full / 2 / 37.0,
                                     "note": "Hermitian half computed on DMMA (mma.sync m8n8k4 f64); ncu: DMMA sub-pipe 88 % "
                                             "active (profiles/r01d_rho_kernel_ncu_full.csv)"}
        del st3, rho3
    except Exception as e:
        out["config3_rho_kernel"] = {"error": repr(e)}
    # config 3, second half: all 66 pair mutual informations of every per-column snapshot of every trajectory
    from quantum_sim.engine.analysis import all_pairs_mutual_information_device
    dp, _ = s12._program(qc, record_steps=True)
    T3, ns, dim3 = 500, dp.prog.n_snapshots, 1 << n
    u3 = ctx.to_device(np.random.default_rng(3).random((T3, dp.prog.n_draws)))
    snaps = ctx.alloc(T3 * ns * dim3 * 16)
    ctx.run(dp, 8, uniforms=u3, uniforms_stride=dp.prog.n_draws, snapshots=snaps, store=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.run(dp, T3, uniforms=u3, uniforms_stride=dp.prog.n_draws, snapshots=snaps, store=False)
    mi = all_pairs_mutual_information_device(n, snaps, 0, T3 * ns)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["config3_layer_mi"] = {"qubits": n, "trials": T3, "layers": ns, "pairs": int(mi.shape[1]), "seconds": dt,
                               "state_layers_per_s": T3 * ns / dt, "mean_mi_bits": float(mi.mean()),
                               "includes": "trajectories with per-column snapshots, RDMs, Jacobi entropies, d2h of MI"}
    del snaps
    # config 4: Steane [[7,1,3]] cycles (13 qubits), batched
    qs = QECSimulator(SteaneCode())
    seeds = list(range(1000, 1000 + 2048))
    qs.run_cycles([t % 2 for t in range(64)], "depolarizing", 0.05, seeds[:64])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    qs.run_cycles([t % 2 for t in range(2048)], "depolarizing", 0.05, seeds)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["config4_steane_cycles"] = {"cycles": 2048, "seconds": dt, "cycles_per_s": 2048 / dt,
                                    "includes": "reference streams (one default_rng per trial), decode table, d2h of syndromes"}
    # throughput mode of config 4: one vectorised generator per sweep point instead of one default_rng per trial
    Tq = 16384
    uq = np.random.default_rng(77).random((Tq, 7))
    qs.run_cycles([t % 2 for t in range(Tq)], "depolarizing", 0.05, None, uniforms=uq)    # warm-up at this batch size
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    qs.run_cycles([t % 2 for t in range(Tq)], "depolarizing", 0.05, None, uniforms=uq)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["config4_steane_cycles_bulk_draws"] = {"cycles": Tq, "seconds": dt, "cycles_per_s": Tq / dt}
    # complex64 mode of the headline workload (BASELINE: reported separately, tolerance 1e-5): same circuit, noise
    # and draws; 2^14 complex64 amplitudes per CTA -> clusters of 4 and all 148 SMs busy
    try:
        from qsb.lowering import lower_circuit
        nm16 = sim._noise_model
        T64, n16 = 2048, qc16.num_qubits
        u64 = ctx.to_device(np.random.default_rng(777).random((T64, 1)))     # placeholder, resized below
        res64 = {}
        for m64 in (14, 13):
            prog64, _ = lower_circuit(n16, qc16.get_ordered_gates(), sim._gate_registry,
                                      lambda name: nm16._channel_specs(name), local_bits=m64, max_local_bits=14)
            D64 = prog64.n_draws
            draws = np.random.default_rng(777).random((T64, D64))
            ctx.set_precision("c64")
            try:
                dp64 = ctx.program(prog64)
                u64 = ctx.to_device(draws)
                st64 = ctx.alloc(T64 * (8 << n16))
                ctx.run(dp64, 256, states=st64, uniforms=u64, uniforms_stride=D64)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.run(dp64, T64, states=st64, uniforms=u64, uniforms_stride=D64, async_=True)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                nrm = ctx.alloc(T64 * 16)
                ctx.overlap(n16, st64, 0, st64, 0, 1, T64, nrm)
                worst = float(np.max(np.abs(nrm.download(np.complex128, (T64,)) - 1.0)))
            finally:
                ctx.set_precision("c128")
            res64[f"local_bits_{m64}"] = {"cluster": 1 << (n16 - m64), "trajectories": T64, "ms": ms,
                                          "trajectories_per_s": T64 / (ms * 1e-3), "max_norm_error": worst}
        out["headline_complex64"] = res64
    except Exception as e:
        out["headline_complex64"] = {"error": repr(e)}
    # config 5 (single-GPU leg): 26-qubit layered circuit streamed through shared memory
    try:
        from qsb.bigstate import BigState
        nb = 26
        gl = [(g.gate_name, list(g.target_qubits), list(g.params)) for g in
              to_gate_instances(layered_circuit(nb, 20, 2026), GateInstance)]
        st = BigState(nb, layout="textbook", distributed=False)   # rank 0 alone: no collectives here
        st.apply_gates(gl[:8])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.apply_gates(gl)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["config5_streamed_26q"] = {"qubits": nb, "gates": len(gl), "seconds": dt, "gate_apps_per_s": len(gl) / dt,
                                       "algorithmic_GBps": len(gl) * 2 * 16 * 2 ** nb / dt / 1e9, "norm2": st.norm2()}
        del st
    except Exception as e:           # never let a secondary measurement break the headline line
        out["config5_streamed_26q"] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--traj", type=int, default=4096, help="trajectories per GPU per step")
    ap.add_argument("--e2e-traj", type=int, default=2048)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", action="store_true", help="(internal) print the cpu_baseline object and exit")
    ap.add_argument("--no-extras", action="store_true", help="skip the short secondary measurements (configs 2-5)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
