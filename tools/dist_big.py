"""Sharded big-state run under torchrun (BASELINE config 5): parity against the oracle at a size it can check,
then the full-size circuit with device timing (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_big.py [--check-n 22] [--qubits 30] [--depth 20] [--engine tma|executor] [--no-fuse]

Prints one JSON object on rank 0.  tests/test_gpu_multi.py runs it on 2 GPUs; bench.py has the same leg.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist


def check_against_oracle(n, depth, world, rank, engine, fuse, where="store"):
    """Reference semantics (axis scramble included) on every rank's shard, gathered and compared on rank 0."""
    from qsb.bigstate import BigState
    from qsb.workloads import layered_circuit
    from test_bigstate import ordered
    gl = ordered(n, layered_circuit(n, depth, 7 + n))
    st = BigState(n, fuse_exchange=fuse, engine=engine, fuse_where=where)
    st.apply_gates(gl)
    shard = torch.from_numpy(st.local_shard().view(np.float64).copy()).cuda()
    parts = [torch.empty_like(shard) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, shard)
    else:
        parts = [shard]
    nrm = st.norm2()
    res = None
    if rank == 0:
        from oracle import qsim_oracle as O
        got = st.to_reference_order([p.cpu().numpy().view(np.complex128) for p in parts])
        ref = np.zeros(2 ** n, dtype=np.complex128)
        ref[0] = 1.0
        for name, targets, params in gl:
            ref = O.apply_gate(ref, n, O.gate_matrix(name, params), targets)
        res = {"n": n, "gates": len(gl), "max_abs_err": float(np.max(np.abs(got - ref))), "norm2": nrm,
               "fused_exchanges": st.fused_exchanges, "launches": st.launches}
    del st, shard, parts
    torch.cuda.empty_cache()
    return res


def timed_run(n, depth, world, rank, engine, fuse, reps=3, where="store"):
    from qsb.bigstate import BigState
    from qsb.workloads import layered_circuit
    from quantum_sim.engine.gate_registry import GateRegistry
    from test_bigstate import ordered
    gl = ordered(n, layered_circuit(n, depth, 2026))
    st = BigState(n, layout="textbook", fuse_exchange=fuse, engine=engine, fuse_where=where)
    lw = st.lowering()
    reg = GateRegistry.instance()
    for name, targets, params in gl:
        lw.gate(name, targets, params, reg.get(name).matrix_func)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = []
    if engine == "tma":
        comp = st.compile(lw)                                   # plan + upload, outside the timed region
        kinds = [s.kind for s in comp[0]] + ["exchange"] * sum(1 for s in comp[0] if s.scatter)
        st.execute(comp)                                        # warm-up (also maps the second buffer)
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            st.execute(comp, sync=False)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    else:
        from qsb.bigstate import plan_distributed
        kinds = [s.kind for s in plan_distributed(lw, st.g, None)[0]]
        st.run(lw)
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            st.run(lw)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    ms = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nrm = st.norm2()
    ms = float(ms.item())
    passes = kinds.count("pass") + kinds.count("reorder")
    shard_bytes = 16 * 2 ** n / world
    g = world.bit_length() - 1
    return {"n": n, "gates": len(gl), "passes": kinds.count("pass"), "reorders": kinds.count("reorder"),
            "exchanges": kinds.count("exchange"), "fused_exchanges_per_run": st.fused_exchanges // (reps + 1),
            "ms": ms, "all_ms": times, "gate_apps_per_s": len(gl) / ms * 1e3,
            "real_GBps_per_gpu": passes * 2 * shard_bytes / ms / 1e6,
            "nvlink_bytes_per_gpu_per_direction": kinds.count("exchange") * (1 - 2.0 ** -g) * shard_bytes,
            "algorithmic_GBps_per_gpu": len(gl) * 2 * shard_bytes / ms / 1e6, "norm2": nrm,
            "symm_error": getattr(st, "_symm_error", None)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check-n", type=int, default=22)
    ap.add_argument("--check-depth", type=int, default=3)
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--engine", default="tma", choices=["tma", "executor"])
    ap.add_argument("--no-fuse", action="store_true", help="NCCL all-to-all exchanges instead of peer-mapped TMA stores / loads")
    ap.add_argument("--fuse-where", default="store", choices=["store", "load"],
                    help="fold an exchange into the store of the pass before it (default) or into the load of the pass after it")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["QSB_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "engine": args.engine, "fuse": not args.no_fuse, "fuse_where": args.fuse_where}
    if args.check_n > 0:
        out["check"] = check_against_oracle(args.check_n, args.check_depth, world, rank, args.engine, not args.no_fuse,
                                            args.fuse_where)
    if args.qubits > 0:
        out["run"] = timed_run(args.qubits, args.depth, world, rank, args.engine, not args.no_fuse, where=args.fuse_where)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
