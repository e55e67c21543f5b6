"""Sharded big-state run under torchrun (BASELINE config 5): parity at a size the oracle can check, then
the full-size circuit with device timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_big.py [--check-n 22] [--qubits 30] [--depth 20]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check-n", type=int, default=22)
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--no-fuse", action="store_true", help="NCCL all-to-all exchanges instead of peer loads folded into the next pass")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["QSB_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from qsb.bigstate import BigState, plan_distributed
    from qsb.workloads import layered_circuit
    from test_bigstate import ordered
    out = {"world": world}
    # ---- parity against the oracle (reference semantics incl. the axis scramble) at a checkable size
    n = args.check_n
    gl = ordered(n, layered_circuit(n, 3, 7 + n))
    st = BigState(n, fuse_exchange=not args.no_fuse)
    st.apply_gates(gl)
    shard = torch.from_numpy(st.local_shard().view(np.float64).copy()).cuda()
    parts = [torch.empty_like(shard) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, shard)
    else:
        parts = [shard]
    nrm = st.norm2()
    if rank == 0:
        from oracle import qsim_oracle as O
        got = st.to_reference_order([p.cpu().numpy().view(np.complex128) for p in parts])
        ref = np.zeros(2 ** n, dtype=np.complex128)
        ref[0] = 1.0
        for name, targets, params in gl:
            ref = O.apply_gate(ref, n, O.gate_matrix(name, params), targets)
        out["check"] = {"n": n, "gates": len(gl), "max_abs_err": float(np.max(np.abs(got - ref))), "norm2": nrm}
    del st, shard, parts
    torch.cuda.empty_cache()
    # ---- full size, timed on the device (max over ranks)
    n = args.qubits
    gl = ordered(n, layered_circuit(n, args.depth, 2026))
    st = BigState(n, layout="textbook", fuse_exchange=not args.no_fuse)
    lw = st.lowering()
    from quantum_sim.engine.gate_registry import GateRegistry
    reg = GateRegistry.instance()
    for name, targets, params in gl:
        lw.gate(name, targets, params, reg.get(name).matrix_func)
    steps, _ = plan_distributed(lw, st.g, None)
    kinds = [s.kind for s in steps]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.run(lw)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nrm = st.norm2()
    if rank == 0:
        ms = float(ms.item())
        out["run"] = {"n": n, "gates": len(gl), "passes": kinds.count("pass"), "reorders": kinds.count("reorder"),
                      "exchanges": kinds.count("exchange"), "ms": ms, "gate_apps_per_s": len(gl) / ms * 1e3,
                      "algorithmic_GBps_per_gpu": len(gl) * 2 * 16 * 2 ** n / world / ms / 1e6, "norm2": nrm,
                      "fused_exchanges": getattr(st, "fused_exchanges", 0), "symm_error": getattr(st, "_symm_error", None)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
