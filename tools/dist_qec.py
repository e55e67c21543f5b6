"""Sharded QEC threshold sweep under torchrun (BASELINE config 4).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/dist_qec.py [--trials 4096]

Rank 0 checks the sharded Steane sweep against the golden sweep of the real reference (15 points, six metrics), then every rank runs its share of a larger sweep and rank 0 prints cycles/s.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=4096)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["QSB_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from quantum_sim.engine.qec import QECSimulator, SteaneCode
    sim = QECSimulator(SteaneCode())
    out = {"world": world}
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        sw = next(x for x in json.load(f)["qec_sweeps"] if x["code"] == "steane")
    pts = sim.threshold_sweep_sharded(sw["probs"], n_trials=sw["trials"], noise_type=sw["noise_type"], seed=sw["seed"])
    if rank == 0:
        fields = ("logical_rate", "success_rate", "avg_fidelity", "logical_z_fidelity", "decoder_success_rate", "projection_logical_rate")
        worst = max(abs(getattr(pt, f) - g[f]) for pt, g in zip(pts, sw["points"]) for f in fields)
        out["parity_vs_reference_sweep"] = {"points": len(pts), "trials": sw["trials"], "max_abs_diff": worst}
    probs2 = [0.001, 0.005, 0.01, 0.02, 0.05, 0.1]
    sim.threshold_sweep_sharded(probs2[:1], n_trials=args.trials, noise_type="depolarizing", seed=1)      # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pts = sim.threshold_sweep_sharded(probs2, n_trials=args.trials, noise_type="depolarizing", seed=7)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        out["sweep"] = {"points": len(probs2), "trials_per_point": args.trials, "seconds": dt,
                        "cycles_per_s": len(probs2) * args.trials / dt,
                        "logical_rates": [pt.logical_rate for pt in pts]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
