"""Sharded trajectories under torchrun: `Simulator.run_with_noise_sharded` and
`ensemble_density_matrix_sharded` against goldens of the real reference, then a timed 16-qubit batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_traj.py
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["QSB_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from conftest import as_gates
    from qsb.workloads import ghz, layered_circuit, to_gate_instances
    from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
    from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise, AmplitudeDampingNoise
    from quantum_sim.engine.simulator import Simulator
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        gold = json.load(f)
    gnpz = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    out = {"world": world}

    def circuit(n, gates):
        qc = QuantumCircuit(n)
        for g in gates:
            qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
        return qc

    # 1. counts of the reference's run_with_noise (GHZ-3, depolarizing 0.1, noise seed 7, seed 42, 200 shots)
    nm = NoiseModel(); nm.add_global_noise(DepolarizingNoise(0.1)); nm.set_seed(7)
    res = Simulator(nm).run_with_noise_sharded(circuit(3, ghz(3)), shots=200, seed=42)
    want = gold["ghz3"]["run_with_noise"]
    out["run_with_noise_counts_equal"] = res.measurement_counts == want and list(res.measurement_counts) == list(want)
    # 2. ensemble rho against the reference's (golden ens4: gates, noise, trials, seed in the fixture)
    e = gold["ens4"]
    nm = NoiseModel()
    for kind, p in e["noise"]["global"]:
        nm.add_global_noise({"depolarizing": DepolarizingNoise, "amplitude_damping": AmplitudeDampingNoise}[kind](p))
    rho = Simulator(nm).ensemble_density_matrix_sharded(circuit(4, as_gates(e["gates"])), n_trials=e["n_trials"], seed=e["seed"])
    out["ensemble_rho_max_abs_diff"] = float(np.max(np.abs(rho - gnpz["ens4_rho"])))
    # 3. timed: 16-qubit noisy trajectories, 4096 shots per rank
    n = 16
    qc = QuantumCircuit(n)
    for g in to_gate_instances(layered_circuit(n, 64, 2026), GateInstance):
        qc.add_gate(g)
    nm = NoiseModel(); nm.add_global_noise(DepolarizingNoise(0.01)); nm.add_global_noise(AmplitudeDampingNoise(0.02)); nm.set_seed(5)
    sim = Simulator(nm)
    shots = 4096 * world
    sim.run_with_noise_sharded(qc, shots=shots, seed=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    r = sim.run_with_noise_sharded(qc, shots=shots, seed=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        out["timed"] = {"shots": shots, "seconds": dt, "trajectories_per_s": shots / dt, "distinct_outcomes": len(r.measurement_counts)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
