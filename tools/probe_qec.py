import os, sys, time
sys.path[:0] = ["/root/repo", "/root/repo/quantum-simulator_b200"]
import numpy as np, torch
from quantum_sim.engine.qec import QECSimulator, SteaneCode
qs = QECSimulator(SteaneCode())
Tq = 16384
uq = np.random.default_rng(77).random((Tq, 7))
for rep in range(3):
    t0 = time.perf_counter()
    qs.run_cycles([t % 2 for t in range(Tq)], "depolarizing", 0.05, None, uniforms=uq)
    torch.cuda.synchronize()
    print("bulk", rep, time.perf_counter() - t0, flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
qs.run_cycles([t % 2 for t in range(Tq)], "depolarizing", 0.05, None, uniforms=uq)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
