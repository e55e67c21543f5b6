"""cProfile of the Philox QEC sweep (developer tool): where the host time of BASELINE config 4's throughput mode goes.
   python tools/probe_qec_profile.py [points=4] [trials=16384]"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200")]
import numpy as np
from quantum_sim.engine.qec import QECSimulator, SteaneCode

points = int(sys.argv[1]) if len(sys.argv) > 1 else 4
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
sim = QECSimulator(SteaneCode())
probs = list(np.linspace(0.005, 0.3, points))
sim.threshold_sweep_philox(probs[:1], n_trials=2048, seed=1)          # warm-up: programs, pools
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
sim.threshold_sweep_philox(probs, n_trials=trials, seed=2)
pr.disable()
dt = time.perf_counter() - t0
print(f"{points} points x {trials} trials: {dt:.3f} s = {points * trials / dt:.0f} cycles/s")
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(28)
print(out.getvalue()[:6000])
