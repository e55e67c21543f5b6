"""ncu target (developer tool): all-pairs RDMs (qsb_rdm_gram_kernel) + MI of 8000 random 12-qubit states."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200")]
import numpy as np
from qsb import capi
n, count = 12, 8000
ctx = capi.get_context()
rng = np.random.default_rng(1)
psi = (rng.normal(size=(count, 2 ** n)) + 1j * rng.normal(size=(count, 2 ** n))) / np.sqrt(2.0 ** (n + 1))
s = ctx.to_device(psi)
npairs = n * (n - 1) // 2
r1, r2 = ctx.alloc(count * n * 64), ctx.alloc(count * npairs * 256)
for rep in range(3):
    ctx.timer_start()
    ctx.rdm_all(n, s, 0, count, r1, r2)
    ms = ctx.timer_stop()
    print(f"rdm_all {count} x {n}q: {ms:.3f} ms = {count / ms * 1e3:.0f} states/s, {count * 2 ** n * 16 / ms / 1e6:.0f} GB/s of state reads, "
          f"{count * npairs * 2 ** (n - 4) * 512 / ms / 1e9:.2f} TFLOP/s of DMMA")
mi = ctx.alloc(count * npairs * 8)
ctx.timer_start()
ctx.mi_all_pairs(n, s, 0, count, mi)
print(f"mi_all_pairs: {ctx.timer_stop():.3f} ms")
