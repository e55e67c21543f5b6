"""Ensemble-rho accumulation probe (developer tool): DMMA kernel time and FP64 rate at config 3's size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200")]
import numpy as np
from qsb import capi
ctx = capi.get_context(0)
for n, N in ((10, 500), (12, 500), (12, 2048), (13, 500)):
    dim = 1 << n
    rng = np.random.default_rng(n)
    psi = (rng.normal(size=(N, dim)) + 1j * rng.normal(size=(N, dim))) / np.sqrt(2 * dim)
    states = ctx.to_device(psi)
    rho = ctx.alloc(16 * dim * dim).zero()
    best = 1e9
    for r in range(3):
        rho.zero()
        ctx.timer_start()
        ctx.rho_accumulate(n, states, 0, N, 1.0 / N, rho)
        best = min(best, ctx.timer_stop())
    flops = 8.0 * dim * dim * N
    err = None
    if n <= 10:
        got = rho.download(np.complex128, (dim, dim))
        ref = (psi.T @ psi.conj()) / N
        err = float(np.max(np.abs(got - ref)))
    print(f"n={n} N={N}: {best:8.3f} ms  {flops / best / 1e9:8.2f} TFLOP/s (8*4^n*N convention; the Hermitian half is computed)"
          f"  max err {err}", flush=True)
