"""Streamed big-state probe (developer tool): pass count, time per pass, achieved HBM GB/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from qsb.bigstate import BigState, plan_distributed
from qsb.workloads import layered_circuit
from test_bigstate import ordered

for n in [int(x) for x in (sys.argv[1:] or ["24", "28"])]:
    gl = ordered(n, layered_circuit(n, 20, 2026))
    st = BigState(n, layout="textbook")
    lw = st.lowering()
    from quantum_sim.engine.gate_registry import GateRegistry
    reg = GateRegistry.instance()
    for name, targets, params in gl:
        lw.gate(name, targets, params, reg.get(name).matrix_func)
    steps, _ = plan_distributed(lw, 0, None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st.run(lw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    passes = len(steps)
    bytes_pass = 2 * 16 * 2 ** n
    print(f"n={n} gates={len(gl)} passes={passes} time={dt*1e3:.1f} ms  {dt/passes*1e3:.2f} ms/pass  "
          f"{passes*bytes_pass/dt/1e9:.0f} GB/s actual  {len(gl)*bytes_pass/dt/1e9:.0f} GB/s algorithmic  "
          f"{len(gl)/dt:.0f} gate-apps/s  norm2={st.norm2():.12f}", flush=True)
