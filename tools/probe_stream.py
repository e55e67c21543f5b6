"""Streamed-state timing probe (developer tool): compile once, time repeated executions of the pass list.
   python tools/probe_stream.py [n=26] [engine=tma|executor] [low_bits] [box_bits]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
from qsb import stream as S
from qsb.bigstate import BigState
from qsb.workloads import layered_circuit
from quantum_sim.engine.gate_registry import GateRegistry
from test_bigstate import ordered

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
engine = sys.argv[2] if len(sys.argv) > 2 else "tma"
low = int(sys.argv[3]) if len(sys.argv) > 3 else None
box = int(sys.argv[4]) if len(sys.argv) > 4 else None
gl = ordered(n, layered_circuit(n, 20, 2026))
reg = GateRegistry.instance()
st = BigState(n, layout="textbook", distributed=False, engine=engine)
lw = st.lowering()
for name, targets, params in gl:
    lw.gate(name, targets, params, reg.get(name).matrix_func)
if engine == "tma":
    if low is not None or box is not None:
        _plan = S.plan
        S.plan = lambda *a, **k: _plan(*a, **{**k, "low_bits": low, "box_bits": box})
    t0 = time.perf_counter()
    comp = st.compile(lw)
    t_compile = time.perf_counter() - t0
    steps = comp[0]
    sweeps = [len(s.spass.blocks) for s in steps]
    gates = sum(len(s.spass.sweeps) for s in steps)
    st.execute(comp)
    ts = []
    for _ in range(5):
        st.ctx.timer_start()
        st.execute(comp, sync=False)
        ts.append(st.ctx.timer_stop())
    ms = min(ts)
    real = len(steps) * 2 * 16 * 2 ** n
    print(f"n={n} engine=tma variant={os.environ.get('QSB_STREAM_VARIANT', '0')} geometry={steps[0].spass.m, steps[0].spass.l, steps[0].spass.e} "
          f"gates={len(gl)} passes={len(steps)} multi-qubit gates + flushes={gates} block sweeps={sum(sweeps)} (per pass {sweeps}) compile={t_compile*1e3:.1f} ms")
    print(f"   best of 5: {ms:.3f} ms = {ms/len(steps):.3f} ms/pass, {real/ms/1e6:.0f} GB/s real (read+write), "
          f"{len(gl)/ms*1e3:.0f} gate-apps/s, all runs {['%.2f' % t for t in ts]}")
    # per pass timing
    for k, s in enumerate(steps[:6]):
        st.ctx.timer_start()
        s.handle.run(st._wrapped[st.cur])
        print(f"   pass {k}: block sweeps={len(s.spass.blocks)} {st.ctx.timer_stop():.3f} ms")
    print("   norm2 =", st.norm2())
else:
    st.run(lw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st.run(lw)
    torch.cuda.synchronize()
    print(f"n={n} engine=executor: {1e3*(time.perf_counter()-t0):.2f} ms wall (plan + create + per-pass sync included)")
