"""ncu target (developer tool): three launches of one streamed pass of the 26-qubit layered circuit.
   python tools/ncu_stream.py [n=26] [pass index=0]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
from qsb.bigstate import BigState
from qsb.workloads import layered_circuit
from quantum_sim.engine.gate_registry import GateRegistry
from test_bigstate import ordered

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
gl = ordered(n, layered_circuit(n, 20, 2026))
reg = GateRegistry.instance()
st = BigState(n, layout="textbook", distributed=False)
lw = st.lowering()
for name, targets, params in gl:
    lw.gate(name, targets, params, reg.get(name).matrix_func)
steps, _, _ = st.compile(lw)
s = steps[k]
for _ in range(3):
    s.handle.run(st._wrapped[st.cur])
st.ctx.sync()
print(f"pass {k}: {len(s.spass.blocks)} block sweeps ({len(s.spass.sweeps)} gates), geometry {(s.spass.m, s.spass.l, s.spass.e)}; 3 launches done")
