"""compute-sanitizer target (developer tool): a small streamed circuit (17 qubits, block sweeps, one reorder pass, a
scatter pass onto two stand-in shards) through qsb_stream_kernel, checked against the CPU replay of the same plan."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from qsb import capi, stream as S
from qsb.workloads import layered_circuit
from test_bigstate import ordered, lower
from test_stream_plan import replay

n = 17
gl = ordered(n, layered_circuit(n, 2, 3))
lw = lower(n, gl, layout="reference")
cdata = lw.pool.array()
steps, _, _ = S.plan(lw.items, cdata, n, 0, list(range(n)))
rng = np.random.default_rng(0)
psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
ctx = capi.get_context()
buf = ctx.to_device(psi)
hs = [ctx.stream_pass(st.spass, cdata) for st in steps]
for h in hs:
    h.run(buf)
ctx.sync()
got = buf.download(np.complex128, (2 ** n,))
print("passes", len(steps), "max |device - replay| =", float(np.max(np.abs(got - replay(steps, n, 0, psi, cdata)))))
# a permuting (reorder) pass and a scatter pass with two stand-in shards
L = 16
new = list(range(5)) + (5 + rng.permutation(L - 5)).tolist()
sp = S.StreamPass(L, 12, 5, 3, list(range(L)), new, [])
src, dst = ctx.to_device(psi[:2 ** L]), ctx.alloc(16 << L).zero()
ctx.stream_pass(sp, np.zeros(2)).run(src, dst)
ctx.sync()
steps2, _, _ = S.plan(lw.items, cdata, n, 1, list(range(n)))
st = next(s for s in steps2 if s.scatter)
outs = [ctx.alloc(16 << L).zero() for _ in range(2)]
h = ctx.stream_pass(st.spass, cdata)
for r in range(2):
    h.run_scatter(ctx.to_device(psi[r << L:(r + 1) << L]), [o.ptr for o in outs], L - 1, r << (L - 1))
ctx.sync()
print("reorder + scatter passes done")
