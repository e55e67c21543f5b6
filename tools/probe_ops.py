"""Per-op cost probe: micro-programs of one op kind each (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200")]
import numpy as np
from qsb import capi
from qsb.compiler import Lowering
from quantum_sim.engine import gates as G

ctx = capi.get_context(0)
CLK = 1.9e9


def run(label, n, build, T, reps=3, m=None, nops=None):
    lw = Lowering(n, layout="textbook")
    build(lw)
    prog = lw.finish(m)
    dp = ctx.program(prog)
    states = ctx.alloc(T * (1 << n) * 16)
    kw = {}
    if prog.n_draws:
        u = np.full((T, prog.n_draws), 0.5) if label.startswith("certain") else np.random.default_rng(0).random((T, prog.n_draws))
        kw.update(uniforms=ctx.to_device(u), uniforms_stride=prog.n_draws)
    best = 1e9
    for r in range(reps):
        ctx.timer_start()
        ctx.run(dp, T, states=states, async_=True, **kw)
        best = min(best, ctx.timer_stop())
    ctx.profile(True)
    ctx.run(dp, T, states=states, **kw)
    pr = ctx.profile(True, read=True).astype(np.float64)[0]
    ctx.profile(False)
    C = 1 << (prog.n - prog.m)
    units = min(T, (148 // C) if C > 1 else 148)
    if C == 8: units = min(T, 15)
    rounds = -(-T // units)
    per_traj = best * 1e-3 / rounds
    k = nops or len(prog.ops)
    print(f"{label:40s} n={n} C={C} ops={len(prog.ops):5d} remaps={prog.n_remaps:4d} {best:8.3f} ms  "
          f"{per_traj * 1e6:9.1f} us/traj  {per_traj / k * CLK:9.0f} cyc/op | sweep busy/desc "
          f"{pr[3] / max(pr[11], 1):7.0f} n={pr[11]:5.0f} wwait {pr[0]:9.0f} ctl {pr[18]:9.0f} ringwait {pr[17]:9.0f}", flush=True)


N = 400
rng = np.random.default_rng(1)
def only_h(lw):
    for i in range(N): lw.matrix(G.H_MATRIX, [i % lw.n])
def only_cx(lw):
    for i in range(N): lw.gate("CNOT", [i % lw.n, (i + 1) % lw.n])
def cx_far(lw):           # always the two lowest qubits (highest slot bits)
    for i in range(N): lw.gate("CNOT", [0, 1])
def only_ccx(lw):
    for i in range(N): lw.gate("Toffoli", [i % lw.n, (i + 1) % lw.n, (i + 2) % lw.n])
def h_then_cx(lw):
    for i in range(N // 2):
        lw.matrix(G.H_MATRIX, [i % lw.n]); lw.gate("CNOT", [i % lw.n, (i + 1) % lw.n])
def only_depol(lw):
    for i in range(N): lw.kraus("depolarizing", 0.01, i % lw.n)
def only_ad(lw):
    for i in range(N): lw.kraus("amplitude_damping", 0.02, i % lw.n)
def empty(lw):
    pass
def u2(lw):
    m = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    for i in range(N): lw.matrix(m, [i % lw.n, (i + 1) % lw.n])

def ad_cx(lw):            # K0 = diag(1, s) pending on both qubits of every CX (uniforms 0.5: always the certain branch)
    for i in range(N // 3):
        lw.kraus("amplitude_damping", 0.02, 0); lw.kraus("amplitude_damping", 0.02, 1); lw.gate("CNOT", [0, 1])
def ad1_cx(lw):
    for i in range(N // 2):
        lw.kraus("amplitude_damping", 0.02, 1); lw.gate("CNOT", [0, 1])
def h2_cx(lw):
    for i in range(N // 3):
        lw.matrix(G.H_MATRIX, [0]); lw.matrix(G.H_MATRIX, [1]); lw.gate("CNOT", [0, 1])
def ad_h_ad_cx(lw):       # the noisy workload's typical pending: K0 . U . K0 on one qubit, K0 on the other
    for i in range(N // 5):
        lw.kraus("amplitude_damping", 0.02, 0); lw.matrix(G.H_MATRIX, [0]); lw.kraus("amplitude_damping", 0.02, 0)
        lw.kraus("amplitude_damping", 0.02, 1); lw.gate("CNOT", [0, 1])
def cx_pairs(lw):          # adjacent CX on disjoint bits: fused into K = 4 sweeps
    for i in range(N // 2):
        lw.gate("CNOT", [0, 1]); lw.gate("CNOT", [2, 3])
def cx_pairs_low(lw):
    for i in range(N // 2):
        lw.gate("CNOT", [12, 11]); lw.gate("CNOT", [10, 9])
def h4_cx_pairs(lw):
    for i in range(N // 6):
        for q in range(4): lw.matrix(G.H_MATRIX, [q])
        lw.gate("CNOT", [0, 1]); lw.gate("CNOT", [2, 3])
for n, T in ((13, 148),):
    run("200 x (CX, CX) disjoint, high bits", n, cx_pairs, T)
    run("200 x (CX, CX) disjoint, low bits", n, cx_pairs_low, T)
    run("(H x4, CX, CX) x 66", n, h4_cx_pairs, T)
    run("certain: (AD, AD, CX) x 133", n, ad_cx, T)
    run("certain: (AD, CX) x 200", n, ad1_cx, T)
    run("(H, H, CX) x 133", n, h2_cx, T)
    run("certain: (AD H AD, AD, CX) x 80", n, ad_h_ad_cx, T)
    run("400 CX same qubits", n, cx_far, T)
for n, T in ((13, 148), (16, 15)):
    run("empty (init+store)", n, empty, T, nops=1)
    run("400 H (scalar pending path)", n, only_h, T)
    run("400 depolarizing draws", n, only_depol, T)
    run("400 amplitude-damping draws", n, only_ad, T)
    run("400 CX cycling qubits", n, only_cx, T)
    run("400 CX same qubits", n, cx_far, T)
    run("400 CCX cycling", n, only_ccx, T)
    run("200 x (H, CX)", n, h_then_cx, T)
    run("400 dense U2", n, u2, T)
