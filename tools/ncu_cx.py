"""Tiny program for an ncu source-level capture of the sweep path (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-simulator_b200")]
import numpy as np
from qsb import capi
from qsb.compiler import Lowering
ctx = capi.get_context(0)
n, T = 13, 148
lw = Lowering(n, layout="textbook")
for i in range(400):
    lw.gate("CNOT", [0, 1])
dp = ctx.program(lw.finish(None))
states = ctx.alloc(T * (1 << n) * 16)
for r in range(3):
    ctx.timer_start()
    ctx.run(dp, T, states=states, async_=True)
    print("ms", ctx.timer_stop())
