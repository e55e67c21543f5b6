"""Recipe: put the UNMODIFIED reference next to the repo as `baseline/_ref/` (git-ignored, not gpurun-ignored).

The reference (justinbrianhwang/Quantum-Simulator) is pure Python + NumPy: no build step, no setup.py /
pyproject (so `pip install --target baseline/_ref /root/reference` has nothing to install -- recorded in
DESIGN.md).  "Installing" it is a byte-for-byte copy of the files the hot path and its acceptance drivers need:

    quantum_sim/__init__.py, quantum_sim/engine/, quantum_sim/core/, quantum_sim/bridge/{protocol,client}.py
    scripts/*.py, test_validation.py

Two users, both outside the product path:
  * `bench.py --impl reference` and bench's `cpu_baseline` leg import `baseline/_ref/quantum_sim` (the reference's own
    `Simulator`) in a process that never imports this repo's engine -- the reference arm is the reference's code;
  * `tests/test_gpu_acceptance.py` runs `baseline/_ref/test_validation.py` and `baseline/_ref/scripts/*.py` UNCHANGED
    on top of the B200 engine through `qsb.launcher` (which also overlays the reference's non-mirrored, compute-free
    engine modules: reference.py, benchmarks.py, algorithms.py, comparison.py).

Nothing under baseline/_ref is tracked by git and nothing in the product imports it.

    python tools/make_baseline_ref.py [--src /root/reference] [--force]
"""

from __future__ import annotations

import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
DEFAULT_SRC = "/root/reference"

FILES = ["test_validation.py", "quantum_sim/__init__.py", "quantum_sim/bridge/__init__.py",
         "quantum_sim/bridge/protocol.py", "quantum_sim/bridge/client.py", "LICENSE"]
DIRS = ["quantum_sim/engine", "quantum_sim/core", "scripts"]


def _wanted(src):
    out = list(FILES)
    for d in DIRS:
        for base, _, names in os.walk(os.path.join(src, d)):
            for nm in sorted(names):
                if nm.endswith(".py"):
                    out.append(os.path.relpath(os.path.join(base, nm), src))
    return [f for f in out if os.path.exists(os.path.join(src, f))]


def make(src=DEFAULT_SRC, force=False, quiet=False):
    """Copy the reference files; returns the destination, or None when the source tree is absent (GPU box)."""
    if not os.path.isdir(os.path.join(src, "quantum_sim", "engine")):
        return DEST if os.path.isdir(os.path.join(DEST, "quantum_sim", "engine")) else None
    copied = 0
    for rel in _wanted(src):
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        if not force and os.path.exists(d) and filecmp.cmp(s, d, shallow=False):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        copied += 1
    if not quiet:
        print(f"baseline/_ref: {copied} file(s) copied from {src}")
    return DEST


def present():
    return os.path.isfile(os.path.join(DEST, "quantum_sim", "engine", "simulator.py"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=DEFAULT_SRC)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    sys.exit(0 if make(a.src, a.force) else 1)
