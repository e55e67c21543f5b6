"""Run the reference's own drivers UNCHANGED on the B200 engine.

The reference has no plugin layer: its scripts, `test_validation.py`, bridge and GUI import
`quantum_sim.engine.*` by name and prepend their own checkout to `sys.path`
(scripts/qec_threshold.py:15, test_validation.py:17).  This launcher makes those imports land on this
package instead:

  1. import this repo's `quantum_sim` / `quantum_sim.engine` first (cached `sys.modules` entries win over
     whatever the script later puts on `sys.path`);
  2. append the reference checkout's `quantum_sim/` and `quantum_sim/engine/` directories to the two
     packages' `__path__`, so modules this repo does NOT mirror because they do no array work
     (`engine/reference.py`, `benchmarks.py`, `algorithms.py`, `comparison.py`, `core/*`, `bridge/*`) are
     the reference's own files -- and their relative imports (`from .simulator import Simulator`) resolve
     to the mirrored, CUDA-backed modules, which come first on `__path__`;
  3. `runpy.run_path(script, run_name="__main__")` with the script's argv.

    PYTHONPATH=quantum-simulator_b200 python -m qsb.launcher /path/to/Quantum-Simulator/scripts/noise_sweep.py \
        --circuit ghz3 --noise depolarizing --steps 4 --trials 20 --seed 42

The reference checkout is found from the script path (its parent holding `quantum_sim/engine`), from
`$QSB_REFERENCE_ROOT`, or given explicitly.
"""

from __future__ import annotations

import os
import runpy
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # quantum-simulator_b200/


def find_reference_root(script_path=None):
    cands = []
    if os.environ.get("QSB_REFERENCE_ROOT"):
        cands.append(os.environ["QSB_REFERENCE_ROOT"])
    if script_path:
        d = os.path.dirname(os.path.abspath(script_path))
        cands += [d, os.path.dirname(d)]
    for c in cands:
        if os.path.isfile(os.path.join(c, "quantum_sim", "engine", "state_vector.py")):
            return os.path.abspath(c)
    return None


def activate(reference_root=None):
    """Make `quantum_sim` this repo's engine (+ the reference's compute-free modules as an overlay).
    Returns the imported `quantum_sim.engine` package."""
    if _PKG_ROOT not in sys.path:
        sys.path.insert(0, _PKG_ROOT)
    loaded = sys.modules.get("quantum_sim")
    if loaded is not None and not os.path.abspath(getattr(loaded, "__file__", "") or "").startswith(_PKG_ROOT):
        raise RuntimeError("another `quantum_sim` is already imported; activate() must run first")
    import quantum_sim
    import quantum_sim.engine as eng
    from . import capi
    capi.load_library()                       # no CPU fallback: fail here, not in the middle of a script
    if reference_root is not None:
        top = os.path.join(reference_root, "quantum_sim")
        sub = os.path.join(top, "engine")
        if not os.path.isdir(sub):
            raise FileNotFoundError(f"{sub}: not a Quantum-Simulator checkout")
        if top not in quantum_sim.__path__:
            quantum_sim.__path__.append(top)
        if sub not in eng.__path__:
            eng.__path__.append(sub)
    return eng


def run_script(script_path, argv=(), reference_root=None):
    """Run one of the reference's drivers as `__main__` on the B200 engine; returns its globals.
    SystemExit propagates (test_validation.py ends with sys.exit(main()))."""
    script_path = os.path.abspath(script_path)
    root = reference_root or find_reference_root(script_path)
    activate(root)
    old_argv = sys.argv
    sys.argv = [script_path] + [str(a) for a in argv]
    try:
        return runpy.run_path(script_path, run_name="__main__")
    finally:
        sys.argv = old_argv


def main(args=None):
    args = list(sys.argv[1:] if args is None else args)
    if not args or args[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    run_script(args[0], args[1:])
    return 0


if __name__ == "__main__":
    sys.exit(main())
