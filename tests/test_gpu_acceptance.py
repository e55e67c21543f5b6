"""The reference's OWN acceptance drivers, unmodified, on the B200 engine.

`baseline/_ref/` is a byte-for-byte copy of the reference (tools/make_baseline_ref.py; git-ignored, shipped to the
GPU box).  `qsb.launcher` imports this repo's `quantum_sim.engine` first, overlays the reference's compute-free
modules (reference.py, benchmarks.py, ...) and `runpy`s the script as `__main__` -- exactly what INTEGRATION.md
tells a maintainer to do.  Expected results were frozen from the same scripts on the reference's NumPy engine
(tests/golden/make_golden_acceptance.py)."""

import io
import json
import os
import sys
from contextlib import redirect_stdout

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden", "acceptance")

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_acceptance import CASES  # noqa: E402  (the command lines, shared with the generator)

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "test_validation.py")),
                               reason="baseline/_ref missing: run `python tools/make_baseline_ref.py` where "
                                      "/root/reference exists (build() does it)")

# wall-clock fields and fields that are pure eigvalsh rounding noise around zero
IGNORED = {"elapsed_seconds"}
LOOSE = {"mean_entropy": 1e-9}


def _compare(got, want, path=""):
    if isinstance(want, dict):
        assert isinstance(got, dict) and sorted(got) == sorted(want), path
        for k in want:
            if k in IGNORED:
                continue
            _compare(got[k], want[k], f"{path}.{k}")
    elif isinstance(want, list):
        assert isinstance(got, list) and len(got) == len(want), path
        for i, (g, w) in enumerate(zip(got, want)):
            _compare(g, w, f"{path}[{i}]")
    elif isinstance(want, bool) or want is None or isinstance(want, (str, int)):
        assert got == want, (path, got, want)
    else:
        tol = LOOSE.get(path.rsplit(".", 1)[-1], 1e-10 if "cost" in path else 1e-12)
        assert abs(got - want) <= tol, (path, got, want)


def _run(script, argv):
    from qsb import launcher
    buf = io.StringIO()
    code = 0
    with redirect_stdout(buf):
        try:
            launcher.run_script(os.path.join(REF, script), argv, reference_root=REF)
        except SystemExit as e:
            code = e.code or 0
    return code, buf.getvalue()


@needs_ref
def test_the_engine_under_the_scripts_is_the_cuda_one():
    from qsb import launcher
    eng = launcher.activate(REF)
    import quantum_sim.engine.simulator as s
    import quantum_sim.engine.reference as r          # not mirrored: the reference's own file, via the overlay
    pkg = os.path.join(ROOT, "quantum-simulator_b200")
    assert os.path.abspath(s.__file__).startswith(pkg)
    assert os.path.abspath(r.__file__).startswith(REF)
    assert r.StateVector is sys.modules["quantum_sim.engine.state_vector"].StateVector
    assert os.path.abspath(sys.modules["quantum_sim.engine.state_vector"].__file__).startswith(pkg)
    assert eng.__path__[0].startswith(pkg)


@needs_ref
def test_validation_harness_33_of_33():
    """/root/reference/test_validation.py:537-576 prints `Results: 33/33 passed`; same lines as on the reference."""
    import re
    code, out = _run("test_validation.py", [])
    clock = lambda ln: re.sub(r"\d+\.\d+s <", "_s <", ln)            # the three wall-clock assertions print their time
    lines = [clock(ln.strip()) for ln in out.splitlines() if ln.strip().startswith("[") or ln.startswith("Results:")]
    with open(os.path.join(GOLD, "test_validation_transcript.json")) as f:
        want = [clock(ln) for ln in json.load(f)["lines"]]
    assert lines == want, "\n".join(lines)
    assert lines[-1].startswith("Results: 33/33 passed, 0 failed") and code == 0


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_script_unchanged(name, tmp_path):
    script, argv = CASES[name]
    with open(os.path.join(GOLD, name + ".json")) as f:
        want = json.load(f)
    assert want["script"] == script and want["argv"] == argv
    out_file = tmp_path / "out.json"
    code, _ = _run(script, argv + ["--output", str(out_file)])
    assert code == 0
    with open(out_file) as f:
        got = json.load(f)
    _compare(got, want["output"], name)
