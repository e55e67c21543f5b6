"""CPU checks of the drop-in import route (INTEGRATION.md option A): `qsb.launcher.activate` keeps this repo's
engine modules in front and lets the reference's compute-free modules (reference.py, benchmarks.py, algorithms.py,
comparison.py, core/, bridge/protocol.py) import on top of them.  No device call is made here."""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantum-simulator_b200")
REF = next((p for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")
            if os.path.isfile(os.path.join(p, "quantum_sim", "engine", "reference.py"))), None)

pytestmark = pytest.mark.skipif(REF is None, reason="no reference checkout (baseline/_ref or /root/reference)")

MIRRORED = ["state_vector", "simulator", "noise", "measurement", "analysis", "qec", "optimizer", "debugger", "circuit",
            "gates", "gate_registry"]
OVERLAID = ["reference", "benchmarks", "algorithms", "comparison"]


def test_overlay_resolves_every_engine_module_the_callers_import():
    from qsb import launcher
    launcher.activate(REF)
    import importlib
    for name in MIRRORED:
        mod = importlib.import_module(f"quantum_sim.engine.{name}")
        assert os.path.abspath(mod.__file__).startswith(PKG), name
    for name in OVERLAID:
        mod = importlib.import_module(f"quantum_sim.engine.{name}")
        assert os.path.abspath(mod.__file__).startswith(REF), name
    # the overlaid modules bind the CUDA-backed classes through their relative imports
    import quantum_sim.engine.comparison as cmp
    import quantum_sim.engine.simulator as sim
    assert cmp.Simulator is sim.Simulator
    # the packages around the engine come from the reference unchanged
    import quantum_sim.core.serialization as ser
    import quantum_sim.bridge.protocol as proto
    assert os.path.abspath(ser.__file__).startswith(REF) and os.path.abspath(proto.__file__).startswith(REF)
    assert ser.QuantumCircuit is sys.modules["quantum_sim.engine.circuit"].QuantumCircuit


def test_algorithm_templates_and_benchmark_suite_build_on_the_mirrored_model():
    """engine/algorithms.py and engine/benchmarks.py only construct circuits (no array work) -- they must accept the
    mirrored QuantumCircuit / GateInstance."""
    from qsb import launcher
    launcher.activate(REF)
    import quantum_sim.engine.algorithms as alg
    import quantum_sim.engine.benchmarks as bm
    T = alg.AlgorithmTemplate
    circuits = [T.bell_state(), T.ghz_state(4), T.quantum_fourier_transform(3), T.inverse_qft(3), T.grover_search(3, 5),
                T.deutsch_jozsa(3), T.quantum_teleportation(), T.bernstein_vazirani("101"), T.superdense_coding()]
    for qc in circuits:
        assert os.path.abspath(sys.modules[type(qc).__module__].__file__).startswith(PKG)
        assert sum(len(col) for col in qc.get_ordered_gates()) == len(qc.gates)
    assert len(T.list_templates()) >= 9
    assert hasattr(bm, "BenchmarkSuite")


def test_launcher_cli_runs_a_script_as_main(tmp_path):
    """`python -m qsb.launcher script args` -- the script sees `quantum_sim.engine` = this repo's package."""
    script = tmp_path / "probe.py"
    script.write_text(
        "import sys, os\n"
        "sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))\n"
        "import quantum_sim.engine.simulator as s, quantum_sim.engine.reference as r\n"
        "print('SIM', s.__file__); print('REF', r.__file__); print('ARGV', sys.argv[1:])\n")
    env = dict(os.environ, PYTHONPATH=PKG, QSB_REFERENCE_ROOT=REF)
    res = subprocess.run([sys.executable, "-m", "qsb.launcher", str(script), "--x", "1"], env=env, capture_output=True,
                         text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stderr
    out = dict(ln.split(" ", 1) for ln in res.stdout.strip().splitlines())
    assert os.path.abspath(out["SIM"]).startswith(PKG)
    assert os.path.abspath(out["REF"]).startswith(REF)
    assert out["ARGV"] == "['--x', '1']"
