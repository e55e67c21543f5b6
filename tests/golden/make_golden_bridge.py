"""Freeze a request/response transcript of the REAL reference bridge handler (bridge/server.py:30-267) and the
`.qsim` / NoiseModel JSON formats (core/serialization.py, circuit.py:154-173, noise.py:262-298) into
tests/golden/golden_bridge.json.

Build container only (needs /root/reference):   python tests/golden/make_golden_bridge.py

bridge/server.py imports PyQt6 at module level (for the QThread worker below the handler); PyQt6 is not installed
here, so a stub module stands in for it -- the command handler itself is plain Python and runs unmodified on the
reference's own engine.
"""
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.append(os.path.join(ROOT, "quantum-simulator_b200"))       # only for qsb.workloads

qt = types.ModuleType("PyQt6")
qtcore = types.ModuleType("PyQt6.QtCore")
qtcore.QObject = object
qtcore.QThread = object
qtcore.pyqtSignal = lambda *a, **k: None
qt.QtCore = qtcore
sys.modules["PyQt6"] = qt
sys.modules["PyQt6.QtCore"] = qtcore

import numpy as np                                                   # noqa: E402
from quantum_sim.bridge.server import BridgeCommandHandler          # noqa: E402
from quantum_sim.bridge.protocol import BridgeMessage               # noqa: E402
from quantum_sim.core.serialization import CircuitSerializer        # noqa: E402
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance  # noqa: E402
from quantum_sim.engine.noise import (NoiseModel, BitFlipNoise, PhaseFlipNoise, DepolarizingNoise,  # noqa: E402
                                      AmplitudeDampingNoise, ReadoutError)
from qsb.workloads import layered_circuit                           # noqa: E402

assert "/root/reference" in sys.modules["quantum_sim"].__file__

out = {"files": {}, "noise_models": {}, "transcript": []}

# ---- .qsim files written by the reference's own serializer -------------------------------------------------
def qsim_text(qc):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c" + CircuitSerializer.FILE_EXTENSION)
        CircuitSerializer.save(qc, p)
        back = CircuitSerializer.load(p)
        assert back.to_dict() == qc.to_dict()
        return open(p, encoding="utf-8").read()

qc5 = QuantumCircuit(5, initial_states=[0, 1, 0, 0, 1])
for g in layered_circuit(5, 6, 31):
    qc5.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
qc5.add_gate(GateInstance("Measure", [2], [], 7))
qc3 = QuantumCircuit(3)
qc3.add_gate(GateInstance("H", [0], [], 0))
qc3.add_gate(GateInstance("CNOT", [0, 1], [], 1))
qc3.add_gate(GateInstance("CNOT", [0, 2], [], 2))
qc10 = QuantumCircuit(10)
for g in layered_circuit(10, 8, 2026):
    qc10.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
out["files"]["layered5.qsim"] = qsim_text(qc5)
out["files"]["ghz3.qsim"] = qsim_text(qc3)
out["files"]["layered10.qsim"] = qsim_text(qc10)

# ---- noise-model dicts ---------------------------------------------------------------------------------------
nm = NoiseModel()
nm.add_global_noise(DepolarizingNoise(0.03))
nm.add_global_noise(AmplitudeDampingNoise(0.05))
nm.add_gate_noise("CNOT", BitFlipNoise(0.04))
nm.add_gate_noise("H", PhaseFlipNoise(0.02))
nm.set_readout_error(ReadoutError(0.02, 0.05))
out["noise_models"]["mixed"] = nm.to_dict()
assert NoiseModel.from_dict(nm.to_dict()).to_dict() == nm.to_dict()
nm2 = NoiseModel()
nm2.add_global_noise(BitFlipNoise(0.1))
out["noise_models"]["bitflip"] = nm2.to_dict()

# ---- transcript through the reference's handler --------------------------------------------------------------
h = BridgeCommandHandler()
seq = [0]


def send(action, **params):
    seq[0] += 1
    req = BridgeMessage(type="request", id=f"m{seq[0]}", action=action, params=params)
    wire = BridgeMessage.from_json(req.to_bytes().decode("utf-8"))       # what the server decodes off the socket
    resp = h.handle(wire)
    back = json.loads(resp.to_bytes().decode("utf-8"))                     # what the client reads back
    out["transcript"].append({"request": json.loads(req.to_json()), "response": back})
    return back


def seed_noise(seed):
    # not a protocol message: the GUI owns the handler's NoiseModel; NoiseModel.from_dict leaves its generator
    # unseeded (noise.py:192), so the replay seeds it the same way to make the noisy counts comparable
    h._noise_model.set_seed(seed)
    out["transcript"].append({"local": "seed_noise", "seed": seed})


send("ping")
send("run")                                            # error: no circuit
send("set_circuit", circuit=json.loads(out["files"]["ghz3.qsim"]))
send("get_circuit")
send("run", shots=512, seed=7)
send("get_state")
send("get_result")
send("get_analysis", metrics=["fidelity", "entropy", "purity", "pauli"])
send("add_gate", gate_name="Ry", target_qubits=[1], params=[0.7], column=3)
send("add_gate", gate_name="Toffoli", target_qubits=[2, 0, 1], params=[], column=4)
send("run", shots=300, seed=8)
send("get_state")
send("set_circuit", circuit=json.loads(out["files"]["layered5.qsim"]))
send("run", shots=1000, seed=21)
send("get_state")
send("get_analysis", metrics=["fidelity", "entropy", "purity", "pauli"])
send("set_noise", noise_model=out["noise_models"]["mixed"])
seed_noise(99)
send("run", shots=200, seed=22)                         # run_with_noise: 200 trajectories, one basis index each
send("get_result")
send("run", shots=0, seed=23)                           # noisy, shots = 0 -> Simulator.run with noise, no sampling
send("get_state")
send("get_analysis", metrics=["fidelity", "entropy", "purity"])
send("clear_noise")
send("sweep_parameter", param="noise_p", values=[0.0, 0.02, 0.1], shots=0, seed=5, trials=12)
send("sweep_parameter", values=[0.05], shots=16, seed=6, trials="bad")       # trials falls back to 50
send("set_circuit", circuit=json.loads(out["files"]["layered10.qsim"]))
send("run", shots=2048, seed=3)
send("get_state")
send("set_noise", noise_model=out["noise_models"]["bitflip"])
seed_noise(4)
send("run", shots=64, seed=9)
send("clear_circuit")
send("get_circuit")
send("run", shots=10, seed=1)
send("bogus")
send("set_noise")                                       # error: missing param

path = os.path.join(HERE, "golden_bridge.json")
with open(path, "w", encoding="utf-8") as f:
    json.dump(out, f, indent=1)
print("wrote", path, os.path.getsize(path), "bytes;", len(out["transcript"]), "transcript entries")
