"""Test infrastructure: replay a golden bridge transcript (tests/golden/golden_bridge.json, recorded from the
reference's own `BridgeCommandHandler`, bridge/server.py:30-267, on the reference engine) against THIS repo's engine.

`ReplayHandler` restates what each command does with the engine API (the reference handler imports
`quantum_sim.engine.*` lazily by name, so a maintainer gets the same effect by putting quantum-simulator_b200/ first
on sys.path and running the reference's bridge unchanged; the reference is not on the GPU box, hence this restatement).
Wire format: one JSON object per line with the fields of bridge/protocol.py:14-37.
"""
import json

import numpy as np

FIELDS = ("type", "id", "action", "params", "status", "data", "error")      # protocol.py:26-33


def decode(raw: bytes) -> dict:
    d = json.loads(raw.decode("utf-8").strip())
    return {"type": d.get("type", "request"), "id": d.get("id", ""), "action": d.get("action", ""),
            "params": d.get("params", {}), "status": d.get("status", ""), "data": d.get("data", {}),
            "error": d.get("error", "")}


def encode(msg: dict) -> bytes:
    return (json.dumps({k: msg[k] for k in FIELDS}, ensure_ascii=False) + "\n").encode("utf-8")


def _ok(rid, data=None):
    return {"type": "response", "id": rid, "action": "", "params": {}, "status": "ok", "data": data or {}, "error": ""}


def _err(rid, text):
    return {"type": "response", "id": rid, "action": "", "params": {}, "status": "error", "data": {}, "error": text}


class ReplayHandler:
    def __init__(self):
        self.circuit = None
        self.noise_model = None
        self.last_result = None
        self.ideal_state = None

    def handle(self, msg):
        fn = getattr(self, "cmd_" + msg["action"], None)
        if fn is None:
            return _err(msg["id"], f"Unknown action: {msg['action']}")                 # server.py:63-66
        try:
            return fn(msg)
        except Exception as e:                                                         # server.py:69-71
            return _err(msg["id"], str(e))

    def cmd_ping(self, m):
        return _ok(m["id"], {"pong": True})

    def cmd_get_circuit(self, m):
        if self.circuit is None:
            return _err(m["id"], "No circuit loaded")
        return _ok(m["id"], self.circuit.to_dict())

    def cmd_set_circuit(self, m):
        from quantum_sim.engine.circuit import QuantumCircuit
        cd = m["params"].get("circuit")
        if cd is None:
            return _err(m["id"], "Missing 'circuit' param")
        self.circuit = QuantumCircuit.from_dict(cd)
        return _ok(m["id"], {"num_qubits": self.circuit.num_qubits, "gate_count": self.circuit.gate_count()})

    def cmd_add_gate(self, m):
        from quantum_sim.engine.circuit import GateInstance
        if self.circuit is None:
            return _err(m["id"], "No circuit loaded")
        p = m["params"]
        self.circuit.add_gate(GateInstance(gate_name=p.get("gate_name", "H"), target_qubits=p.get("target_qubits", [0]),
                                           params=p.get("params", []), column=p.get("column", 0)))
        return _ok(m["id"], {"gate_count": self.circuit.gate_count()})

    def cmd_clear_circuit(self, m):
        if self.circuit is None:
            return _err(m["id"], "No circuit loaded")
        self.circuit.clear()
        return _ok(m["id"])

    def cmd_run(self, m):                                                              # server.py:117-142
        from quantum_sim.engine.simulator import Simulator
        if self.circuit is None:
            return _err(m["id"], "No circuit loaded")
        shots, seed = m["params"].get("shots", 1024), m["params"].get("seed")
        sim = Simulator(noise_model=self.noise_model)
        if self.noise_model is not None and shots > 0:
            res = sim.run_with_noise(self.circuit, shots=shots, seed=seed)
        else:
            res = sim.run(self.circuit, shots=shots, seed=seed)
        self.last_result = res
        if self.noise_model is None:
            self.ideal_state = res.final_state
        return _ok(m["id"], {"measurement_counts": res.measurement_counts, "num_shots": res.num_shots, "seed": res.seed})

    def cmd_get_state(self, m):                                                        # server.py:144-157
        if self.last_result is None:
            return _err(m["id"], "No simulation result")
        sv = self.last_result.final_state
        return _ok(m["id"], {"num_qubits": sv.num_qubits,
                             "amplitudes": [{"re": float(a.real), "im": float(a.imag)} for a in sv.data],
                             "probabilities": sv.probabilities.tolist()})

    def cmd_get_result(self, m):
        if self.last_result is None:
            return _err(m["id"], "No simulation result")
        r = self.last_result
        return _ok(m["id"], {"measurement_counts": r.measurement_counts, "num_shots": r.num_shots, "seed": r.seed})

    def cmd_set_noise(self, m):
        from quantum_sim.engine.noise import NoiseModel
        nd = m["params"].get("noise_model")
        if nd is None:
            return _err(m["id"], "Missing 'noise_model' param")
        self.noise_model = NoiseModel.from_dict(nd)
        return _ok(m["id"])

    def cmd_clear_noise(self, m):
        self.noise_model = None
        return _ok(m["id"])

    def cmd_get_analysis(self, m):                                                     # server.py:181-208
        from quantum_sim.engine.analysis import StateAnalysis
        if self.last_result is None:
            return _err(m["id"], "No simulation result")
        state = self.last_result.final_state
        data = {}
        for k in m["params"].get("metrics", ["fidelity", "entropy", "purity"]):
            if k == "fidelity" and self.ideal_state is not None:
                data["fidelity"] = StateAnalysis.state_fidelity(self.ideal_state.data, state.data)
            elif k == "entropy":
                data["entropy"] = StateAnalysis.von_neumann_entropy(state)
            elif k == "purity":
                data["purity"] = StateAnalysis.purity(state)
            elif k == "pauli":
                data["pauli"] = {f"q{q}": {p: StateAnalysis.pauli_expectation(state, p, q) for p in "XYZ"}
                                 for q in range(state.num_qubits)}
        return _ok(m["id"], data)

    def cmd_sweep_parameter(self, m):                                                  # server.py:210-267
        from quantum_sim.engine.simulator import Simulator
        from quantum_sim.engine.noise import NoiseModel, DepolarizingNoise
        from quantum_sim.engine.analysis import StateAnalysis
        if self.circuit is None:
            return _err(m["id"], "No circuit loaded")
        p = m["params"]
        values, shots, seed = p.get("values", [0.01, 0.05, 0.1]), p.get("shots", 0), p.get("seed")
        try:
            n_trials = max(1, int(p.get("trials", 50)))
        except (TypeError, ValueError):
            n_trials = 50
        rng = np.random.default_rng(seed)
        ideal = Simulator().run(self.circuit, shots=0, rng=np.random.default_rng(rng.integers(0, 2 ** 63))).final_state
        sweep = []
        for val in values:
            if float(val) == 0.0:
                sweep.append({"value": val, "fidelity": 1.0, "purity": 1.0})
                continue
            fid = pur = 0.0
            for _ in range(n_trials):
                model = NoiseModel()
                model.add_global_noise(DepolarizingNoise(float(val)))
                model.set_seed(int(rng.integers(0, 2 ** 63)))
                child = np.random.default_rng(rng.integers(0, 2 ** 63))
                res = Simulator(noise_model=model).run(self.circuit, shots=shots, rng=child)
                fid += StateAnalysis.state_fidelity(ideal.data, res.final_state.data)
                pur += StateAnalysis.purity(res.final_state)
            sweep.append({"value": val, "fidelity": fid / n_trials, "purity": pur / n_trials, "trials": n_trials})
        return _ok(m["id"], {"sweep": sweep})


def assert_same(got, want, path="", tol=1e-12):
    """Exact for strings / ints / bools / keys and key ORDER of count dicts; `tol` (absolute) for floats."""
    if isinstance(want, dict):
        assert isinstance(got, dict), path
        assert list(got.keys()) == list(want.keys()), f"{path}: keys {list(got.keys())[:6]} != {list(want.keys())[:6]}"
        for k in want:
            assert_same(got[k], want[k], f"{path}/{k}", tol)
    elif isinstance(want, list):
        assert isinstance(got, list) and len(got) == len(want), path
        for i, (g, w) in enumerate(zip(got, want)):
            assert_same(g, w, f"{path}[{i}]", tol)
    elif isinstance(want, float) and not isinstance(want, bool):
        assert isinstance(got, (int, float)) and abs(float(got) - want) <= tol, f"{path}: {got} != {want}"
    else:
        assert got == want and type(got) is type(want), f"{path}: {got!r} != {want!r}"
