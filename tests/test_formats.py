"""the on-disk / on-wire formats around the path go through this backend unchanged.
Goldens: tests/golden/golden_bridge.json, written by tests/golden/make_golden_bridge.py from the reference's own
CircuitSerializer (core/serialization.py), NoiseModel.to_dict (noise.py:262-298) and BridgeCommandHandler
(bridge/server.py:30-267) running on the reference engine."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "golden_bridge.json"), encoding="utf-8") as f:
        return json.load(f)


def test_qsim_files_round_trip(golden, tmp_path):
    """`.qsim` text written by the reference loads into this repo's data model and is written back byte for byte
    (json.dump(indent=2, ensure_ascii=False), serialization.py:19-22)."""
    from quantum_sim.engine.circuit import QuantumCircuit
    for name, text in golden["files"].items():
        data = json.loads(text)
        qc = QuantumCircuit.from_dict(data)
        assert qc.to_dict() == data, name
        assert json.dumps(qc.to_dict(), indent=2, ensure_ascii=False) == text, name
        assert qc.num_qubits == data["num_qubits"] and qc.gate_count() == len(data["gates"])
        p = tmp_path / name
        p.write_text(json.dumps(qc.to_dict(), indent=2, ensure_ascii=False), encoding="utf-8")
        assert QuantumCircuit.from_dict(json.loads(p.read_text(encoding="utf-8"))).to_dict() == data


def test_noise_model_dict_round_trip(golden):
    from quantum_sim.engine.noise import NoiseModel
    for name, d in golden["noise_models"].items():
        nm = NoiseModel.from_dict(d)
        assert nm.to_dict() == d, name
        assert json.dumps(nm.to_dict()) == json.dumps(d), name             # key order too
    nm = NoiseModel.from_dict(golden["noise_models"]["mixed"])
    assert nm.readout_error.p01 == 0.02 and nm.readout_error.p10 == 0.05
    with pytest.raises(KeyError):
        NoiseModel.from_dict({"global": [{"type": "NoSuchNoise", "probability": 0.1}]})


def test_wire_codec_matches_protocol(golden):
    """encode/decode of the replay helper reproduce the golden messages (newline-terminated JSON objects)."""
    from bridge_replay import decode, encode
    for t in golden["transcript"]:
        if "local" in t:
            continue
        for side in ("request", "response"):
            raw = encode(t[side])
            assert raw.endswith(b"\n") and raw.count(b"\n") == 1
            assert decode(raw) == t[side]


@pytest.mark.gpu
def test_bridge_transcript_on_gpu_backend(golden):
    """Every request of the recorded session, replayed on the CUDA backend, gives the reference's response:
    counts (and their key order), seeds and error texts exactly; amplitudes, probabilities and metrics to 1e-12."""
    from bridge_replay import ReplayHandler, assert_same, decode, encode
    h = ReplayHandler()
    n_checked = 0
    for t in golden["transcript"]:
        if t.get("local") == "seed_noise":
            h.noise_model.set_seed(t["seed"])
            continue
        resp = h.handle(decode(encode(t["request"])))
        got = decode(encode(resp))                          # through the wire codec, like the client would see it
        tol = 1e-9 if t["request"]["action"] == "get_analysis" else 1e-12   # entropies of ~1e-16 eigenvalues
        assert_same(got, t["response"], t["request"]["action"] + "#" + t["request"]["id"], tol)
        n_checked += 1
    assert n_checked == sum(1 for t in golden["transcript"] if "request" in t)
