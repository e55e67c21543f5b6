"""ThreadSanitizer pass over the executor protocol (CPU only).

The CUDA executor and the host emulator share qsb_exec.cuh; the emulator maps CUDA named barriers /
barrier.cluster onto std::barrier one to one.  Running representative programs (cluster remaps, rank-slot
flushes, state-dependent Kraus draws, in-place LOAD+STORE, snapshots) under TSan catches missing barriers
that a GPU run can hide by timing."""

import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from oracle import qsim_oracle as O
from qsb.compiler import Lowering
from qsb.lowering import lower_circuit
from qsb.workloads import layered_circuit
from quantum_sim.engine.circuit import QuantumCircuit, GateInstance
from quantum_sim.engine.gate_registry import GateRegistry

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu_tsan_main.cpp")
BIN = os.path.join(HERE, "emu", "emu_tsan.bin")
HDRS = [os.path.join(ROOT, "quantum-simulator_b200", "csrc", "qsb_exec.cuh"), os.path.join(ROOT, "include", "qsb.h"),
        os.path.join(HERE, "emu", "emu_exec.cpp")]


def _build():
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < newest:
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++20", "-fsanitize=thread", "-pthread",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "quantum-simulator_b200", "csrc"),
                               "-I" + os.path.join(HERE, "emu"), "-o", BIN, SRC])
    return BIN


def _dump(path, prog, W, count=1, states=None, uniforms=None, params=None, out_of_place=False):
    dim = 1 << prog.n
    flags = 2 | (1 if states is not None else 0) | (4 if prog.normalize else 0)
    st = np.zeros((count, dim), dtype=np.complex128) if states is None else np.ascontiguousarray(states, dtype=np.complex128)
    with open(path, "wb") as f:
        def w(arr):
            arr = np.ascontiguousarray(arr)
            f.write(struct.pack("<q", arr.size if arr.dtype != prog.ops.dtype else len(arr)))
            f.write(arr.tobytes())
        w(np.array([prog.n, prog.m, W, prog.load_perm, prog.store_perm, prog.n_snapshots, flags, count, prog.ops_stride, int(out_of_place)],
                   dtype=np.int64))
        w(prog.ops)
        w(np.asarray(prog.cdata, dtype=np.float64))
        w(np.asarray(prog.idata, dtype=np.int32))
        w(np.zeros(0) if uniforms is None else np.asarray(uniforms, dtype=np.float64))
        w(np.zeros(0) if params is None else np.asarray(params, dtype=np.float64))
        w(st.reshape(-1))


def _circuit(n, gates):
    qc = QuantumCircuit(n)
    for g in gates:
        qc.add_gate(GateInstance(g[0], list(g[1]), list(g[2]), g[3]))
    return qc


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_executor_protocol_is_race_free(tmp_path):
    try:
        exe = _build()
    except subprocess.CalledProcessError:
        pytest.skip("ThreadSanitizer runtime not available")
    reg = GateRegistry.instance()
    rng = np.random.default_rng(5)
    files = []
    # 1. in-place 1-qubit gate on a cluster of 4 (LOAD + STORE with different bit permutations)
    lw = Lowering(5)
    lw.matrix(np.array([[0.6, 0.8j], [0.8j, 0.6]]), [4])
    psi = rng.normal(size=32) + 1j * rng.normal(size=32)
    files.append(str(tmp_path / "inplace.bin"))
    _dump(files[-1], lw.finish(3), 8, states=psi[None])
    # 2. noisy trajectories on clusters of 1, 2, 8 with snapshots (remaps, rank-slot flushes, AD reductions)
    noise = {"global": [("depolarizing", 0.2), ("amplitude_damping", 0.3)], "gate": {}}
    for n, m in ((6, 6), (6, 5), (7, 4)):
        gates = layered_circuit(n, 4, 11)
        prog, _ = lower_circuit(n, _circuit(n, gates).get_ordered_gates(), reg,
                                lambda name: [(k, p, None) for k, p in O.channels_for(noise, name)],
                                record_steps=True, local_bits=m)
        files.append(str(tmp_path / f"noisy_{n}_{m}.bin"))
        _dump(files[-1], prog, 8, count=3, uniforms=rng.random((3, prog.n_draws)))
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66")
    r = subprocess.run([exe] + files, capture_output=True, text=True, env=env, timeout=600)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr[:2000])
