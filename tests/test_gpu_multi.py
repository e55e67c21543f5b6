"""Multi-GPU legs that need real peers (skipped on a one-GPU box; the CPU side of the same logic runs under gloo in
tests/test_sharding_gloo.py): the sharded big state of BASELINE config 5 under torchrun, parity against the oracle."""

import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_big.py")] + args
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.skipif(_gpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("fuse", ["store", "load", "nccl"])
def test_sharded_state_on_all_gpus_matches_the_oracle(fuse):
    """Global qubits sharded over the GPUs of the box (all of them, a power of two); exchanges folded into the store of
    the preceding pass (peer-mapped TMA stores), into the load of the following pass (peer-mapped TMA loads), or done by
    the NCCL all-to-all: 24 qubits against the oracle, 28 qubits unit norm."""
    n = 1 << (_gpus().bit_length() - 1)
    extra = ["--no-fuse"] if fuse == "nccl" else ["--fuse-where", fuse]
    out = _torchrun(n, ["--check-n", "24", "--check-depth", "3", "--qubits", "28", "--depth", "6"] + extra)
    fuse = fuse != "nccl"
    assert out["world"] == n
    assert out["check"]["max_abs_err"] < 1e-12 and abs(out["check"]["norm2"] - 1.0) < 1e-12
    assert out["run"]["exchanges"] >= 1 and abs(out["run"]["norm2"] - 1.0) < 1e-10
    if fuse and out["run"]["symm_error"] is None:
        assert out["run"]["fused_exchanges_per_run"] >= 1
