"""Shared test plumbing: markers, import paths, golden fixtures."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantum-simulator_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        j = json.load(f)
    a = np.load(os.path.join(GOLDEN_DIR, "golden.npz"))
    return j, a


def as_gates(raw):
    """JSON round-trips tuples as lists; normalise to (name, targets, params, column)."""
    return [(g[0], list(g[1]), list(g[2]), g[3]) for g in raw]


def as_noise(raw):
    if raw is None:
        return None
    return {"global": [tuple(c) for c in raw.get("global", [])],
            "gate": {k: [tuple(c) for c in v] for k, v in raw.get("gate", {}).items()},
            "readout": tuple(raw["readout"]) if raw.get("readout") else None}
