"""Host-side helpers that replaced per-shot / per-trial Python loops (CPU)."""

import numpy as np
import pytest

from qsb.distributed import merge_counts_in_shot_order


def per_shot_loop(indices, n):
    """The reference's way of building the counts dict (simulator.py:144-145)."""
    out = {}
    for i in indices.tolist():
        key = format(i, f"0{n}b")
        out[key] = out.get(key, 0) + 1
    return out


@pytest.mark.parametrize("n,shots", [(1, 9), (3, 1000), (8, 300), (9, 4000), (16, 16384), (16, 500), (20, 5000), (24, 3000), (30, 64)])
def test_counts_dict_matches_the_per_shot_loop(n, shots):
    rng = np.random.default_rng(1000 * n + shots)
    idx = rng.integers(0, 2 ** n, shots)
    got, ref = merge_counts_in_shot_order(idx, n), per_shot_loop(idx, n)
    assert got == ref
    assert list(got) == list(ref)                                   # insertion order = order of first occurrence
    assert all(type(k) is str and len(k) == n for k in got) and all(type(v) is int for v in got.values())


def test_counts_dict_edge_cases():
    assert merge_counts_in_shot_order(np.zeros(0, dtype=np.int64), 5) == {}
    assert merge_counts_in_shot_order(np.array([3, 3, 3]), 2) == {"11": 3}
    assert list(merge_counts_in_shot_order(np.array([2, 0, 2, 1]), 2).items()) == [("10", 2), ("00", 1), ("01", 1)]
    assert merge_counts_in_shot_order([5], 3) == {"101": 1}         # a plain list is accepted
