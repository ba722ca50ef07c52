"""Rigour of the bounds of ring_nn_filter_kernel, checked on the CPU: the numpy model of its decision
logic (tests/ring_filter_model.py) against the oracle = the reference's own binned search. A row the
filter accepts must carry exactly the id the reference returns -- including where the reference stops
after ring 0 and where its result is not the true nearest neighbour (tall-Dec fields)."""
import numpy as np
import pytest

from ring_filter_model import run_filter
from vif_b200 import catalogs as G


def _box(seed, n, r0, r1, d0, d1):
    return r0 + (r1 - r0)*G.uniform(seed, 0, n), d0 + (d1 - d0)*G.uniform(seed, 1, n)


def _pairs(seed, n, r0, r1, d0, d1, jitter_deg):
    """catalog 2 = catalog 1 displaced by a small astrometric error (the cross-matching use case)"""
    a, b = _box(seed, n, r0, r1, d0, d1)
    return (a, b), (a + jitter_deg*(G.uniform(seed, 2, n) - 0.5), b + jitter_deg*(G.uniform(seed, 3, n) - 0.5))


CASES = {
    "c1_box": lambda: (_box(1, 20000, 0.0, 1.0, 0.0, 1.0), _box(2, 20000, 0.0, 1.0, 0.0, 1.0), False),
    "tall_dec_field": lambda: (_box(3, 15000, 10.0, 11.0, 20.0, 70.0), _box(4, 15000, 10.0, 11.0, 20.0, 70.0), False),
    "high_dec_box": lambda: (_box(5, 15000, 100.0, 104.0, 66.0, 69.0), _box(6, 15000, 100.0, 104.0, 66.0, 69.0), False),
    "wide_strip": lambda: (_box(7, 15000, 0.0, 40.0, -1.0, 1.0), _box(8, 15000, 0.0, 40.0, -1.0, 1.0), False),
    "few_candidates": lambda: (_box(9, 20000, 50.0, 52.0, -31.0, -29.0), _box(10, 700, 50.0, 52.0, -31.0, -29.0), False),
    "many_candidates": lambda: (_box(11, 3000, 50.0, 52.0, -31.0, -29.0), _box(12, 60000, 50.0, 52.0, -31.0, -29.0), False),
    "displaced_copies": lambda: (*_pairs(13, 20000, 200.0, 201.0, 40.0, 41.0, 2.0/3600.0), False),
    # cells close to the filter's size limit (5e-3 rad): the small-angle forms and the stop-rule slack at their loosest
    "large_cells": lambda: (_box(17, 20000, 150.0, 153.3, 30.0, 33.3), _box(18, 2000, 150.0, 153.3, 30.0, 33.3), False),
    "large_cells_high_dec": lambda: (_box(19, 20000, 150.0, 153.7, 55.0, 57.05), _box(20, 2000, 150.0, 153.7, 55.0, 57.05), False),
    "self": lambda: (_box(15, 20000, 300.0, 301.5, -10.0, -9.0), None, True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_accepted_rows_carry_the_reference_id(oracle, name):
    c1, c2, self_match = CASES[name]()
    if c2 is None:
        c2 = c1
    bb1 = (c1[0].min(), c1[0].max(), c1[1].min(), c1[1].max())
    bb2 = (c2[0].min(), c2[0].max(), c2[1].min(), c2[1].max())
    grid = oracle.oracle_grid(bb1, bb2, c2[0].size)
    acc, win, ring0 = run_filter(c1[0], c1[1], c2[0], c2[1], grid, self_match)
    exp = oracle.oracle_qxmatch(c1[0], c1[1], c2[0], c2[1], self=self_match, thread=8, no_mirror=True)
    ref_id = exp["id"][0].astype(np.int64)
    assert acc.sum() > 0.9*acc.size, (name, int(acc.sum()), grid)        # the filter decides almost every row
    wrong = np.flatnonzero(acc & (win != ref_id))
    assert wrong.size == 0, (name, wrong[:10], win[wrong[:10]], ref_id[wrong[:10]], grid)
