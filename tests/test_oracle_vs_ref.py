"""Pins the oracle port against the reference itself on seeded random inputs (larger than the
fixtures). Needs oracle/_ref (built in the authoring container from /root/reference; the binary
travels to the GPU box). Skipped where neither exists."""
import os

import numpy as np
import pytest

from vif_b200 import catalogs as G

REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                      "libqxmatch_ref.so")
pytestmark = pytest.mark.skipif(not (os.path.exists(REF_SO) or os.path.isdir("/root/reference/include/vif")),
                                reason="compiled reference not available")


def same(a, b):
    for k in ("id", "rid"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("d", "rd"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("kw", [dict(), dict(nth=3), dict(nth=5, self=True), dict(no_mirror=True, nth=2)],
                         ids=["nth1", "nth3", "self5", "nomirror"])
def test_binned_random(oracle, kw):
    n = 20000
    a1, b1 = G.c2(31, n)
    a2, b2 = (a1, b1) if kw.get("self") else G.c2(32, n)
    same(oracle.oracle_qxmatch(a1, b1, a2, b2, thread=4, **kw), oracle.ref_qxmatch(a1, b1, a2, b2, **kw))


def test_brute_random(oracle):
    a1, b1 = G.fullsky(33, 1500)
    a2, b2 = G.fullsky(34, 1300)
    for kw in (dict(brute_force=True), dict(brute_force=True, nth=3, self=False)):
        same(oracle.oracle_qxmatch(a1, b1, a2, b2, thread=3, **kw), oracle.ref_qxmatch(a1, b1, a2, b2, **kw))


def test_threaded_reference_equals_single_thread_binned(oracle):
    """The reference's binned result does not depend on `thread`."""
    a1, b1 = G.c1(35, 5000)
    a2, b2 = G.c1(36, 5000)
    same(oracle.ref_qxmatch(a1, b1, a2, b2, nth=2, thread=1), oracle.ref_qxmatch(a1, b1, a2, b2, nth=2, thread=4))


def test_scalar_helpers(oracle):
    for box in ([0, 1, 0, 1], [0, 10, 60, 70], [3, 3, 1, 2], [350, 360, -5, 5]):
        assert oracle.oracle_field_area_bbox(*box) == oracle.ref_field_area_bbox(*box)
    for p in ([0.1, 0.2, 0.3, 0.4], [359.9, 89, 0.1, 89.5]):
        assert oracle.oracle_angdist(*p) == oracle.ref_angdist(*p)
