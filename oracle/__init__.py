"""TEST INFRASTRUCTURE ONLY: loaders for the CPU oracle (port) and the compiled reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this package.
Nothing under vif_b200/ or include/ does (checked by tests/test_abi.py::test_no_oracle_in_product).
"""
from .oracle import (OracleGrid, have_ref, oracle_qxmatch, oracle_grid, oracle_ring,  # noqa: F401
                     oracle_field_area_bbox, oracle_angdist, ref_qxmatch, ref_ring,
                     ref_field_area_bbox, ref_angdist, build, NPOS,
                     oracle_angdist_vec, oracle_angdist_less, oracle_qdist,
                     ref_angdist_vec, ref_angdist_less, ref_qdist)
