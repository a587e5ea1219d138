"""Writes tests/golden/ref_mcl_prune.npz: outputs of the UNMODIFIED reference's MCLPruneRecoverySelect (ParFriends.h:186,
through oracle/_ref/libref_oracle.so, built from /root/reference) on the seeded inputs of tests/test_prune_oracle.py.
Run in the build container:  python tests/golden/make_golden_prune.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle  # noqa: E402
from tests.test_prune_oracle import CASES, skewed_stochastic  # noqa: E402

ref = RefOracle()
out = {}
for case, (m, n, hard, select, recover, pct, dt) in enumerate(CASES):
    A = skewed_stochastic(m, n, 100 + case, dt)
    C = ref.mcl_prune(A, hard, select, recover, pct, sr=0 if dt == np.float64 else 1)
    out[f"in_vals_{case}"] = A.vals
    out[f"colptr_{case}"] = C.colptr
    out[f"rows_{case}"] = C.rows
    out[f"vals_{case}"] = C.vals
    print(case, A.nnz, "->", C.nnz)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_mcl_prune.npz"), **out)
