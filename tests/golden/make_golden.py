"""Generates the committed golden fixtures. Run in the build container (needs /root/reference and the
reference oracle oracle/_ref/libref_oracle.so):  python tests/golden/make_golden.py

 * bcsstk01_*.npz : the reference tree's own known-answer vector, 3DSpGEMM/matlab/bcsstk01.mtx squared ==
   3DSpGEMM/matlab/C.mtx (written by MATLAB, multwrite.m). Stored as CSC arrays.
 * ref_*.npz      : outputs of the UNMODIFIED reference (LocalHybridSpGEMM / Mult_AnXBn_Synch / MultiwayMerge)
   on small seeded inputs for every semiring, so that the GPU box (which has no /root/reference) can still
   check against the real reference's output bits.
"""
import os
import sys

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import (Csc, RefOracle, SR_DTYPES, REF_LOCAL_HYBRID, REF_LOCAL_HASH_SORTED, REF_DIST_SYNCH)  # noqa: E402
from tests.util import random_pair  # noqa: E402

REF = "/root/reference"


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name), **arrs)


def csc_arrays(prefix, M: Csc):
    return {prefix + "_shape": np.array([M.m, M.n]), prefix + "_colptr": M.colptr, prefix + "_rows": M.rows, prefix + "_vals": M.vals}


def main():
    A = Csc.from_scipy(sio.mmread(f"{REF}/3DSpGEMM/matlab/bcsstk01.mtx", spmatrix=True))
    Cm = Csc.from_scipy(sio.mmread(f"{REF}/3DSpGEMM/matlab/C.mtx", spmatrix=True))
    save("bcsstk01_squared.npz", **csc_arrays("A", A), **csc_arrays("C", Cm))
    R = RefOracle()
    for sr in range(9):
        A, B = random_pair(180, 150, 200, 0.06, 0.05, 100 + sr, SR_DTYPES[sr])
        a, b = Csc.from_scipy(A, SR_DTYPES[sr][0]), Csc.from_scipy(B, SR_DTYPES[sr][1])
        c = R.spgemm(a, b, sr, REF_LOCAL_HYBRID)
        c2 = R.spgemm(a, b, sr, REF_LOCAL_HASH_SORTED)
        assert np.array_equal(c.rows, c2.rows)
        # three partial products merged with the heap MultiwayMerge
        parts = []
        for i in range(3):
            Ai, _ = random_pair(180, 150, 200, 0.05, 0.05, 1000 + 10 * sr + i, SR_DTYPES[sr])
            parts.append(R.spgemm(Csc.from_scipy(Ai, SR_DTYPES[sr][0]), b, sr, REF_LOCAL_HYBRID))
        mg = R.merge(parts, sr, hash=False)
        arrs = {**csc_arrays("A", a), **csc_arrays("B", b), **csc_arrays("C", c), **csc_arrays("M", mg)}
        for i, p in enumerate(parts):
            arrs.update(csc_arrays(f"P{i}", p))
        save(f"ref_sr{sr}.npz", **arrs)
    # distributed driver at P=1 (Mult_AnXBn_Synch) on a small R-MAT-like input
    A, _ = random_pair(256, 256, 256, 0.04, 0.04, 77, SR_DTYPES[0])
    a = Csc.from_scipy(A, np.float64)
    c = R.spgemm(a, a, 0, REF_DIST_SYNCH)
    save("ref_synch_sr0.npz", **csc_arrays("A", a), **csc_arrays("C", c))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
