"""Golden fixture of the indexing semirings. Run in the build container (needs oracle/_ref/libref_oracle.so, i.e. /root/reference):
     python tests/golden/make_golden_subsref.py
ref_subsref.npz: for NT = double, int64, bool the operands of SpParMat::SubsRef_SR (SpParMat.cpp:2515-2566) -- A, the boolean row
selector S and column selector T -- and the outputs of the UNMODIFIED reference's LocalHybridSpGEMM for S*A with
BoolCopy2ndSRing<NT> and (S*A)*T with BoolCopy1stSRing<NT> (Semirings.h:51-138), so that the GPU box (no /root/reference) can check
the device path and the C restatement against the real reference's output bits."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import Csc, RefOracle  # noqa: E402
from tests.util import subsref_operands  # noqa: E402

PAIRS = {"f64": (9, 10, np.float64), "i64": (11, 12, np.int64), "bool": (13, 14, np.uint8)}


def csc_arrays(prefix, M: Csc):
    return {prefix + "_shape": np.array([M.m, M.n]), prefix + "_colptr": M.colptr, prefix + "_rows": M.rows, prefix + "_vals": M.vals}


def main():
    R = RefOracle()
    arrs = {}
    for name, (sr2, sr1, dt) in PAIRS.items():
        A, S, T, ri, ci = subsref_operands(700, 640, 520, 480, 300 + sr2, dt)
        a, s, t = Csc.from_scipy(A, dt), Csc.from_scipy(S, np.uint8), Csc.from_scipy(T, np.uint8)
        sa = R.spgemm(s, a, sr2)
        sat = R.spgemm(sa, t, sr1)
        dense = A.toarray()[ri][:, ci]
        got = np.zeros(dense.shape)
        got[sat.rows, sat.cols_expanded()] = sat.vals
        assert np.array_equal(got, dense), "the reference's S*A*T is A[ri][:, ci]"
        for p, M in (("A", a), ("S", s), ("T", t), ("SA", sa), ("SAT", sat)):
            arrs.update(csc_arrays(f"{name}_{p}", M))
    np.savez_compressed(os.path.join(HERE, "ref_subsref.npz"), **arrs)
    print("written", os.path.join(HERE, "ref_subsref.npz"))


if __name__ == "__main__":
    main()
