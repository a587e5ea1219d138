"""GPU: the indexing semirings (BoolCopy2ndSRing / BoolCopy1stSRing, Semirings.h:51-138) against the committed outputs of the
unmodified reference (tests/golden/ref_subsref.npz; the GPU box has no /root/reference). Bit for bit."""
import os

import numpy as np
import pytest

import combblas_b200 as cb
from oracle.oracle import Csc
from tests.util import assert_same

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name, prefix):
    z = np.load(os.path.join(GOLD, name))
    m, n = z[prefix + "_shape"]
    return Csc(int(m), int(n), z[prefix + "_colptr"], z[prefix + "_rows"], z[prefix + "_vals"])


def dcsc_of(c):
    return cb.SpDCCols.from_csc(c.m, c.n, c.colptr, c.rows, c.vals, np.int64)


@pytest.mark.parametrize("name,sr2,sr1", [("f64", 9, 10), ("i64", 11, 12), ("bool", 13, 14)])
def test_committed_subsref_outputs(ctx, name, sr2, sr1):
    f = "ref_subsref.npz"
    A, S, T = load(f, name + "_A"), load(f, name + "_S"), load(f, name + "_T")
    sa = cb.LocalHybridSpGEMM(ctx, sr2, dcsc_of(S), dcsc_of(A))
    assert_same(sa, load(f, name + "_SA"), sr2)
    sat = cb.LocalHybridSpGEMM(ctx, sr1, dcsc_of(load(f, name + "_SA")), dcsc_of(T))
    assert_same(sat, load(f, name + "_SAT"), sr1)
