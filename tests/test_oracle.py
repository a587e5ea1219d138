"""CPU: pins the oracle. The C restatement (oracle/spgemm_oracle.c) must reproduce
 (1) the reference tree's own known-answer vector bcsstk01^2 == C.mtx (3DSpGEMM/matlab),
 (2) the committed outputs of the UNMODIFIED reference for every semiring (tests/golden/ref_sr*.npz),
 (3) where the compiled reference is present (build container), the reference itself on fresh seeded inputs,
     including its unsorted hash-table order and its symbolic counts."""
import os

import numpy as np
import pytest

from oracle.oracle import (Csc, SR_DTYPES, SR_NAMES, REF_LOCAL_HYBRID, REF_LOCAL_HASH_SORTED, REF_LOCAL_HASH_UNSORTED,
                           REF_LOCAL_HEAP, REF_DIST_SYNCH, REF_DIST_DOUBLEBUFF, REF_DIST_MEMEFF_HASH, REF_DIST_MEMEFF_HEAP,
                           REF_DIST_SUMMA3D)
from tests.util import random_pair, to_csc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name, prefix):
    z = np.load(os.path.join(GOLD, name))
    m, n = z[prefix + "_shape"]
    return Csc(int(m), int(n), z[prefix + "_colptr"], z[prefix + "_rows"], z[prefix + "_vals"])


def same_pattern(a, b):
    return a.nnz == b.nnz and np.array_equal(a.colptr, b.colptr) and np.array_equal(a.rows, b.rows)


def test_port_matches_reference_known_answer(port_oracle):
    A, G = load("bcsstk01_squared.npz", "A"), load("bcsstk01_squared.npz", "C")
    c = port_oracle.spgemm(A, A, 0)
    assert same_pattern(c, G) and c.nnz == 1292
    assert np.max(np.abs(c.vals - G.vals) / np.abs(G.vals)) < 1e-12


@pytest.mark.parametrize("sr", range(9))
def test_port_matches_committed_reference_outputs(port_oracle, sr):
    f = f"ref_sr{sr}.npz"
    A, B, Cg = load(f, "A"), load(f, "B"), load(f, "C")
    c = port_oracle.spgemm(A, B, sr)
    assert same_pattern(c, Cg)
    if sr in (0, 1, 6):  # LocalHybridSpGEMM mixes heap and hash columns: floating sums may be ordered differently
        assert np.allclose(c.vals, Cg.vals, rtol=1e-12 if sr != 1 else 1e-5, atol=0)
    else:
        assert np.array_equal(c.vals, Cg.vals)
    parts = [load(f, f"P{i}") for i in range(3)]
    mg = port_oracle.merge(parts, sr)
    Mg = load(f, "M")
    assert same_pattern(mg, Mg)
    if sr in (0, 1, 6):
        assert np.allclose(mg.vals, Mg.vals, rtol=1e-12 if sr != 1 else 1e-5, atol=0)
    else:
        assert np.array_equal(mg.vals, Mg.vals)


def test_port_matches_committed_synch_output(port_oracle):
    A, Cg = load("ref_synch_sr0.npz", "A"), load("ref_synch_sr0.npz", "C")
    c = port_oracle.spgemm(A, A, 0)
    assert same_pattern(c, Cg) and np.allclose(c.vals, Cg.vals, rtol=1e-12, atol=0)


@pytest.mark.parametrize("sr", range(9))
def test_port_is_bit_identical_to_reference_hash_kernel(port_oracle, ref_oracle, sr):
    A, B = random_pair(300, 220, 260, 0.05, 0.04, 7 + sr, SR_DTYPES[sr])
    a, b = to_csc(A, SR_DTYPES[sr][0]), to_csc(B, SR_DTYPES[sr][1])
    # sorted: the reference hash kernel and the restatement agree bit for bit (same accumulation order)
    r = ref_oracle.spgemm(a, b, sr, REF_LOCAL_HASH_SORTED)
    p = port_oracle.spgemm(a, b, sr)
    assert same_pattern(r, p) and np.array_equal(r.vals, p.vals)
    # unsorted: even the reference's hash-table emission order is reproduced
    ru = ref_oracle.spgemm(a, b, sr, REF_LOCAL_HASH_UNSORTED, canonical=False)
    pu = port_oracle.spgemm(a, b, sr, sort=False)
    assert np.array_equal(ru.rows, pu.rows) and np.array_equal(ru.vals, pu.vals)
    # symbolic
    fl, nz = ref_oracle.symbolic(a, b, sr)
    fl2, nz2 = port_oracle.symbolic(a, b)
    ne = np.diff(b.colptr) > 0
    assert np.array_equal(fl, fl2[ne]) and np.array_equal(nz, nz2[ne])


def test_reference_kernels_agree_with_each_other(ref_oracle):
    """the cross-checks the reference's own tests make: heap == hash == hybrid; Synch == DoubleBuff == MemEfficient ==
    3D (layers=1) (ReleaseTests/MultTest.cpp:162-182, SpGEMM3DTest.cpp:73-100)"""
    A, B = random_pair(200, 200, 200, 0.05, 0.05, 3, SR_DTYPES[0])
    a, b = to_csc(A, np.float64), to_csc(B, np.float64)
    base = ref_oracle.spgemm(a, b, 0, REF_LOCAL_HYBRID)
    for routine in (REF_LOCAL_HASH_SORTED, REF_LOCAL_HEAP, REF_DIST_SYNCH, REF_DIST_DOUBLEBUFF, REF_DIST_MEMEFF_HASH,
                    REF_DIST_MEMEFF_HEAP, REF_DIST_SUMMA3D):
        c = ref_oracle.spgemm(a, b, 0, routine, phases=3 if routine in (REF_DIST_MEMEFF_HASH, REF_DIST_MEMEFF_HEAP) else 1)
        assert same_pattern(base, c), routine
        assert np.allclose(base.vals, c.vals, rtol=1e-12, atol=0), routine


@pytest.mark.parametrize("sr", [0, 2, 3, 4, 5, 8])
def test_port_merge_matches_reference_merges(port_oracle, ref_oracle, sr):
    _, B = random_pair(150, 120, 140, 0.06, 0.06, 50 + sr, SR_DTYPES[sr])
    b = to_csc(B, SR_DTYPES[sr][1])
    parts = []
    for i in range(4):
        Ai, _ = random_pair(150, 120, 140, 0.06, 0.06, 500 + 7 * sr + i, SR_DTYPES[sr])
        parts.append(port_oracle.spgemm(to_csc(Ai, SR_DTYPES[sr][0]), b, sr))
    heap = ref_oracle.merge(parts, sr, hash=False)
    mine = port_oracle.merge(parts, sr)
    assert same_pattern(heap, mine)
    if sr == 0:
        assert np.allclose(heap.vals, mine.vals, rtol=1e-12, atol=0)
    else:
        assert np.array_equal(heap.vals, mine.vals)
    if sr != 5:  # the reference's MultiwayMergeHash does not terminate for bool values (see oracle/ref_oracle.cpp)
        hashed = ref_oracle.merge(parts, sr, hash=True)
        assert same_pattern(hashed, mine) and np.array_equal(hashed.vals, mine.vals)


def test_empty_and_degenerate_inputs(port_oracle):
    z = Csc(5, 7, np.zeros(8, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64))
    a = Csc.from_coo(4, 5, [0, 3], [1, 4], np.array([2.0, 3.0]))
    c = port_oracle.spgemm(a, z, 0)
    assert c.nnz == 0 and c.m == 4 and c.n == 7
    # explicit zeros are kept (no value-based dropping anywhere on the path)
    x = Csc.from_coo(2, 2, [0, 0], [0, 1], np.array([1.0, -1.0]))
    y = Csc.from_coo(2, 1, [0, 1], [0, 0], np.array([1.0, 1.0]))
    c = port_oracle.spgemm(x, y, 0)
    assert c.nnz == 1 and c.vals[0] == 0.0


@pytest.mark.parametrize("pair", [(9, 10, np.float64), (11, 12, np.int64), (13, 14, np.uint8)])
def test_bool_copy_semirings_subsref(port_oracle, pair):
    """BoolCopy2ndSRing / BoolCopy1stSRing (Semirings.h:51-138), the pair SpParMat::SubsRef_SR multiplies with: S*A*T with
    boolean selectors equals A[ri][:, ci] entry for entry (explicit zeros and signs kept); the restatement equals the compiled
    reference where that is present; a product in which an output would receive two values fails, as the reference's add throws"""
    from tests.util import subsref_operands
    from oracle.oracle import RefOracle

    sr2, sr1, dt = pair
    A, S, T, ri, ci = subsref_operands(300, 260, 180, 140, 5 + sr2, dt)
    a, s_, t_ = to_csc(A, dt), to_csc(S, np.uint8), to_csc(T, np.uint8)
    sa = port_oracle.spgemm(s_, a, sr2)
    sat = port_oracle.spgemm(sa, t_, sr1)
    dense = A.toarray()[ri][:, ci]
    got = np.zeros(dense.shape)
    got[sat.rows, sat.cols_expanded()] = sat.vals
    assert np.array_equal(got, dense)
    if RefOracle.available():
        ref = RefOracle()
        rsa = ref.spgemm(s_, a, sr2)
        rsat = ref.spgemm(rsa, t_, sr1)
        assert same_pattern(sa, rsa) and np.array_equal(sa.vals, rsa.vals)
        assert same_pattern(sat, rsat) and np.array_equal(sat.vals, rsat.vals)
    # a selector row with two entries on rows of A that share a column: the output needs add -> "Add should not happen"
    X = sp_two_hits(A, 300)
    with pytest.raises(RuntimeError):
        port_oracle.spgemm(to_csc(X, np.uint8), a, sr2)


def sp_two_hits(A, m):
    """a 1 x m boolean row with two entries on rows of A that share a column: S*A then needs an add"""
    import scipy.sparse as sp

    Ac = A.tocsc()
    col = int(np.argmax(np.diff(Ac.indptr)))
    r0, r1 = Ac.indices[Ac.indptr[col]], Ac.indices[Ac.indptr[col] + 1]
    X = sp.coo_matrix((np.ones(2), (np.zeros(2, int), np.array([r0, r1]))), shape=(1, m)).tocsc()
    X.sort_indices()
    return X


@pytest.mark.parametrize("name,sr2,sr1", [("f64", 9, 10), ("i64", 11, 12), ("bool", 13, 14)])
def test_port_matches_committed_subsref_outputs(port_oracle, name, sr2, sr1):
    """the C restatement against the committed outputs of the unmodified reference for the indexing semirings
    (tests/golden/ref_subsref.npz, written by tests/golden/make_golden_subsref.py): bit for bit"""
    f = "ref_subsref.npz"
    A, S, T = load(f, name + "_A"), load(f, name + "_S"), load(f, name + "_T")
    SA, SAT = load(f, name + "_SA"), load(f, name + "_SAT")
    sa = port_oracle.spgemm(S, A, sr2)
    assert same_pattern(sa, SA) and np.array_equal(sa.vals, SA.vals)
    sat = port_oracle.spgemm(sa, T, sr1)
    assert same_pattern(sat, SAT) and np.array_equal(sat.vals, SAT.vals)
