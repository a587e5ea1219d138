"""CPU: the numpy restatement of MCLPruneRecoverySelect (oracle/oracle.py, ParFriends.h:186-354) is pinned against the
unmodified reference (oracle/_ref) and against committed reference outputs (tests/golden/ref_mcl_prune.npz)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.oracle import Csc, kselect1, mcl_prune_recovery_select

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mcl_prune.npz")

# (m, n, hard, select, recover, pct, dtype): chosen so that the recover, select and recover-after-select rules all fire
CASES = [
    (400, 160, 2e-2, 10, 14, 0.9, np.float64),
    (400, 160, 1e-3, 6, 0, 0.9, np.float64),
    (300, 200, 5e-2, 8, 12, 0.5, np.float64),
    (600, 90, 1e-4, 1100, 1400, 0.9, np.float64),
    (400, 160, 0.05, 0, 14, 0.9, np.float64),
    (400, 160, 0.2, 3, 5, 0.99, np.float64),
    (400, 160, 2e-2, 10, 14, 0.9, np.float32),
]


def skewed_stochastic(m, n, seed, dtype):
    """column-stochastic matrix with column lengths from 0 to ~m/2 (what an MCL expansion slab looks like)"""
    rng = np.random.default_rng(seed)
    lens = np.minimum(m, (rng.pareto(1.2, n) * 6).astype(np.int64))
    lens[:3] = [0, 1, m // 2]
    rows = np.concatenate([np.sort(rng.choice(m, int(k), replace=False)) for k in lens]) if lens.sum() else np.zeros(0, np.int64)
    cols = np.repeat(np.arange(n), lens)
    vals = rng.random(len(rows)) ** 3 + 1e-6
    M = sp.csc_matrix((vals, (rows, cols)), shape=(m, n))
    s = np.asarray(M.sum(0)).ravel()
    s[s == 0] = 1
    M = sp.csc_matrix(M @ sp.diags(1 / s))
    M.sort_indices()
    return Csc.from_scipy(M, dtype)


def same(a: Csc, b: Csc):
    return a.nnz == b.nnz and np.array_equal(a.colptr, b.colptr) and np.array_equal(a.rows, b.rows) and np.array_equal(a.vals, b.vals)


def test_kselect1_rule():
    v = np.array([0.5, 0.1, 0.9, 0.3])
    assert kselect1(v, 1) == 0.9 and kselect1(v, 3) == 0.3 and kselect1(v, 4) == 0.1
    assert kselect1(v, 7) == 0.1  # fewer than k entries: the smallest one (SpParMat.cpp:1683)
    assert kselect1(v[:0], 3) == np.finfo(np.float64).tiny  # empty: numeric_limits::min() (:1681)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_restatement_equals_reference(ref_oracle, case):
    m, n, hard, select, recover, pct, dt = CASES[case]
    A = skewed_stochastic(m, n, 100 + case, dt)
    want = ref_oracle.mcl_prune(A, hard, select, recover, pct, sr=0 if dt == np.float64 else 1)
    got, _ = mcl_prune_recovery_select(A, hard, select, recover, pct)
    assert same(got, want)
    if case == 0:
        assert 0 < got.nnz < A.nnz


def test_restatement_equals_committed_reference_outputs():
    z = np.load(GOLDEN)
    for case in range(len(CASES)):
        m, n, hard, select, recover, pct, dt = CASES[case]
        A = skewed_stochastic(m, n, 100 + case, dt)
        assert np.array_equal(A.vals, z[f"in_vals_{case}"]), "the seeded input changed: regenerate the fixture"
        got, _ = mcl_prune_recovery_select(A, hard, select, recover, pct)
        assert np.array_equal(got.colptr, z[f"colptr_{case}"]) and np.array_equal(got.rows, z[f"rows_{case}"])
        assert np.array_equal(got.vals, z[f"vals_{case}"])


def test_every_rule_fires():
    """the seeded cases exercise recover, select and recover-after-select"""
    fired = set()
    for case in range(len(CASES)):
        m, n, hard, select, recover, pct, dt = CASES[case]
        A = skewed_stochastic(m, n, 100 + case, dt)
        hard_t, pct_t = dt(hard), dt(pct)
        for j in range(A.n):
            v = A.vals[A.colptr[j]:A.colptr[j + 1]]
            pr = v[v > hard_t]
            if len(pr) < recover and len(v) > len(pr) and pr.sum(dtype=dt) < pct_t:
                fired.add("recover")
            elif select > 0 and len(pr) > select:
                fired.add("select")
                if recover > 0:
                    t = kselect1(v, select)
                    sel = v[~(v < t)]
                    if len(sel) < recover and sel.sum(dtype=dt) < pct_t:
                        fired.add("recover_after_select")
    assert fired == {"recover", "select", "recover_after_select"}, fired


def test_restatement_fuzz_against_reference(ref_oracle):
    """80 seeded random (shape, parameter) draws, including k larger than any column, zero / negative thresholds and ties"""
    rng = np.random.default_rng(2026)
    for it in range(80):
        m, n = int(rng.integers(1, 400)), int(rng.integers(3, 120))
        A = skewed_stochastic(max(m, 8), n, 1000 + it, np.float64)
        if it % 4 == 0:  # ties: few distinct values
            A = Csc(A.m, A.n, A.colptr, A.rows, np.round(A.vals * 8) / 8 + (rng.integers(0, 2, A.nnz) * 0.125))
        hard = float(rng.choice([0.0, -1.0, 1e-4, 1e-2, 0.05, 0.3]))
        select = int(rng.choice([0, 1, 2, 5, 17, 1100]))
        recover = int(rng.choice([0, 1, 3, 9, 25, 1400]))
        pct = float(rng.choice([0.0, 0.3, 0.9, 1.5, 1e9]))
        want = ref_oracle.mcl_prune(A, hard, select, recover, pct)
        got, _ = mcl_prune_recovery_select(A, hard, select, recover, pct)
        assert same(got, want), (it, m, n, hard, select, recover, pct, got.nnz, want.nnz)
