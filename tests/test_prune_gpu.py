"""GPU: the device MCL pruning epilogue (cbgpu_mcl_prune, cbgpu_mat_make_col_stochastic, cbgpu_mat_inflate) and the
phased multiply with pruning (MemEfficientSpGEMM mirror) against the oracle: the numpy restatement of
MCLPruneRecoverySelect (pinned against the reference in tests/test_prune_oracle.py) and the reference itself."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import combblas_b200 as cb
from oracle.oracle import Csc, mcl_prune_recovery_select
from tests.test_prune_oracle import CASES, skewed_stochastic
from tests.util import rmat

pytestmark = pytest.mark.gpu


def to_host_block(A: Csc):
    return cb.SpDCCols.from_csc(A.m, A.n, A.colptr, A.rows, A.vals, np.int64)


def device_prune(ctx, A: Csc, hard, select, recover, pct):
    dA = ctx.upload(to_host_block(A))
    D, st = ctx.mcl_prune(dA, hard, select, recover, pct, want_stats=True)
    rows, cols, vals = ctx.download_coo(D)
    inf = D.info()
    D.free()
    dA.free()
    return rows, cols, vals, st, inf


def assert_bit_exact(got, want: Csc):
    rows, cols, vals = got[:3]
    assert len(rows) == want.nnz, f"nnz {len(rows)} vs {want.nnz}"
    assert np.array_equal(cols, want.cols_expanded()), "columns differ"
    assert np.array_equal(rows, want.rows), "rows differ"
    assert vals.dtype == want.vals.dtype and np.array_equal(vals, want.vals), "values differ"


@pytest.mark.parametrize("case", range(len(CASES)))
def test_prune_matches_oracle(ctx, case):
    m, n, hard, select, recover, pct, dt = CASES[case]
    A = skewed_stochastic(m, n, 100 + case, dt)
    want, _ = mcl_prune_recovery_select(A, hard, select, recover, pct)
    got = device_prune(ctx, A, hard, select, recover, pct)
    assert_bit_exact(got, want)
    st, inf = got[3], got[4]
    assert st.nnz_in == A.nnz and st.nnz_out == want.nnz == inf.nnz
    assert inf.nzc == int(np.count_nonzero(np.diff(want.colptr)))
    if case == 0:
        assert st.cols_recovered > 0 and st.cols_selected > 0 and st.cols_recovered_after_select > 0


def test_prune_edge_cases(ctx):
    # empty block
    E = Csc(50, 40, np.zeros(41, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64))
    got = device_prune(ctx, E, 1e-3, 5, 7, 0.9)
    assert len(got[0]) == 0 and got[4].nnz == 0 and got[4].nzc == 0
    # ties, negative values, columns shorter than k, a column entirely below the hard threshold
    rng = np.random.default_rng(3)
    m, n = 6000, 12
    cols, rows, vals = [], [], []
    pool = np.array([-0.5, -0.0, 0.0, 0.01, 0.02, 0.02, 0.3, 0.3, 0.3, 1.0])
    for j in range(n):
        k = [0, 1, 3, 7, 8, 50, 500, 5000, 40, 40, 2, 6000][j]
        r = np.sort(rng.choice(m, k, replace=False))
        v = rng.choice(pool, k) if j != 8 else np.full(k, 1e-6)
        if j == 9:
            v = np.full(k, 0.25)  # all equal: every entry is the k-th largest
        rows.append(r)
        cols.append(np.full(k, j))
        vals.append(v)
    M = Csc.from_coo(m, n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals).astype(np.float64))
    for hard, select, recover, pct in [(0.015, 7, 9, 0.9), (0.015, 7, 9, 1e9), (0.5, 3, 0, 0.9), (-1.0, 4, 4, 0.1), (0.015, 0, 9, 100.0)]:
        want, _ = mcl_prune_recovery_select(M, hard, select, recover, pct)
        assert_bit_exact(device_prune(ctx, M, hard, select, recover, pct), want)


def test_prune_long_columns(ctx):
    """columns of 1e5..3e5 entries (radix select over long columns) next to many short ones; MCL's default parameters"""
    rng = np.random.default_rng(11)
    m, n = 400000, 300
    lens = np.minimum(m, (rng.pareto(0.9, n) * 40).astype(np.int64))
    lens[:6] = [300000, 100000, 1100, 1401, 20000, 12000]
    rows = np.concatenate([np.sort(rng.choice(m, int(k), replace=False)) for k in lens])
    cols = np.repeat(np.arange(n), lens)
    vals = rng.random(len(rows)) ** 4
    # ties that do not fit the shared-memory candidate buffer: a column of equal values, and one whose selectNum-th
    # largest entry lies inside a run of 9000 equal values
    o4, o5 = int(lens[:4].sum()), int(lens[:5].sum())
    vals[o4:o4 + 20000] = 0.125
    vals[o5:o5 + 9000] = 0.25
    vals[o5 + 9000:o5 + 12000] = 0.25 + rng.random(3000)
    vals[o5:o5 + 12000] = rng.permutation(vals[o5:o5 + 12000])
    M = sp.csc_matrix((vals, (rows, cols)), shape=(m, n))
    s = np.asarray(M.sum(0)).ravel()
    s[s == 0] = 1
    A = Csc.from_scipy(sp.csc_matrix(M @ sp.diags(1 / s)), np.float64)
    want, _ = mcl_prune_recovery_select(A, 1e-4, 1100, 1400, 0.9)
    got = device_prune(ctx, A, 1e-4, 1100, 1400, 0.9)
    assert_bit_exact(got, want)
    assert got[3].cols_selected >= 4
    # a select threshold inside the tie run (column 5: 3000 larger values, then 9000 equal ones)
    want2, _ = mcl_prune_recovery_select(A, 1e-9, 5000, 0, 0.9)
    assert_bit_exact(device_prune(ctx, A, 1e-9, 5000, 0, 0.9), want2)
    # recover rule on long columns: everything is below the hard threshold, the top recoverNum entries come back
    want3, _ = mcl_prune_recovery_select(A, 0.9, 0, 7000, 10.0)
    assert_bit_exact(device_prune(ctx, A, 0.9, 0, 7000, 10.0), want3)


def test_make_col_stochastic_and_inflate(ctx):
    A = skewed_stochastic(300, 300, 7, np.float64)
    A = Csc(A.m, A.n, A.colptr, A.rows, A.vals * 3.7)
    dA = ctx.upload(to_host_block(A))
    ctx.make_col_stochastic(dA)
    _, _, vals = ctx.download_coo(dA)
    M = A.to_scipy()
    s = np.asarray(M.sum(0)).ravel()
    s[s == 0] = 1
    want = sp.csc_matrix(M @ sp.diags(1 / s))
    want.sort_indices()
    assert np.allclose(vals, want.data, rtol=1e-12, atol=0)
    ctx.inflate(dA, 2.0)
    _, _, vals2 = ctx.download_coo(dA)
    P = want.copy()
    P.data = P.data ** 2.0
    s = np.asarray(P.sum(0)).ravel()
    s[s == 0] = 1
    want2 = sp.csc_matrix(P @ sp.diags(1 / s))
    want2.sort_indices()
    assert np.allclose(vals2, want2.data, rtol=1e-11, atol=0)
    # the block is still a valid multiply operand afterwards (cached operand copies were dropped)
    C = ctx.spgemm(0, dA, dA)
    assert C.info().nnz > 0
    C.free()
    dA.free()


def dyadic_rmat(scale, seed):
    """R-MAT pattern with weights k/256: every product and column sum of A^2 is exact in double precision, so the device
    and the reference produce bit-identical expansion slabs and must take identical pruning decisions"""
    M = rmat(scale, 8, seed)
    rng = np.random.default_rng(seed)
    M.data = rng.integers(1, 256, len(M.data)).astype(np.float64) / 256.0
    return M


@pytest.mark.parametrize("phases", [1, 4])
def test_memefficient_spgemm_with_pruning_equals_reference(ctx, ref_oracle, phases):
    """BASELINE config 4 (HipMCL expansion: A^2 with column pruning / top-k) at reduced size, single rank"""
    M = dyadic_rmat(10, 3)
    A = Csc.from_scipy(M, np.float64)
    for hard, select, recover, pct in [(0.4, 20, 25, 40.0), (0.05, 12, 0, 0.9), (2.0, 30, 40, 1e6)]:
        want = ref_oracle.memeff_prune(A, A, phases, hard, select, recover, pct)
        got = cb.MemEfficientSpGEMM(ctx, cb.PlusTimesSRing_f64, cb.SpDCCols.from_scipy(M, np.float64), cb.SpDCCols.from_scipy(M, np.float64),
                                    phases, hard, select, recover, pct)
        assert got.getnnz() == want.nnz
        assert np.array_equal(got.cols, want.cols_expanded()) and np.array_equal(got.rows, want.rows)
        assert np.array_equal(got.vals, want.vals)
        assert 0 < want.nnz


def test_memefficient_spgemm_device_entry(ctx, ref_oracle):
    """cbgpu_memefficient_spgemm on resident operands: automatic phase count, statistics, integer results unpruned"""
    M = dyadic_rmat(10, 3)
    A = Csc.from_scipy(M, np.float64)
    dA = ctx.upload(cb.SpDCCols.from_scipy(M, np.float64))
    want = ref_oracle.memeff_prune(A, A, 1, 0.4, 20, 25, 40.0)
    for phases in (0, 5):
        D, st = ctx.memefficient_spgemm(cb.PlusTimesSRing_f64, dA, dA, phases, 0.4, 20, 25, 40.0, want_stats=True)
        rows, cols, vals = ctx.download_coo(D)
        assert np.array_equal(cols, want.cols_expanded()) and np.array_equal(rows, want.rows) and np.array_equal(vals, want.vals)
        assert st.phases == (phases if phases else 1) and st.nnz_out == want.nnz and st.nnz_unpruned > st.nnz_out
        assert st.cols_selected > 0 and st.flops > 0
        D.free()
    dA.free()
    Mi = M.copy()
    Mi.data = np.floor(Mi.data * 256)
    dI = ctx.upload(cb.SpDCCols.from_scipy(Mi, np.int64))
    D, st = ctx.memefficient_spgemm(cb.PlusTimesSRing_i64, dI, dI, 3, 0.4, 20, 25, 40.0, want_stats=True)
    assert st.nnz_out == st.nnz_unpruned == (Mi @ Mi).nnz
    D.free()
    dI.free()


def test_hipmcl_expansion_step(ctx, ref_oracle):
    """column-stochastic weighted R-MAT (MCL.cpp:389-394), one expansion with MCL-style parameters scaled to the size"""
    M = rmat(11, 8, 5)
    rng = np.random.default_rng(5)
    M.data = rng.random(len(M.data)) + 1e-3
    M = sp.csc_matrix(M + sp.identity(M.shape[0], format="csc"))  # AddLoops
    s = np.asarray(M.sum(0)).ravel()
    M = sp.csc_matrix(M @ sp.diags(1 / s))
    M.sort_indices()
    A = Csc.from_scipy(M, np.float64)
    want = ref_oracle.memeff_prune(A, A, 3, 1e-3, 60, 80, 0.9)
    H = cb.SpDCCols.from_scipy(M, np.float64)
    got = cb.MemEfficientSpGEMM(ctx, cb.PlusTimesSRing_f64, H, H, 3, 1e-3, 60, 80, 0.9)
    assert got.getnnz() == want.nnz
    assert np.array_equal(got.cols, want.cols_expanded()) and np.array_equal(got.rows, want.rows)
    assert np.allclose(got.vals, want.vals, rtol=1e-12, atol=0)


@pytest.mark.skipif(os.environ.get("CBGPU_TEST_EXPERIMENTAL", "0") != "1",
                    reason="written without GPU time left to run it once; enable with CBGPU_TEST_EXPERIMENTAL=1")
def test_prune_fuzz_against_oracle(ctx):
    """the seeded draws of tests/test_prune_oracle.py::test_restatement_fuzz_against_reference on the device"""
    rng = np.random.default_rng(2026)
    for it in range(80):
        m, n = int(rng.integers(1, 400)), int(rng.integers(3, 120))
        A = skewed_stochastic(max(m, 8), n, 1000 + it, np.float64)
        if it % 4 == 0:
            A = Csc(A.m, A.n, A.colptr, A.rows, np.round(A.vals * 8) / 8 + (rng.integers(0, 2, A.nnz) * 0.125))
        hard = float(rng.choice([0.0, -1.0, 1e-4, 1e-2, 0.05, 0.3]))
        select = int(rng.choice([0, 1, 2, 5, 17, 1100]))
        recover = int(rng.choice([0, 1, 3, 9, 25, 1400]))
        pct = float(rng.choice([0.0, 0.3, 0.9, 1.5, 1e9]))
        want, _ = mcl_prune_recovery_select(A, hard, select, recover, pct)
        assert_bit_exact(device_prune(ctx, A, hard, select, recover, pct), want)
