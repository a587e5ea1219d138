"""User-defined semirings (include/combblas_b200/device_semiring.cuh): structs with the reference's static-member semiring
interface (Semirings.h:143-255; KTipsTest.cpp:12-20) instantiated into the accumulation engine by a translation unit of the
application (tests/user_semiring/my_semirings.cu -> libmy_semirings.so) and registered with libcbgpu.so at run time."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

import combblas_b200 as cb
from oracle.oracle import Csc, esc_spgemm
from tests.util import rmat, typed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
USER_LIB = os.path.join(ROOT, "tests", "user_semiring", "libmy_semirings.so")

# (exported symbol, value dtypes (A, B, C), multiply on numpy arrays, add ufunc)
I32_INF = np.iinfo(np.int32).max
USER = {
    "ktips_or_and_id": ((np.uint8, np.uint8, np.uint8), lambda x, y: x & y, np.logical_or),
    "max_times_f64_id": ((np.float64, np.float64, np.float64), lambda x, y: x * y, np.maximum),
    "min_plus_i32_id": ((np.int32, np.int32, np.int32),
                        lambda x, y: np.where((x == I32_INF) | (y == I32_INF), I32_INF, x.astype(np.int64) + y).astype(np.int32), np.minimum),
}


def user_lib():
    if not os.path.exists(USER_LIB):
        pytest.skip("tests/user_semiring/libmy_semirings.so not built (__graft_entry__.build() builds it)")
    return USER_LIB


def test_registration_and_type_registry():
    """no GPU needed: the application's unit registers once per semiring, ids start at CBGPU_SR_USER_BASE and the library
    reports the declared value types; an unknown id stays an error"""
    lib = cb.load_library()
    ids = [cb.load_user_semiring(user_lib(), sym) for sym in USER]
    assert ids == sorted(ids) and len(set(ids)) == 3 and min(ids) >= 64
    assert [cb.load_user_semiring(user_lib(), sym) for sym in USER] == ids  # idempotent
    for sr, sym in zip(ids, USER):
        assert tuple(np.dtype(t) for t in cb.semiring_types(sr)) == tuple(np.dtype(t) for t in USER[sym][0])
    a = C.c_int()
    assert lib.cbgpu_semiring_types(max(ids) + 1, C.byref(a), C.byref(a), C.byref(a)) == -1
    assert lib.cbgpu_semiring_types(63, C.byref(a), C.byref(a), C.byref(a)) == -1


def test_esc_oracle_is_pinned_to_the_reference(oracle):
    """the expand-sort-fold oracle used for user semirings equals the compiled reference (or its C restatement) on the library
    semirings that have the same arithmetic: OR-AND (5), MinPlus f64 (4), SelectMax (8), PlusTimes i64 (2)"""
    rng = np.random.default_rng(3)
    A = sp.random(300, 200, density=0.05, random_state=rng, format="csc")
    B = sp.random(200, 260, density=0.04, random_state=rng, format="csc")
    cases = [(5, np.uint8, lambda x, y: x & y, np.logical_or), (4, np.float64, lambda x, y: x + y, np.minimum),
             (8, np.int64, lambda x, y: x * y, np.maximum), (2, np.int64, lambda x, y: x * y, np.add)]
    for sr, dt, mul, add in cases:
        a, b = Csc.from_scipy(typed(A, dt), dt), Csc.from_scipy(typed(B, dt), dt)
        got, want = esc_spgemm(a, b, mul, add, dt), oracle.spgemm(a, b, sr)
        assert np.array_equal(got.colptr, want.colptr) and np.array_equal(got.rows, want.rows)
        assert np.array_equal(got.vals, want.vals)


def values_for(M, dt, seed):
    M = M.tocsc().copy()
    M.sort_indices()
    rng = np.random.default_rng(seed)
    if np.dtype(dt) == np.uint8:
        M.data = np.ones(M.nnz)
    elif np.dtype(dt) == np.int32:
        M.data = rng.integers(1, 1000, M.nnz).astype(np.float64)
    else:
        M.data = rng.integers(1, 64, M.nnz) / 8.0
    return M


def check(ctx, sym, A, B):
    (ta, tb, tc), mul, add = USER[sym]
    sr = cb.load_user_semiring(user_lib(), sym)
    A, B = values_for(A, ta, 1), values_for(B, tb, 2)
    got = cb.LocalHybridSpGEMM(ctx, sr, cb.SpDCCols.from_scipy(A, ta), cb.SpDCCols.from_scipy(B, tb))
    want = esc_spgemm(Csc.from_scipy(A, ta), Csc.from_scipy(B, tb), mul, add, tc)
    assert got.getnnz() == want.nnz
    assert np.array_equal(got.cols, want.cols_expanded()) and np.array_equal(got.rows, want.rows)
    assert got.vals.dtype == np.dtype(tc) and np.array_equal(got.vals, want.vals)
    return sr, got


@pytest.mark.gpu
@pytest.mark.parametrize("sym", list(USER))
def test_user_semiring_multiply(ctx, sym):
    rng = np.random.default_rng(5)
    for (m, k, n, da, db) in [(300, 220, 260, 0.05, 0.04), (64, 2000, 50, 0.02, 0.3), (3000, 40, 3000, 0.2, 0.01)]:
        A = sp.random(m, k, density=da, random_state=rng, format="csc")
        B = sp.random(k, n, density=db, random_state=rng, format="csc")
        check(ctx, sym, A, B)


@pytest.mark.gpu
@pytest.mark.parametrize("sym", list(USER))
def test_user_semiring_every_kernel_class(ctx, sym):
    """R-MAT squared with small row windows and small shared-accumulator capacities: hash per warp / per CTA, bitmap with
    shared-memory accumulators (exchange protocol on the user's add) and bitmap with accumulators in C (compare-and-swap on
    the user's add; for bool a byte of C inside its aligned word)"""
    G = rmat(12, 8, seed=3)
    for opts in [{}, {"bitmap_window_log2": 10, "shared_acc_max": 256}, {"bitmap_window_log2": 11, "shared_acc": 0}, {"force_path": 1}]:
        old = {k: ctx.get_option(k) for k in opts}
        for k, v in opts.items():
            ctx.set_option(k, v)
        try:
            check(ctx, sym, G, G)
        finally:
            for k, v in old.items():
                ctx.set_option(k, v)


@pytest.mark.gpu
def test_user_or_and_equals_library_or_and_and_merges(ctx, oracle):
    """the KTips struct gives bit for bit what the library's OR-AND id gives; MultiwayMerge with the user's add"""
    G = rmat(11, 8, seed=9)
    sr, got = check(ctx, "ktips_or_and_id", G, G)
    M = values_for(G, np.uint8, 0)
    lib = cb.LocalHybridSpGEMM(ctx, cb.OrAndSRing_bool, cb.SpDCCols.from_scipy(M, np.uint8), cb.SpDCCols.from_scipy(M, np.uint8))
    assert np.array_equal(got.rows, lib.rows) and np.array_equal(got.cols, lib.cols) and np.array_equal(got.vals, lib.vals)
    # merge: max of three partial results == the oracle's fold
    srm = cb.load_user_semiring(user_lib(), "max_times_f64_id")
    rng = np.random.default_rng(11)
    parts = [values_for(sp.random(500, 400, density=0.03, random_state=rng, format="csc"), np.float64, 20 + i) for i in range(3)]
    mg = cb.MultiwayMerge(ctx, srm, [cb.SpDCCols.from_scipy(P, np.float64) for P in parts])
    dense = np.zeros((500, 400))
    for P in parts:
        dense = np.maximum(dense, P.toarray())
    W = sp.csc_matrix(dense)
    W.sort_indices()
    assert mg.getnnz() == W.nnz and np.array_equal(mg.rows, W.indices) and np.array_equal(mg.vals, W.data)
