"""Shared helpers for the parity tests: seeded inputs, canonical comparison."""
import numpy as np
import scipy.sparse as sp

import combblas_b200 as cb
from oracle.oracle import Csc, SR_DTYPES


def typed(M, dt):
    """scipy matrix -> values of the requested operand type (bool: all true; ints: small positive)."""
    M = M.tocsc().copy()
    M.sort_indices()
    if np.dtype(dt) == np.uint8:
        M.data = np.ones_like(M.data)
    elif np.issubdtype(np.dtype(dt), np.integer):
        M.data = np.floor(np.abs(M.data) * 97) + 1
    return M


def random_pair(m, k, n, da, db, seed, dtypes):
    rng = np.random.default_rng(seed)
    A = sp.random(m, k, density=da, random_state=rng, format="csc")
    B = sp.random(k, n, density=db, random_state=rng, format="csc")
    ta, tb, _ = dtypes
    return typed(A, ta), typed(B, tb)


def rmat(scale, edgefactor, seed, a=0.57, b=0.19, c=0.19, scramble=True):
    """The library's own seeded R-MAT (host arithmetic identical to the device generator); duplicates summed."""
    import ctypes as C

    lib = cb.load_library()
    ne = edgefactor << scale
    rows = np.empty(ne, np.int64)
    cols = np.empty(ne, np.int64)
    rc = lib.cbgpu_rmat_edges_host(scale, ne, seed, a, b, c, int(scramble), rows.ctypes.data, cols.ctypes.data)
    assert rc == 0
    n = 1 << scale
    M = sp.coo_matrix((np.ones(ne), (rows, cols)), shape=(n, n)).tocsc()
    M.sum_duplicates()
    M.sort_indices()
    return M


def to_csc(M, dt):
    return Csc.from_scipy(M, dt)


def to_dcsc(M, dt, idx=np.int64):
    return cb.SpDCCols.from_scipy(M, dt, idx)


def assert_same(got, want, sr, rtol=None):
    """got: cb.SpTuples (device result), want: oracle Csc (canonical). Pattern bit-exact; values bit-exact for
    integer/bool/select-max/min-plus, relative tolerance for floating-point PlusTimes (accumulation order differs)."""
    wcols = want.cols_expanded()
    assert got.getnnz() == want.nnz, f"nnz differs: {got.getnnz()} vs {want.nnz}"
    assert np.array_equal(got.cols, wcols), "column indices differ"
    assert np.array_equal(got.rows, want.rows), "row indices differ (or are not ascending per column)"
    dc = SR_DTYPES[sr][2]
    assert got.vals.dtype == np.dtype(dc)
    if sr in (0, 6):
        tol = 1e-12 if rtol is None else rtol
        err = np.abs(got.vals - want.vals)
        assert np.all(err <= tol * np.maximum(np.abs(got.vals), np.abs(want.vals))), f"max rel err {np.max(err / np.maximum(np.abs(want.vals), 1e-300))}"
    elif sr == 1:
        tol = 1e-5 if rtol is None else rtol
        err = np.abs(got.vals.astype(np.float64) - want.vals.astype(np.float64))
        assert np.all(err <= tol * np.maximum(np.abs(got.vals), np.abs(want.vals)))
    else:
        assert np.array_equal(got.vals, want.vals), "values differ (bit-exact semiring)"


def subsref_operands(m, n, nri, nci, seed, dt):
    """Operands of SpParMat::SubsRef_SR (SpParMat.cpp:2515-2566): A (m x n, values of type dt), the boolean row selector
    S (nri x m, S[i, ri[i]] = 1) and the boolean column selector T (n x nci, T[ci[j], j] = 1). ri / ci may repeat indices;
    every entry of S*A and of (S*A)*T still receives exactly one product, which is what BoolCopy2nd/1stSRing rely on."""
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density=0.03, random_state=rng, format="csc")
    A.data = np.floor(A.data * 1000) - 300.0  # negative values and zeros included: the product must copy them bit for bit
    if np.dtype(dt) == np.uint8:
        A.data = (A.data > 0).astype(np.float64)
    ri = rng.integers(0, m, nri)
    ci = rng.integers(0, n, nci)
    S = sp.coo_matrix((np.ones(nri), (np.arange(nri), ri)), shape=(nri, m)).tocsc()
    T = sp.coo_matrix((np.ones(nci), (ci, np.arange(nci))), shape=(n, nci)).tocsc()
    for M in (A, S, T):
        M.sort_indices()
    return A, S, T, ri, ci
