"""ctypes front-ends for the two CPU parity oracles.

TEST INFRASTRUCTURE ONLY. May be imported by tests/, by ``__graft_entry__.smoke()`` and by
``bench.py``'s cpu_baseline / ``--impl reference`` legs -- never by the product package
``combblas_b200`` (a test asserts this).

* ``RefOracle``  -- oracle/_ref/libref_oracle.so: the UNMODIFIED reference (CombBLAS v2.0.1) compiled
  from /root/reference in the build container against a single-rank mpi.h stand-in (oracle/Makefile).
* ``PortOracle`` -- oracle/libspgemm_oracle.so: plain-C restatement of the same algorithm
  (oracle/spgemm_oracle.c, each function cites the reference file:line it follows).

Matrices are (m, n, colptr[int64 n+1], rows[int64 nnz], vals[typed nnz]) CSC triples ("Csc").
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# semiring ids -- identical numbering to include/cbgpu.h (cbgpu_semiring)
SR_PLUS_TIMES_F64 = 0
SR_PLUS_TIMES_F32 = 1
SR_PLUS_TIMES_I64 = 2
SR_SELECT_MAX_BOOL_I64 = 3
SR_MIN_PLUS_F64 = 4
SR_OR_AND_BOOL = 5
SR_PLUS_TIMES_BOOL_F64 = 6
SR_PLUS_TIMES_I32 = 7
SR_SELECT_MAX_I64 = 8

# (A dtype, B dtype, C dtype) per semiring; bool travels as uint8
SR_DTYPES = {
    0: (np.float64, np.float64, np.float64),
    1: (np.float32, np.float32, np.float32),
    2: (np.int64, np.int64, np.int64),
    3: (np.uint8, np.int64, np.int64),
    4: (np.float64, np.float64, np.float64),
    5: (np.uint8, np.uint8, np.uint8),
    6: (np.uint8, np.float64, np.float64),
    7: (np.int32, np.int32, np.int32),
    8: (np.int64, np.int64, np.int64),
    9: (np.uint8, np.float64, np.float64),
    10: (np.float64, np.uint8, np.float64),
    11: (np.uint8, np.int64, np.int64),
    12: (np.int64, np.uint8, np.int64),
    13: (np.uint8, np.uint8, np.uint8),
    14: (np.uint8, np.uint8, np.uint8),
}
SR_NAMES = {
    0: "PlusTimes<f64>", 1: "PlusTimes<f32>", 2: "PlusTimes<i64>", 3: "SelectMax<bool,i64>", 4: "MinPlus<f64>",
    5: "OrAnd<bool>", 6: "PlusTimes<bool,f64>", 7: "PlusTimes<i32>", 8: "SelectMax<i64>",
    9: "BoolCopy2nd<f64>", 10: "BoolCopy1st<f64>", 11: "BoolCopy2nd<i64>", 12: "BoolCopy1st<i64>", 13: "BoolCopy2nd<bool>",
    14: "BoolCopy1st<bool>",
}

# reference routines (oracle/ref_oracle.h)
REF_LOCAL_HYBRID, REF_LOCAL_HASH_SORTED, REF_LOCAL_HASH_UNSORTED, REF_LOCAL_HEAP = 0, 1, 2, 3
REF_DIST_SYNCH, REF_DIST_DOUBLEBUFF, REF_DIST_MEMEFF_HASH, REF_DIST_MEMEFF_HEAP, REF_DIST_SUMMA3D = 10, 11, 12, 13, 14


@dataclass
class Csc:
    m: int
    n: int
    colptr: np.ndarray  # int64 [n+1]
    rows: np.ndarray  # int64 [nnz]
    vals: np.ndarray  # typed [nnz]

    @property
    def nnz(self) -> int:
        return int(self.colptr[-1])

    def astype(self, dt) -> "Csc":
        return Csc(self.m, self.n, self.colptr, self.rows, np.ascontiguousarray(self.vals.astype(dt)))

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csc_matrix((self.vals, self.rows, self.colptr), shape=(self.m, self.n))

    @staticmethod
    def from_scipy(M, dtype=None) -> "Csc":
        M = M.tocsc()
        M.sort_indices()
        v = M.data if dtype is None else M.data.astype(dtype)
        return Csc(M.shape[0], M.shape[1], M.indptr.astype(np.int64), M.indices.astype(np.int64), np.ascontiguousarray(v))

    @staticmethod
    def from_coo(m, n, rows, cols, vals) -> "Csc":
        """column-major, rows ascending; duplicates must already be combined."""
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        order = np.lexsort((rows, cols))
        rows, cols, vals = rows[order], cols[order], np.asarray(vals)[order]
        colptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(colptr, cols + 1, 1)
        np.cumsum(colptr, out=colptr)
        return Csc(m, n, colptr, np.ascontiguousarray(rows), np.ascontiguousarray(vals))

    def cols_expanded(self) -> np.ndarray:
        return np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.colptr))


class _RefCsc(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("nnz", C.c_int64), ("colptr", C.c_void_p), ("rows", C.c_void_p),
                ("vals", C.c_void_p)]


def _as_ref(M: Csc, dt):
    cp = np.ascontiguousarray(M.colptr, dtype=np.int64)
    ro = np.ascontiguousarray(M.rows, dtype=np.int64)
    va = np.ascontiguousarray(M.vals, dtype=dt)
    s = _RefCsc(M.m, M.n, int(cp[-1]), cp.ctypes.data, ro.ctypes.data, va.ctypes.data)
    return s, (cp, ro, va)


def _coo_to_csc(m, n, rows, cols, vals, keep_order):
    """COO as produced by an oracle -> Csc. With keep_order the within-column order is preserved."""
    colptr = np.zeros(n + 1, dtype=np.int64)
    if len(cols):
        if np.any(np.diff(cols) < 0):
            raise AssertionError("oracle output is not column-grouped ascending")
        np.add.at(colptr, cols + 1, 1)
    np.cumsum(colptr, out=colptr)
    return Csc(m, n, colptr, rows, vals)


class _Base:
    lib = None

    def _result(self, h, m, n, out_dt):
        nnz = self.lib.ref_result_nnz(h) if self._prefix == "ref" else self.lib.port_result_nnz(h)
        rows = np.empty(nnz, dtype=np.int64)
        cols = np.empty(nnz, dtype=np.int64)
        vals = np.empty(nnz, dtype=out_dt)
        copy = getattr(self.lib, self._prefix + "_result_copy")
        free = getattr(self.lib, self._prefix + "_result_free")
        copy(h, rows.ctypes.data, cols.ctypes.data, vals.ctypes.data)
        free(h)
        return _coo_to_csc(m, n, rows, cols, vals, True)


class RefOracle(_Base):
    """The unmodified reference, prebuilt into oracle/_ref/libref_oracle.so."""

    _prefix = "ref"
    PATH = os.path.join(HERE, "_ref", "libref_oracle.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = C.CDLL(self.PATH)
        L = self.lib
        L.ref_spgemm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_double)]
        L.ref_merge.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                C.POINTER(C.c_double)]
        L.ref_symbolic.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_result_nnz.argtypes = [C.c_void_p]
        L.ref_result_nnz.restype = C.c_int64
        L.ref_result_copy.argtypes = [C.c_void_p] * 4
        L.ref_result_free.argtypes = [C.c_void_p]
        L.ref_set_num_threads.argtypes = [C.c_int]
        if hasattr(L, "ref_mcl_prune"):
            L.ref_mcl_prune.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_int64, C.c_int64, C.c_double, C.c_int,
                                        C.POINTER(C.c_void_p)]
            L.ref_memeff_prune.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_int64,
                                           C.c_double, C.c_int, C.c_int, C.POINTER(C.c_void_p)]

    def num_threads(self) -> int:
        return int(self.lib.ref_num_threads())

    def set_num_threads(self, n: int):
        self.lib.ref_set_num_threads(int(n))

    def mcl_prune(self, A: Csc, hard: float, select: int, recover: int, pct: float, sr: int = 0, kselect_version: int = 1):
        """The reference's MCLPruneRecoverySelect (ParFriends.h:186-354) on A as a P=1 SpParMat."""
        dt = SR_DTYPES[sr][2]
        sa, ka = _as_ref(A, dt)
        h = C.c_void_p()
        rc = self.lib.ref_mcl_prune(sr, C.byref(sa), hard, select, recover, pct, kselect_version, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ref_mcl_prune failed rc={rc}")
        return self._result(h, A.m, A.n, dt)

    def memeff_prune(self, A: Csc, B: Csc, phases: int, hard: float, select: int, recover: int, pct: float, sr: int = 0,
                     kselect_version: int = 1, kernel: int = 1):
        """The reference's MemEfficientSpGEMM (ParFriends.h:453-777) at P=1 with its pruning parameters."""
        da, db, dc = SR_DTYPES[sr]
        sa, ka = _as_ref(A, da)
        sb, kb = _as_ref(B, db)
        h = C.c_void_p()
        rc = self.lib.ref_memeff_prune(sr, C.byref(sa), C.byref(sb), phases, hard, select, recover, pct, kselect_version,
                                       kernel, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ref_memeff_prune failed rc={rc}")
        return self._result(h, A.m, B.n, dc)

    def spgemm(self, A: Csc, B: Csc, sr: int = 0, routine: int = REF_LOCAL_HYBRID, canonical: bool = True,
               phases: int = 1, want_time: bool = False):
        da, db, dc = SR_DTYPES[sr]
        sa, ka = _as_ref(A, da)
        sb, kb = _as_ref(B, db)
        h = C.c_void_p()
        sec = C.c_double(0)
        rc = self.lib.ref_spgemm(routine, sr, C.byref(sa), C.byref(sb), phases, int(canonical), C.byref(h), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"ref_spgemm failed rc={rc}")
        out = self._result(h, A.m, B.n, dc)
        del ka, kb
        return (out, sec.value) if want_time else out

    def merge(self, lists, sr: int = 0, hash: bool = False, sorted: bool = True, canonical: bool = True):
        dc = SR_DTYPES[sr][2]
        arr = (_RefCsc * len(lists))()
        keep = []
        for i, M in enumerate(lists):
            s, k = _as_ref(M, dc)
            arr[i] = s
            keep.append(k)
        h = C.c_void_p()
        sec = C.c_double(0)
        rc = self.lib.ref_merge(int(hash), sr, len(lists), arr, int(sorted), int(canonical), C.byref(h), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"ref_merge failed rc={rc}")
        return self._result(h, lists[0].m, lists[0].n, dc)

    def symbolic(self, A: Csc, B: Csc, sr: int = 0):
        """(flop, nnz) per NON-EMPTY column of B, as estimateFLOP / estimateNNZ_Hash return them."""
        da, db, _ = SR_DTYPES[sr]
        sa, ka = _as_ref(A, da)
        sb, kb = _as_ref(B, db)
        nzc = C.c_int64(0)
        pf, pn = C.c_void_p(), C.c_void_p()
        rc = self.lib.ref_symbolic(sr, C.byref(sa), C.byref(sb), C.byref(nzc), C.byref(pf), C.byref(pn))
        if rc != 0:
            raise RuntimeError("ref_symbolic failed")
        k = nzc.value
        if k == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        flop = np.ctypeslib.as_array(C.cast(pf, C.POINTER(C.c_int64)), shape=(k,)).copy()
        nnz = np.ctypeslib.as_array(C.cast(pn, C.POINTER(C.c_int64)), shape=(k,)).copy()
        self.lib.ref_free(pf)
        self.lib.ref_free(pn)
        return flop, nnz


class PortOracle(_Base):
    """Plain-C restatement (oracle/spgemm_oracle.c); always buildable, travels as source."""

    _prefix = "port"
    PATH = os.path.join(HERE, "libspgemm_oracle.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.PATH)

    @classmethod
    def build(cls):
        import subprocess

        subprocess.check_call(["make", "-s", "-C", HERE, "port"])

    def __init__(self):
        if not self.available():
            self.build()
        self.lib = C.CDLL(self.PATH)
        L = self.lib
        L.port_spgemm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_double)]
        L.port_merge.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_double)]
        L.port_symbolic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.port_symbolic.restype = C.c_int64
        L.port_result_nnz.argtypes = [C.c_void_p]
        L.port_result_nnz.restype = C.c_int64
        L.port_result_copy.argtypes = [C.c_void_p] * 4
        L.port_result_free.argtypes = [C.c_void_p]
        L.port_set_num_threads.argtypes = [C.c_int]

    def num_threads(self) -> int:
        return int(self.lib.port_num_threads())

    def set_num_threads(self, n: int):
        self.lib.port_set_num_threads(int(n))

    def spgemm(self, A: Csc, B: Csc, sr: int = 0, sort: bool = True, want_time: bool = False):
        da, db, dc = SR_DTYPES[sr]
        sa, ka = _as_ref(A, da)
        sb, kb = _as_ref(B, db)
        h = C.c_void_p()
        sec = C.c_double(0)
        rc = self.lib.port_spgemm(sr, C.byref(sa), C.byref(sb), int(sort), C.byref(h), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"port_spgemm failed rc={rc}")
        out = self._result(h, A.m, B.n, dc)
        del ka, kb
        return (out, sec.value) if want_time else out

    def merge(self, lists, sr: int = 0, sort: bool = True):
        dc = SR_DTYPES[sr][2]
        arr = (_RefCsc * len(lists))()
        keep = []
        for i, M in enumerate(lists):
            s, k = _as_ref(M, dc)
            arr[i] = s
            keep.append(k)
        h = C.c_void_p()
        sec = C.c_double(0)
        rc = self.lib.port_merge(sr, len(lists), arr, int(sort), C.byref(h), C.byref(sec))
        if rc != 0:
            raise RuntimeError(f"port_merge failed rc={rc}")
        return self._result(h, lists[0].m, lists[0].n, dc)

    def symbolic(self, A: Csc, B: Csc):
        """(flop, nnz) per column of B (all n columns, zeros for empty ones)."""
        sa, ka = _as_ref(A, A.vals.dtype)
        sb, kb = _as_ref(B, B.vals.dtype)
        flop = np.zeros(B.n, dtype=np.int64)
        nnz = np.zeros(B.n, dtype=np.int64)
        self.lib.port_symbolic(C.byref(sa), C.byref(sb), flop.ctypes.data, nnz.ctypes.data)
        return flop, nnz


def kselect1(col_vals: np.ndarray, k: int):
    """Kselect1 for one column (SpParMat.cpp:1413-1700, result rule :1672-1684): the k-th largest entry; a column with
    fewer than k entries yields its smallest entry, an empty one numeric_limits<NT>::min()."""
    n = len(col_vals)
    if n == 0:
        return np.finfo(col_vals.dtype).tiny
    s = np.sort(col_vals)[::-1]
    return s[k - 1] if n >= k >= 1 else s[-1]


def mcl_prune_recovery_select(A: Csc, hard, select: int, recover: int, pct):
    """numpy restatement of MCLPruneRecoverySelect (ParFriends.h:186-354) for a matrix whose columns are whole
    (P = 1): per column, statistics of the entries above the hard threshold (:196-201), the recover rule (:208-243),
    the select rule with the second recovery check (:248-335), then PruneColumn(thresholds, std::less) (:339).
    Returns (pruned Csc, per-column thresholds)."""
    dt = A.vals.dtype
    hard = dt.type(hard)
    pct = dt.type(pct)
    thr = np.full(A.n, hard, dtype=dt)
    keep = np.zeros(A.nnz, dtype=bool)
    for j in range(A.n):
        b, e = int(A.colptr[j]), int(A.colptr[j + 1])
        v = A.vals[b:e]
        pruned = v[v > hard]  # A.Prune(val <= hardThreshold)
        n_all, n_pr = len(v), len(pruned)
        s_pr = pruned.sum(dtype=dt) if n_pr else dt.type(0)
        t = hard
        if n_pr < recover and n_all > n_pr and s_pr < pct:
            t = kselect1(v, recover)
        elif select > 0 and n_pr > select:
            t = kselect1(v, select)
            if recover > 0:
                sel = v[~(v < t)]
                if len(sel) < recover and sel.sum(dtype=dt) < pct:
                    t = kselect1(v, recover)
        thr[j] = t
        keep[b:e] = ~(v < t)
    colptr = np.zeros(A.n + 1, dtype=np.int64)
    cols = A.cols_expanded()
    np.add.at(colptr, cols[keep] + 1, 1)
    np.cumsum(colptr, out=colptr)
    return Csc(A.m, A.n, colptr, A.rows[keep].copy(), A.vals[keep].copy()), thr



def esc_spgemm(A: Csc, B: Csc, multiply, add_ufunc, out_dtype) -> Csc:
    """Oracle for semirings outside the compiled list (a driver's own struct): expand every product
    multiply(A(i,k), B(k,j)) (the products the column loop of mtSpGEMM.h:362-440 forms), sort them by (column, row) and fold
    equal keys with the semiring's add. Only for associative, commutative adds whose numpy ufunc is exact (max, min,
    logical or, integer +): the fold order then does not matter, as it does not between the reference's own heap and hash
    kernels. `multiply` takes two numpy arrays (A values, B values of the same length)."""
    bcols = np.repeat(np.arange(B.n, dtype=np.int64), np.diff(B.colptr))  # column of every B entry
    k = B.rows.astype(np.int64)                                           # its row = the column of A it scales
    lens = (A.colptr[k + 1] - A.colptr[k]).astype(np.int64)
    total = int(lens.sum())
    if total == 0:
        return Csc(A.m, B.n, np.zeros(B.n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0, out_dtype))
    owner = np.repeat(np.arange(len(k), dtype=np.int64), lens)             # B entry of every product
    first = np.cumsum(lens) - lens
    apos = A.colptr[k][owner] + (np.arange(total, dtype=np.int64) - first[owner])
    rows = A.rows[apos].astype(np.int64)
    cols = bcols[owner]
    vals = multiply(A.vals[apos], B.vals[owner]).astype(out_dtype)
    order = np.lexsort((rows, cols))
    rows, cols, vals = rows[order], cols[order], vals[order]
    head = np.ones(total, bool)
    head[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    starts = np.flatnonzero(head)
    out_vals = add_ufunc.reduceat(vals, starts).astype(out_dtype)
    out_rows, out_cols = rows[starts], cols[starts]
    colptr = np.zeros(B.n + 1, np.int64)
    np.add.at(colptr, out_cols + 1, 1)
    return Csc(A.m, B.n, np.cumsum(colptr), out_rows, out_vals)


def best_oracle():
    """The reference build when present (build container and, prebuilt, the GPU box), else the port."""
    return RefOracle() if RefOracle.available() else PortOracle()


def rmat_edges(scale, nedges, seed, a=0.57, b=0.19, c=0.19, scramble=True):
    """seeded R-MAT edge stream of the benchmark inputs on the host cores (oracle/spgemm_oracle.c port_rmat_edges: the same
    arithmetic as the library's device generator, asserted equal in tests/test_oracle.py)"""
    lib = C.CDLL(os.path.join(HERE, "libspgemm_oracle.so"))
    lib.port_rmat_edges.argtypes = [C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    rows = np.empty(nedges, np.int64)
    cols = np.empty(nedges, np.int64)
    if lib.port_rmat_edges(scale, nedges, seed, a, b, c, int(scramble), rows.ctypes.data, cols.ctypes.data) != 0:
        raise RuntimeError("port_rmat_edges failed")
    return rows, cols


def rmat_csc(scale, edgefactor, seed, a=0.57, b=0.19, c=0.19, scramble=True):
    """the benchmark's R-MAT as Csc: duplicates summed into the value (SpTuples.cpp:70-124 semantics), rows ascending"""
    rows, cols = rmat_edges(scale, edgefactor << scale, seed, a, b, c, scramble)
    n = 1 << scale
    key = cols * n + rows  # column-major order
    del rows, cols
    key.sort()
    uniq, counts = np.unique(key, return_counts=True)
    del key
    ucols = uniq // n
    urows = uniq - ucols * n
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(ucols, minlength=n), out=colptr[1:])
    return Csc(n, n, colptr, np.ascontiguousarray(urows), counts.astype(np.float64))


# ------------------------------------------------------------------------------------------------ checksums
# numpy restatement of cbgpu_mat_checksum_at (combblas_b200/csrc/api.cu): order-independent sums over the entries of a matrix
# placed at (row_offset, col_offset). Used by the checkers to compare column slabs of products that are too large to keep.
def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xff51afd7ed558ccd)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xc4ceb9fe1a85ec53)
    x ^= x >> np.uint64(33)
    return x


def matrix_checksum(rows, cols, vals, row_offset=0, col_offset=0):
    """(pattern_sum, value_sum) as the device computes them; vals is a float64 / int64 / ... array (bit pattern hashed)."""
    with np.errstate(over="ignore"):
        r = (np.asarray(rows, dtype=np.int64) + row_offset).astype(np.uint64) & np.uint64(0xFFFFFFFF)
        c = (np.asarray(cols, dtype=np.int64) + col_offset).astype(np.uint64)
        key = (c << np.uint64(32)) ^ r
        h = _mix64(key + np.uint64(0x9E3779B97F4A7C15))
        v = np.ascontiguousarray(vals)
        nb = v.dtype.itemsize
        if nb == 8:
            vb = v.view(np.uint64).copy()
            vb[vb == np.uint64(0x8000000000000000)] = 0
        elif nb == 4:
            vb = v.view(np.uint32).astype(np.uint64)
            vb[vb == np.uint64(0x80000000)] = 0
        else:
            vb = v.view(np.uint8).astype(np.uint64)
        vs = _mix64(h ^ _mix64(vb + np.uint64(0x632BE59BD9B4E019)))
        return int(h.sum(dtype=np.uint64)), int(vs.sum(dtype=np.uint64))
