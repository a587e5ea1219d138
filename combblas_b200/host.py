"""Host-side mirror of the reference interface for the SpGEMM path (names, argument meaning and error
behaviour follow the reference; the arithmetic always runs in libcbgpu.so on the GPU).

Reference types mirrored here (paths under the CombBLAS tree):
  SpDCCols<IT,NT>   include/CombBLAS/SpDCCols.h:51, Dcsc arrays dcsc.h:125-132
  SpTuples<IT,NT>   include/CombBLAS/SpTuples.h:64 (column-sorted triples)
  semiring structs  include/CombBLAS/Semirings.h:143-255
  local multiply    include/CombBLAS/mtSpGEMM.h:213 (LocalHybridSpGEMM), :463 (LocalSpGEMMHash), :74 (LocalSpGEMM)
  merge             include/CombBLAS/MultiwayMerge.h:428 (MultiwayMerge), :553 (MultiwayMergeHash)
  distributions     SpParMat::Owner SpParMat.cpp:5081, SpParMat3D::Owner SpParMat3D.cpp:337
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import lib as _l

# semiring ids (include/cbgpu.h cbgpu_semiring)
PlusTimesSRing_f64 = 0
PlusTimesSRing_f32 = 1
PlusTimesSRing_i64 = 2
SelectMaxSRing_bool_i64 = 3
MinPlusSRing_f64 = 4
OrAndSRing_bool = 5
PlusTimesSRing_bool_f64 = 6
PlusTimesSRing_i32 = 7
SelectMaxSRing_i64 = 8
# the indexing pair of SpParMat::SubsRef_SR (Semirings.h:51-138): add must not happen
BoolCopy2ndSRing_f64 = 9
BoolCopy1stSRing_f64 = 10
BoolCopy2ndSRing_i64 = 11
BoolCopy1stSRing_i64 = 12
BoolCopy2ndSRing_bool = 13
BoolCopy1stSRing_bool = 14

SEMIRINGS = {
    0: ("PlusTimesSRing<double,double>", np.float64, np.float64, np.float64),
    1: ("PlusTimesSRing<float,float>", np.float32, np.float32, np.float32),
    2: ("PlusTimesSRing<int64_t,int64_t>", np.int64, np.int64, np.int64),
    3: ("SelectMaxSRing<bool,int64_t>", np.uint8, np.int64, np.int64),
    4: ("MinPlusSRing<double,double>", np.float64, np.float64, np.float64),
    5: ("OrAndSRing<bool>", np.uint8, np.uint8, np.uint8),
    6: ("PlusTimesSRing<bool,double>", np.uint8, np.float64, np.float64),
    7: ("PlusTimesSRing<int32_t,int32_t>", np.int32, np.int32, np.int32),
    8: ("SelectMaxSRing<int64_t,int64_t>", np.int64, np.int64, np.int64),
    9: ("BoolCopy2ndSRing<double>", np.uint8, np.float64, np.float64),
    10: ("BoolCopy1stSRing<double>", np.float64, np.uint8, np.float64),
    11: ("BoolCopy2ndSRing<int64_t>", np.uint8, np.int64, np.int64),
    12: ("BoolCopy1stSRing<int64_t>", np.int64, np.uint8, np.int64),
    13: ("BoolCopy2ndSRing<bool>", np.uint8, np.uint8, np.uint8),
    14: ("BoolCopy1stSRing<bool>", np.uint8, np.uint8, np.uint8),
}


def semiring_types(sr: int):
    _, a, b, c = SEMIRINGS[sr]
    return a, b, c


_DTYPE_NP = {_l.F64: np.float64, _l.F32: np.float32, _l.I64: np.int64, _l.I32: np.int32, _l.BOOL: np.uint8}


def load_user_semiring(path: str, symbol: str, name: str = None) -> int:
    """A semiring the application defines itself (a struct with the reference's static id/add/multiply interface,
    Semirings.h:143-255, KTipsTest.cpp:12-20): `path` is the shared library the application built from its .cu with
    CBGPU_DEFINE_SEMIRING(symbol, ...) (include/combblas_b200/device_semiring.cuh). Returns the semiring id to pass wherever
    a library id goes; operand/result types come from the library's registry (cbgpu_semiring_types)."""
    import ctypes as C

    lib = _l.load_library()
    user = C.CDLL(os.path.abspath(path), mode=C.RTLD_GLOBAL)
    fn = getattr(user, symbol)
    fn.restype = C.c_int
    fn.argtypes = []
    sr = fn()
    if sr < 0:
        raise _l.CbgpuError(sr, f"{symbol} did not register with libcbgpu.so (built against another version?)")
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    rc = lib.cbgpu_semiring_types(sr, C.byref(a), C.byref(b), C.byref(c))
    if rc != 0:
        raise _l.CbgpuError(rc, f"semiring {sr} unknown to libcbgpu.so")
    SEMIRINGS[sr] = (name or symbol, _DTYPE_NP[a.value], _DTYPE_NP[b.value], _DTYPE_NP[c.value])
    return sr


@dataclass
class SpTuples:
    """Column-sorted triples (SpTuples.h:64): what the reference's local kernels and merges return."""

    m: int
    n: int
    rows: np.ndarray
    cols: np.ndarray
    vals: np.ndarray

    def getnnz(self):
        return len(self.rows)


class SpDCCols:
    """Host DCSC block: cp[nzc+1], jc[nzc], ir[nnz], numx[nnz] (dcsc.h:125-132); essentials {nnz,m,n,nzc}."""

    def __init__(self, m, n, jc, cp, ir, numx):
        self.m, self.n = int(m), int(n)
        self.jc, self.cp, self.ir, self.numx = jc, cp, ir, numx

    @property
    def nnz(self):
        return len(self.ir)

    @property
    def nzc(self):
        return len(self.jc)

    def getnrow(self):
        return self.m

    def getncol(self):
        return self.n

    def getnnz(self):
        return self.nnz

    def isZero(self):
        return self.nnz == 0

    @staticmethod
    def from_csc(m, n, colptr, rows, vals, idx_dtype=np.int64) -> "SpDCCols":
        colptr = np.asarray(colptr, dtype=np.int64)
        cnt = np.diff(colptr)
        jc = np.nonzero(cnt)[0].astype(idx_dtype)
        cp = np.concatenate([colptr[jc.astype(np.int64)], colptr[-1:]]).astype(idx_dtype) if len(jc) else np.zeros(1, idx_dtype)
        return SpDCCols(m, n, jc, cp, np.ascontiguousarray(rows, dtype=idx_dtype), np.ascontiguousarray(vals))

    @staticmethod
    def from_coo(m, n, rows, cols, vals, idx_dtype=np.int64) -> "SpDCCols":
        """SpDCCols(const SpTuples&, false): column-major, rows ascending. Duplicates must already be combined."""
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        vals = np.asarray(vals)
        order = np.lexsort((rows, cols))
        rows, cols, vals = rows[order], cols[order], vals[order]
        colptr = np.zeros(n + 1, dtype=np.int64)
        if len(cols):
            np.add.at(colptr, cols + 1, 1)
        np.cumsum(colptr, out=colptr)
        return SpDCCols.from_csc(m, n, colptr, rows, vals, idx_dtype)

    @staticmethod
    def from_scipy(M, dtype=None, idx_dtype=np.int64) -> "SpDCCols":
        M = M.tocsc()
        M.sort_indices()
        v = M.data if dtype is None else M.data.astype(dtype)
        return SpDCCols.from_csc(M.shape[0], M.shape[1], M.indptr, M.indices, v, idx_dtype)

    def to_csc(self):
        """-> (colptr[n+1] int64, rows int64, vals)"""
        colptr = np.zeros(self.n + 1, dtype=np.int64)
        if self.nzc:
            colptr[np.asarray(self.jc, dtype=np.int64) + 1] = np.diff(np.asarray(self.cp, dtype=np.int64))
        np.cumsum(colptr, out=colptr)
        return colptr, np.asarray(self.ir, dtype=np.int64), self.numx

    def to_tuples(self) -> SpTuples:
        cp = np.asarray(self.cp, dtype=np.int64)
        cols = np.repeat(np.asarray(self.jc, dtype=np.int64), np.diff(cp)) if self.nzc else np.zeros(0, np.int64)
        return SpTuples(self.m, self.n, np.asarray(self.ir, dtype=np.int64), cols, self.numx)

    def astype(self, dt) -> "SpDCCols":
        return SpDCCols(self.m, self.n, self.jc, self.cp, self.ir, np.ascontiguousarray(self.numx.astype(dt)))

    def colslice(self, c0, c1) -> "SpDCCols":
        jc = np.asarray(self.jc, dtype=np.int64)
        lo, hi = np.searchsorted(jc, [c0, c1])
        cp = np.asarray(self.cp, dtype=np.int64)
        p0, p1 = (cp[lo], cp[hi]) if self.nzc else (0, 0)
        return SpDCCols(self.m, c1 - c0, (jc[lo:hi] - c0).astype(self.jc.dtype), (cp[lo:hi + 1] - p0).astype(self.cp.dtype),
                        self.ir[p0:p1], self.numx[p0:p1])

    def submatrix(self, r0, r1, c0, c1) -> "SpDCCols":
        """block [r0,r1) x [c0,c1) with local indices (what SpParMat's 2D distribution hands each rank)."""
        t = self.to_tuples()
        keep = (t.rows >= r0) & (t.rows < r1) & (t.cols >= c0) & (t.cols < c1)
        return SpDCCols.from_coo(r1 - r0, c1 - c0, t.rows[keep] - r0, t.cols[keep] - c0, t.vals[keep], self.ir.dtype if len(self.ir) else np.int64)


def _download_tuples(ctx, D) -> SpTuples:
    rows, cols, vals = ctx.download_coo(D)
    m, n = D.shape
    return SpTuples(m, n, rows, cols, vals)


def _check_types(sr, A: SpDCCols, B: SpDCCols):
    a, b, _ = semiring_types(sr)
    if A.numx.dtype != np.dtype(a) or B.numx.dtype != np.dtype(b):
        raise TypeError(f"semiring {SEMIRINGS[sr][0]} needs operands of types ({np.dtype(a)}, {np.dtype(b)}), got "
                        f"({A.numx.dtype}, {B.numx.dtype})")


def LocalHybridSpGEMM(ctx, SR: int, A: SpDCCols, B: SpDCCols, clearA=False, clearB=False) -> SpTuples:
    """mtSpGEMM.h:213-217. Host blocks in, column-sorted tuples (rows ascending per column) out.
    clearA/clearB exist for signature parity (Python objects are garbage collected)."""
    _check_types(SR, A, B)
    D = ctx.spgemm_host(SR, A, B)
    out = _download_tuples(ctx, D)
    D.free()
    return out


def LocalSpGEMMHash(ctx, SR: int, A: SpDCCols, B: SpDCCols, clearA=False, clearB=False, sort=True) -> SpTuples:
    """mtSpGEMM.h:463-467. The device path always emits sorted columns; `sort=False` callers (which accept any
    within-column order) therefore also receive sorted output."""
    return LocalHybridSpGEMM(ctx, SR, A, B, clearA, clearB)


def LocalSpGEMM(ctx, SR: int, A: SpDCCols, B: SpDCCols, clearA=False, clearB=False) -> SpTuples:
    """mtSpGEMM.h:74-78 (heap kernel): same result contract as LocalHybridSpGEMM."""
    return LocalHybridSpGEMM(ctx, SR, A, B, clearA, clearB)


def CalculateNumberOfPhases(max_local_nnz_A: int, nnz_product_per_process: int, perProcessMemory: int, idx_bytes: int = 8,
                            in_val_bytes: int = 8, out_val_bytes: int = 8) -> int:
    """ParFriends.h:779-832 (host arithmetic only, no GPU): phases = 1 + asquareMem / remainingMem. The reference feeds
    it an estimate of the per-process nnz of the product; ctx.symbolic(A, B) gives the exact number."""
    from .lib import load_library

    r = load_library().cbgpu_calculate_phases(int(max_local_nnz_A), int(nnz_product_per_process), idx_bytes, in_val_bytes,
                                              out_val_bytes, int(perProcessMemory))
    if r < 1:
        raise ValueError("the operands alone do not fit perProcessMemory")
    return r


def MCLPruneRecoverySelect(ctx, A: SpDCCols, hardThreshold, selectNum: int, recoverNum: int, recoverPct,
                           kselectVersion: int = 1) -> SpTuples:
    """ParFriends.h:186-354 for a block holding whole columns (P = 1 / one rank per process column). Host block in,
    pruned column-sorted tuples out. Both Kselect versions of the reference give the same thresholds, so
    kselectVersion is accepted for signature parity only."""
    dA = ctx.upload(A)
    D = ctx.mcl_prune(dA, float(hardThreshold), int(selectNum), int(recoverNum), float(recoverPct))
    out = _download_tuples(ctx, D)
    D.free()
    dA.free()
    return out


def MemEfficientSpGEMM(ctx, SR: int, A: SpDCCols, B: SpDCCols, phases: int, hardThreshold, selectNum: int, recoverNum: int,
                       recoverPct, kselectVersion: int = 1, computationKernel: int = 1, perProcessMemory: int = 0) -> SpTuples:
    """ParFriends.h:452-777 at P = 1 (the HipMCL expansion step): B is cut into `phases` column slabs (ColSplit rule),
    every slab of C = A (x) B(:, slab) is pruned on the device by MCLPruneRecoverySelect before the next one is
    multiplied, and the pruned slabs are concatenated (ColConcatenate). Unpruned C never exists as a whole
    (cbgpu_memefficient_spgemm). phases <= 0 picks the phase count from the symbolic pass and the free HBM.
    computationKernel (hash / heap) and perProcessMemory are accepted for signature parity."""
    _check_types(SR, A, B)
    dA, dB = ctx.upload(A), ctx.upload(B)
    D = ctx.memefficient_spgemm(SR, dA, dB, int(phases), float(hardThreshold), int(selectNum), int(recoverNum), float(recoverPct))
    out = _download_tuples(ctx, D)
    D.free()
    dA.free()
    dB.free()
    return out


def MultiwayMerge(ctx, SR: int, ArrSpTups, mdim=0, ndim=0, delarrs=False) -> SpTuples:
    """MultiwayMerge.h:428-429: k column-sorted lists -> one, SR::add on equal (row, col).
    ArrSpTups: list of SpTuples (or SpDCCols). Zero lists -> empty mdim x ndim result."""
    if len(ArrSpTups) == 0:
        return SpTuples(mdim, ndim, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, semiring_types(SR)[2]))
    mats = []
    for t in ArrSpTups:
        d = t if isinstance(t, SpDCCols) else SpDCCols.from_coo(t.m, t.n, t.rows, t.cols, t.vals)
        mats.append(ctx.upload(d))
    D = ctx.merge(SR, mats)
    out = _download_tuples(ctx, D)
    D.free()
    for mth in mats:
        mth.free()
    return out


def MultiwayMergeHash(ctx, SR: int, ArrSpTups, mdim=0, ndim=0, delarrs=False, sorted=True) -> SpTuples:
    """MultiwayMerge.h:553-554: same contract; output columns are always sorted on the device path."""
    return MultiwayMerge(ctx, SR, ArrSpTups, mdim, ndim, delarrs)


def EstimateFLOP(ctx, A: SpDCCols, B: SpDCCols):
    """ParFriends.h:357 / estimateFLOP + estimateNNZ_Hash: (products, nnz(C)) of the local block pair."""
    dA, dB = ctx.upload(A), ctx.upload(B)
    r = ctx.symbolic(dA, dB)
    dA.free()
    dB.free()
    return r


# ---- distributions (pure host arithmetic, served by the C ABI so that C++ and Python callers agree)
def block_range(dim: int, parts: int, index: int):
    b, e = C.c_int64(), C.c_int64()
    rc = _l.load_library().cbgpu_block_range(dim, parts, index, C.byref(b), C.byref(e))
    if rc != 0:
        raise _l.CbgpuError(rc, "bad block range request")
    return b.value, e.value


def block_owner(dim: int, parts: int, gi: int) -> int:
    return int(_l.load_library().cbgpu_block_owner(dim, parts, gi))


def local_range(grid, m, n, split_cols: bool):
    r0, r1, c0, c1 = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    rc = _l.load_library().cbgpu_grid_local_range(C.byref(grid), m, n, int(split_cols), C.byref(r0), C.byref(r1), C.byref(c0), C.byref(c1))
    if rc != 0:
        raise _l.CbgpuError(rc, "bad grid")
    return r0.value, r1.value, c0.value, c1.value


def partition_2d(M: SpDCCols, grid) -> SpDCCols:
    """The block of the global matrix M owned by `grid.rank` under SpParMat::Owner (SpParMat.cpp:5081-5107)."""
    r0, r1, c0, c1 = local_range(grid, M.m, M.n, True)
    return M.submatrix(r0, r1, c0, c1)


def partition_3d(M: SpDCCols, grid, split_cols: bool) -> SpDCCols:
    """3D block: the 2D block of the layer grid cut by columns (A, C) or rows (B) into `layers` chunks
    (SpParMat3D::Owner / LocalDim, SpParMat3D.cpp:337-436)."""
    r0, r1, c0, c1 = local_range(grid, M.m, M.n, split_cols)
    return M.submatrix(r0, r1, c0, c1)
