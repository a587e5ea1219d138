"""ctypes binding of libcbgpu.so (include/cbgpu.h). Plumbing only; every compute call goes to the CUDA library."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

F64, F32, I64, I32, BOOL = 0, 1, 2, 3, 4
DTYPE_TO_NUMPY = {F64: np.float64, F32: np.float32, I64: np.int64, I32: np.int32, BOOL: np.uint8}
NUMPY_TO_DTYPE = {np.dtype(np.float64): F64, np.dtype(np.float32): F32, np.dtype(np.int64): I64,
                  np.dtype(np.int32): I32, np.dtype(np.uint8): BOOL, np.dtype(np.bool_): BOOL}

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    # CBGPU_LIB: a tuning build of the same library (tools/phase_timing.py); the product is always libcbgpu.so next to this file
    return os.environ.get("CBGPU_LIB") or os.path.join(_HERE, "libcbgpu.so")


class CbgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cbgpu error {code}: {msg}")
        self.code = code


class _DcscView(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("nnz", C.c_int64), ("nzc", C.c_int64), ("cp", C.c_void_p),
                ("jc", C.c_void_p), ("ir", C.c_void_p), ("numx", C.c_void_p), ("idx_bytes", C.c_int), ("dtype", C.c_int)]


class _DcscOut(C.Structure):
    _fields_ = [("cp", C.c_void_p), ("jc", C.c_void_p), ("ir", C.c_void_p), ("numx", C.c_void_p), ("idx_bytes", C.c_int)]


class _MatInfo(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("nnz", C.c_int64), ("nzc", C.c_int64), ("dtype", C.c_int),
                ("device_bytes", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("flops", C.c_int64), ("nnz_out", C.c_int64), ("nzc_out", C.c_int64), ("tasks", C.c_int64),
                ("kernel_launches", C.c_int64), ("ms_setup", C.c_float), ("ms_symbolic", C.c_float),
                ("ms_numeric", C.c_float), ("ms_total", C.c_float),
                ("tasks_hash_warp", C.c_int64), ("tasks_hash_cta", C.c_int64), ("tasks_bitmap_smem", C.c_int64),
                ("tasks_bitmap_gmem", C.c_int64), ("flops_hash_warp", C.c_int64), ("flops_hash_cta", C.c_int64),
                ("flops_bitmap_smem", C.c_int64), ("flops_bitmap_gmem", C.c_int64),
                ("nnz_hash_warp", C.c_int64), ("nnz_hash_cta", C.c_int64), ("nnz_bitmap_smem", C.c_int64),
                ("nnz_bitmap_gmem", C.c_int64), ("ms_kernel", C.c_float * 16), ("flops_sym", C.c_int64 * 5),
                ("class_tasks", C.c_int64 * 16), ("class_flops", C.c_int64 * 16), ("class_nnz", C.c_int64 * 16)]

    KERNELS = ["sym_bitmap", "sym_regsort", "sym_hash_cta", "sym_hash_warp", "sym_hash_warp_small",
               "num_bitmap_gmem", "num_bitmap_smem", "num_hash_cta", "num_hash_warp", "num_hash_warp_small", "flop",
               "num_hash_warp_mid", "num_sacc_medium", "num_sacc_small", "sym_bitmap_small", "num_regsort"]

    def as_dict(self):
        skip = ("ms_kernel", "flops_sym", "class_tasks", "class_flops", "class_nnz")
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in skip}
        d["classes"] = {n: {"tasks": int(self.class_tasks[i]), "flops": int(self.class_flops[i]), "nnz": int(self.class_nnz[i]),
                            "ms": round(float(self.ms_kernel[i]), 4)}
                        for i, n in enumerate(self.KERNELS) if self.class_tasks[i] > 0}
        d["ms_kernel"] = {n: round(float(self.ms_kernel[i]), 4) for i, n in enumerate(self.KERNELS) if self.ms_kernel[i] > 0}
        d["flops_sym"] = [int(x) for x in self.flops_sym]
        return d


class DistStats(C.Structure):
    _fields_ = [("local", Stats), ("ms_bcast", C.c_float), ("ms_multiply", C.c_float), ("ms_merge", C.c_float),
                ("ms_fiber_exchange", C.c_float), ("ms_fiber_merge", C.c_float), ("ms_total", C.c_float),
                ("bytes_bcast", C.c_int64), ("bytes_fiber", C.c_int64), ("stages", C.c_int)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "local"}
        d["local"] = self.local.as_dict()
        return d


class PruneStats(C.Structure):
    _fields_ = [("nnz_in", C.c_int64), ("nnz_out", C.c_int64), ("nzc_out", C.c_int64), ("cols_recovered", C.c_int64),
                ("cols_selected", C.c_int64), ("cols_recovered_after_select", C.c_int64), ("ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MemEffStats(C.Structure):
    _fields_ = [("phases", C.c_int), ("flops", C.c_int64), ("nnz_unpruned", C.c_int64), ("nnz_out", C.c_int64),
                ("cols_recovered", C.c_int64), ("cols_selected", C.c_int64), ("cols_recovered_after_select", C.c_int64),
                ("ms_multiply", C.c_float), ("ms_prune", C.c_float), ("ms_total", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SlabResult(C.Structure):
    _fields_ = [("nnz", C.c_int64), ("nzc", C.c_int64), ("pattern_sum", C.c_uint64), ("value_sum", C.c_uint64)]


class Grid(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("layers", C.c_int), ("grid_rows", C.c_int),
                ("grid_cols", C.c_int), ("my_layer", C.c_int), ("my_row", C.c_int), ("my_col", C.c_int)]


_lib = None

# every symbol include/cbgpu.h declares: (name, restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "cbgpu_version": (C.c_int, []),
    "cbgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cbgpu_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "cbgpu_destroy": (C.c_int, [_P]),
    "cbgpu_last_error": (C.c_char_p, [_P]),
    "cbgpu_sync": (C.c_int, [_P]),
    "cbgpu_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "cbgpu_get_option": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "cbgpu_launch_count": (C.c_int64, [_P]),
    "cbgpu_mat_upload": (C.c_int, [_P, C.POINTER(_DcscView), C.POINTER(_P)]),
    "cbgpu_mat_from_device_csc": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, C.c_int, C.POINTER(_P)]),
    "cbgpu_mat_info": (C.c_int, [_P, C.POINTER(_MatInfo)]),
    "cbgpu_mat_download": (C.c_int, [_P, _P, C.POINTER(_DcscOut)]),
    "cbgpu_mat_download_coo": (C.c_int, [_P, _P, _P, _P, _P, C.c_int]),
    "cbgpu_mat_device_arrays": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "cbgpu_mat_free": (C.c_int, [_P, _P]),
    "cbgpu_mat_checksum": (C.c_int, [_P, _P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "cbgpu_mat_checksum_at": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "cbgpu_mat_colsplit": (C.c_int, [_P, _P, C.c_int, C.POINTER(_P)]),
    "cbgpu_mat_colslice": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "cbgpu_mat_colconcat": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P)]),
    "cbgpu_mat_submatrix": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "cbgpu_spgemm_local": (C.c_int, [_P, C.c_int, _P, _P, C.POINTER(_P), C.POINTER(Stats)]),
    "cbgpu_spgemm_symbolic": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cbgpu_semiring_types": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cbgpu_spgemm_local_host": (C.c_int, [_P, C.c_int, C.POINTER(_DcscView), C.POINTER(_DcscView), C.POINTER(_P), C.POINTER(Stats)]),
    "cbgpu_merge": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P), C.POINTER(Stats)]),
    "cbgpu_mcl_prune": (C.c_int, [_P, _P, C.c_double, C.c_int64, C.c_int64, C.c_double, C.POINTER(_P), C.POINTER(PruneStats)]),
    "cbgpu_memefficient_spgemm": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_double, C.c_int64, C.c_int64, C.c_double,
                                            C.POINTER(_P), C.POINTER(MemEffStats)]),
    "cbgpu_calculate_phases": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "cbgpu_mat_make_col_stochastic": (C.c_int, [_P, _P]),
    "cbgpu_mat_inflate": (C.c_int, [_P, _P, C.c_double]),
    "cbgpu_grid_make": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(Grid)]),
    "cbgpu_grid_make_ccgrid": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(Grid)]),
    "cbgpu_block_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cbgpu_block_owner": (C.c_int, [C.c_int64, C.c_int, C.c_int64]),
    "cbgpu_grid_local_range": (C.c_int, [C.POINTER(Grid), C.c_int64, C.c_int64, C.c_int] + [C.POINTER(C.c_int64)] * 4),
    "cbgpu_nccl_unique_id": (C.c_int, [_P]),
    "cbgpu_comm_create": (C.c_int, [_P, C.POINTER(Grid), _P, C.POINTER(_P)]),
    "cbgpu_comm_destroy": (C.c_int, [_P]),
    "cbgpu_summa2d": (C.c_int, [_P, _P, C.c_int, _P, _P, C.POINTER(_P), C.POINTER(DistStats)]),
    "cbgpu_summa3d": (C.c_int, [_P, _P, C.c_int, _P, _P, C.POINTER(_P), C.POINTER(DistStats)]),
    "cbgpu_summa_phased": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(SlabResult), C.POINTER(DistStats)]),
    "cbgpu_memefficient_spgemm_dist": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, C.c_double, C.c_int64, C.c_int64, C.c_double,
                                                 C.POINTER(_P), C.POINTER(MemEffStats), C.POINTER(DistStats)]),
    "cbgpu_phase_columns": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cbgpu_spgemm_symbolic_columns": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "cbgpu_mat_transpose": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "cbgpu_mat_validate": (C.c_int, [_P, _P]),
    "cbgpu_redistribute": (C.c_int, [_P, _P, _P, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "cbgpu_memory_in_use": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "cbgpu_summa_symbolic": (C.c_int, [_P, _P, C.c_int, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "cbgpu_summa_phased_global": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, C.c_int64, C.c_int64, C.POINTER(_P), C.POINTER(SlabResult),
                                            C.POINTER(DistStats)]),
    "cbgpu_rmat_edges_host": (C.c_int, [C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_int, _P, _P]),
    "cbgpu_gen_rmat": (C.c_int, [_P, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.c_int, C.POINTER(_P)]),
    "cbgpu_gen_rmat_block": (C.c_int, [_P, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                       C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_P)]),
}


def load_library():
    """Load libcbgpu.so. Raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise CbgpuError(-100, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DeviceMatrix:
    """A DCSC block resident in HBM (opaque cbgpu_mat handle)."""

    def __init__(self, ctx: "Context", handle):
        self.ctx = ctx
        self.handle = handle

    def info(self):
        inf = _MatInfo()
        self.ctx._check(self.ctx.lib.cbgpu_mat_info(self.handle, C.byref(inf)))
        return inf

    @property
    def shape(self):
        i = self.info()
        return (i.m, i.n)

    @property
    def nnz(self):
        return self.info().nnz

    @property
    def nzc(self):
        return self.info().nzc

    @property
    def dtype(self):
        return self.info().dtype

    def free(self):
        if self.handle is not None and self.ctx.handle is not None:
            self.ctx.lib.cbgpu_mat_free(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One per (process, GPU). Raises if no CUDA device is usable."""

    def __init__(self, device: int = 0, stream=None):
        self.lib = load_library()
        h = _P()
        rc = self.lib.cbgpu_create(device, stream, C.byref(h))
        if rc != 0:
            raise CbgpuError(rc, f"cbgpu_create(device={device}) failed: no usable CUDA device; there is no CPU fallback")
        self.handle = h
        self.device = device

    def close(self):
        if self.handle is not None:
            self.lib.cbgpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CbgpuError(rc, self.lib.cbgpu_last_error(self.handle).decode(errors="replace"))

    def sync(self):
        self._check(self.lib.cbgpu_sync(self.handle))

    def set_option(self, name: str, value: int):
        self._check(self.lib.cbgpu_set_option(self.handle, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int64()
        self._check(self.lib.cbgpu_get_option(self.handle, name.encode(), C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        return int(self.lib.cbgpu_launch_count(self.handle))

    # ---- staging
    @staticmethod
    def _view(m, n, jc, cp, ir, numx):
        idx_bytes = ir.dtype.itemsize if len(ir) else jc.dtype.itemsize if len(jc) else 8
        jc = np.ascontiguousarray(jc)
        cp = np.ascontiguousarray(cp)
        ir = np.ascontiguousarray(ir)
        numx = np.ascontiguousarray(numx)
        for a in (jc, cp, ir):
            if len(a) and a.dtype.itemsize != idx_bytes:
                raise ValueError("index arrays must share one integer width")
        v = _DcscView(m, n, len(ir), len(jc), cp.ctypes.data, jc.ctypes.data, ir.ctypes.data, numx.ctypes.data, idx_bytes,
                      NUMPY_TO_DTYPE[numx.dtype])
        return v, (jc, cp, ir, numx)

    def upload(self, M) -> DeviceMatrix:
        """M: host SpDCCols (combblas_b200.host)."""
        v, keep = self._view(M.m, M.n, M.jc, M.cp, M.ir, M.numx)
        h = _P()
        self._check(self.lib.cbgpu_mat_upload(self.handle, C.byref(v), C.byref(h)))
        del keep
        return DeviceMatrix(self, h)

    def from_device_csc(self, m, n, nnz, colptr_ptr, rows_ptr, vals_ptr, dtype) -> DeviceMatrix:
        h = _P()
        self._check(self.lib.cbgpu_mat_from_device_csc(self.handle, m, n, nnz, colptr_ptr, rows_ptr, vals_ptr, dtype, C.byref(h)))
        return DeviceMatrix(self, h)

    def download(self, D: DeviceMatrix, idx_dtype=np.int64):
        """-> (m, n, jc, cp, ir, numx) numpy arrays (DCSC, Dcsc layout of dcsc.h:125-132)."""
        inf = D.info()
        idx_dtype = np.dtype(idx_dtype)
        jc = np.empty(inf.nzc, dtype=idx_dtype)
        cp = np.zeros(inf.nzc + 1, dtype=idx_dtype)
        ir = np.empty(inf.nnz, dtype=idx_dtype)
        numx = np.empty(inf.nnz, dtype=DTYPE_TO_NUMPY[inf.dtype])
        o = _DcscOut(cp.ctypes.data, jc.ctypes.data, ir.ctypes.data, numx.ctypes.data, idx_dtype.itemsize)
        self._check(self.lib.cbgpu_mat_download(self.handle, D.handle, C.byref(o)))
        return inf.m, inf.n, jc, cp, ir, numx

    def download_coo(self, D: DeviceMatrix, idx_dtype=np.int64):
        inf = D.info()
        idx_dtype = np.dtype(idx_dtype)
        rows = np.empty(inf.nnz, dtype=idx_dtype)
        cols = np.empty(inf.nnz, dtype=idx_dtype)
        vals = np.empty(inf.nnz, dtype=DTYPE_TO_NUMPY[inf.dtype])
        self._check(self.lib.cbgpu_mat_download_coo(self.handle, D.handle, rows.ctypes.data, cols.ctypes.data,
                                                    vals.ctypes.data, idx_dtype.itemsize))
        return rows, cols, vals

    def checksum(self, D: DeviceMatrix, row_offset: int = 0, col_offset: int = 0):
        """order-independent (pattern, value) sums; with offsets, of the block placed at that position of a larger matrix"""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self.lib.cbgpu_mat_checksum_at(self.handle, D.handle, row_offset, col_offset, C.byref(a), C.byref(b)))
        return a.value, b.value

    def colslice(self, D: DeviceMatrix, c0: int, c1: int) -> DeviceMatrix:
        h = _P()
        self._check(self.lib.cbgpu_mat_colslice(self.handle, D.handle, c0, c1, C.byref(h)))
        return DeviceMatrix(self, h)

    def submatrix(self, D: DeviceMatrix, r0: int, r1: int, c0: int, c1: int) -> DeviceMatrix:
        h = _P()
        self._check(self.lib.cbgpu_mat_submatrix(self.handle, D.handle, r0, r1, c0, c1, C.byref(h)))
        return DeviceMatrix(self, h)

    def colsplit(self, D: DeviceMatrix, parts: int):
        arr = (_P * parts)()
        self._check(self.lib.cbgpu_mat_colsplit(self.handle, D.handle, parts, arr))
        return [DeviceMatrix(self, _P(arr[i])) for i in range(parts)]

    def colconcat(self, mats):
        arr = (_P * len(mats))(*[m.handle for m in mats])
        h = _P()
        self._check(self.lib.cbgpu_mat_colconcat(self.handle, len(mats), arr, C.byref(h)))
        return DeviceMatrix(self, h)

    # ---- compute
    def spgemm(self, sr: int, A: DeviceMatrix, B: DeviceMatrix, want_stats=False):
        h = _P()
        st = Stats()
        self._check(self.lib.cbgpu_spgemm_local(self.handle, sr, A.handle, B.handle, C.byref(h), C.byref(st)))
        out = DeviceMatrix(self, h)
        return (out, st) if want_stats else out

    def spgemm_host(self, sr: int, A, B, want_stats=False):
        va, ka = self._view(A.m, A.n, A.jc, A.cp, A.ir, A.numx)
        vb, kb = self._view(B.m, B.n, B.jc, B.cp, B.ir, B.numx)
        h = _P()
        st = Stats()
        self._check(self.lib.cbgpu_spgemm_local_host(self.handle, sr, C.byref(va), C.byref(vb), C.byref(h), C.byref(st)))
        out = DeviceMatrix(self, h)
        return (out, st) if want_stats else out

    def symbolic(self, A: DeviceMatrix, B: DeviceMatrix):
        f, z = C.c_int64(), C.c_int64()
        self._check(self.lib.cbgpu_spgemm_symbolic(self.handle, A.handle, B.handle, C.byref(f), C.byref(z)))
        return f.value, z.value

    def validate(self, D: DeviceMatrix) -> None:
        """structural check of a resident block on the device (cbgpu_mat_validate); raises CbgpuError naming what is wrong"""
        self._check(self.lib.cbgpu_mat_validate(self.handle, D.handle))

    def transpose(self, D: DeviceMatrix) -> DeviceMatrix:
        h = _P()
        self._check(self.lib.cbgpu_mat_transpose(self.handle, D.handle, C.byref(h)))
        return DeviceMatrix(self, h)

    def memory_in_use(self) -> int:
        v = C.c_int64()
        self._check(self.lib.cbgpu_memory_in_use(self.handle, C.byref(v)))
        return v.value

    def symbolic_columns(self, A: DeviceMatrix, B: DeviceMatrix):
        """(products, outputs) per NON-EMPTY column of B: what estimateFLOP / estimateNNZ_Hash return (mtSpGEMM.h:1058, :807)"""
        nzc = B.info().nzc
        flop = np.zeros(nzc, dtype=np.int64)
        nnz = np.zeros(nzc, dtype=np.int64)
        f, z = C.c_int64(), C.c_int64()
        self._check(self.lib.cbgpu_spgemm_symbolic_columns(self.handle, A.handle, B.handle, C.byref(f), C.byref(z),
                                                           flop.ctypes.data, nnz.ctypes.data))
        return flop, nnz

    def merge(self, sr: int, mats, want_stats=False):
        arr = (_P * len(mats))(*[m.handle for m in mats])
        h = _P()
        st = Stats()
        self._check(self.lib.cbgpu_merge(self.handle, sr, len(mats), arr, C.byref(h), C.byref(st)))
        out = DeviceMatrix(self, h)
        return (out, st) if want_stats else out

    def mcl_prune(self, A: DeviceMatrix, hard_threshold: float, select_num: int, recover_num: int, recover_pct: float,
                  want_stats=False):
        """MCLPruneRecoverySelect (ParFriends.h:186-354) of a block holding whole columns -> new DeviceMatrix."""
        h = _P()
        st = PruneStats()
        self._check(self.lib.cbgpu_mcl_prune(self.handle, A.handle, C.c_double(hard_threshold), C.c_int64(select_num),
                                             C.c_int64(recover_num), C.c_double(recover_pct), C.byref(h), C.byref(st)))
        out = DeviceMatrix(self, h)
        return (out, st) if want_stats else out

    def memefficient_spgemm(self, sr: int, A: DeviceMatrix, B: DeviceMatrix, phases: int, hard_threshold: float, select_num: int,
                            recover_num: int, recover_pct: float, want_stats=False):
        """MemEfficientSpGEMM (ParFriends.h:452-777) at P = 1 on resident operands; phases <= 0 = automatic."""
        h = _P()
        st = MemEffStats()
        self._check(self.lib.cbgpu_memefficient_spgemm(self.handle, sr, A.handle, B.handle, int(phases), C.c_double(hard_threshold),
                                                       C.c_int64(select_num), C.c_int64(recover_num), C.c_double(recover_pct),
                                                       C.byref(h), C.byref(st)))
        out = DeviceMatrix(self, h)
        return (out, st) if want_stats else out

    def make_col_stochastic(self, A: DeviceMatrix):
        self._check(self.lib.cbgpu_mat_make_col_stochastic(self.handle, A.handle))

    def inflate(self, A: DeviceMatrix, power: float):
        self._check(self.lib.cbgpu_mat_inflate(self.handle, A.handle, C.c_double(power)))

    def gen_rmat_block(self, scale, nedges, seed, r0, r1, c0, c1, a=0.57, b=0.19, c=0.19, scramble=True, dtype=F64, value_mode=0):
        """block [r0,r1) x [c0,c1) of gen_rmat's matrix, local indices, without materialising the whole matrix"""
        h = _P()
        self._check(self.lib.cbgpu_gen_rmat_block(self.handle, scale, nedges, seed, a, b, c, int(scramble), dtype, value_mode,
                                                  r0, r1, c0, c1, C.byref(h)))
        return DeviceMatrix(self, h)

    def gen_rmat(self, scale, nedges, seed, a=0.57, b=0.19, c=0.19, scramble=True, dtype=F64, value_mode=0):
        h = _P()
        self._check(self.lib.cbgpu_gen_rmat(self.handle, scale, nedges, seed, a, b, c, int(scramble), dtype, value_mode, C.byref(h)))
        return DeviceMatrix(self, h)


OPTION_NAMES = ["bitmap_window_log2", "bitmap_min_nnz", "shared_acc", "shared_acc_max", "shared_acc_small_max", "bitmap_cta_threads", "bitmap_small_threads",
                "bitmap_small_minblocks", "bitmap_save_mb", "bitmap_save_min_flop", "light_max", "force_path", "merge_engine", "summa_fused", "fiber_fused", "fiber_pipeline", "regsort", "regsort_packed",
                "sacc_v2", "sacc_overflow", "merge_tma", "validate_uploads"]


class SlabPipeline:
    """Phased multiply C(:, slab) = A x B(:, slab) (MemEfficientSpGEMM's column phases, ParFriends.h:553-772) with the
    slabs spread over `nstreams` contexts of one GPU, each a host thread with its own CUDA stream. The slabs are
    independent, so the symbolic pass of one (instruction bound) overlaps the numeric pass of another (bound by the
    L2 reductions), and the host-side gaps of one stream (size read-backs, allocation) are covered by the other.
    The first slab runs alone: it builds the per-operand caches of A (dense column index, window-major copy)."""

    def __init__(self, ctx: Context, nstreams: int = 2):
        self.ctxs = [ctx] + [Context(ctx.device) for _ in range(max(1, nstreams) - 1)]
        self.sync_options()

    def sync_options(self):
        for name in OPTION_NAMES:
            v = self.ctxs[0].get_option(name)
            for c in self.ctxs[1:]:
                c.set_option(name, v)

    def launch_count(self) -> int:
        return sum(c.launch_count() for c in self.ctxs)

    def run(self, sr: int, A: DeviceMatrix, slabs, consume):
        """consume(index, C_slab, stats) is called on the worker thread that produced the slab and must free C_slab."""
        import threading

        self.ctxs[0].sync()  # everything the caller queued on the main stream (operands, L2 flush) is complete
        errors = []

        def work(ctx, items):
            try:
                for i in items:
                    Cs, st = ctx.spgemm(sr, A, slabs[i], want_stats=True)
                    consume(i, Cs, st)
            except Exception as e:  # surfaced on the calling thread
                errors.append(e)

        work(self.ctxs[0], [0])
        n = len(self.ctxs)
        rest = list(range(1, len(slabs)))
        threads = [threading.Thread(target=work, args=(self.ctxs[w], rest[w::n])) for w in range(n)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]


def make_grid(world: int, rank: int, layers: int = 1, ccgrid: bool = False) -> Grid:
    """CommGrid / CommGrid3D rank map, or (ccgrid=True) the one of the older 3D code (3DSpGEMM/CCGrid.h:14-17)."""
    g = Grid()
    fn = load_library().cbgpu_grid_make_ccgrid if ccgrid else load_library().cbgpu_grid_make
    rc = fn(world, rank, layers, C.byref(g))
    if rc != 0:
        raise CbgpuError(rc, f"no {layers}-layer square grid over {world} ranks (reference: NOTSQUARE / GRIDMISMATCH)")
    return g


class Comm:
    """NCCL communicators (world, row, column, fiber) of this rank; bootstrap via torch.distributed (plumbing)."""

    def __init__(self, ctx: Context, grid: Grid, unique_id: bytes):
        self.ctx = ctx
        self.grid = grid
        buf = C.create_string_buffer(unique_id, 128)
        h = _P()
        ctx._check(ctx.lib.cbgpu_comm_create(ctx.handle, C.byref(grid), buf, C.byref(h)))
        self.handle = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load_library().cbgpu_nccl_unique_id(buf)
        if rc != 0:
            raise CbgpuError(rc, "ncclGetUniqueId failed (NCCL not loadable?)")
        return buf.raw

    def summa2d(self, sr, A: DeviceMatrix, B: DeviceMatrix):
        h = _P()
        st = DistStats()
        self.ctx._check(self.ctx.lib.cbgpu_summa2d(self.ctx.handle, self.handle, sr, A.handle, B.handle, C.byref(h), C.byref(st)))
        return DeviceMatrix(self.ctx, h), st

    def summa3d(self, sr, A: DeviceMatrix, B: DeviceMatrix):
        h = _P()
        st = DistStats()
        self.ctx._check(self.ctx.lib.cbgpu_summa3d(self.ctx.handle, self.handle, sr, A.handle, B.handle, C.byref(h), C.byref(st)))
        return DeviceMatrix(self.ctx, h), st

    def memefficient_spgemm(self, sr, A: DeviceMatrix, B: DeviceMatrix, phases, hard_threshold, select_num, recover_num, recover_pct):
        """MemEfficientSpGEMM / MemEfficientSpGEMM3D with the distributed pruning epilogue; returns (pruned block of C, stats, dist stats)"""
        h = _P()
        ms, ds = MemEffStats(), DistStats()
        self.ctx._check(self.ctx.lib.cbgpu_memefficient_spgemm_dist(self.ctx.handle, self.handle, sr, A.handle, B.handle, int(phases),
                                                                    C.c_double(hard_threshold), int(select_num), int(recover_num),
                                                                    C.c_double(recover_pct), C.byref(h), C.byref(ms), C.byref(ds)))
        return DeviceMatrix(self.ctx, h), ms, ds

    def redistribute(self, D: DeviceMatrix, source, target):
        """source / target: one (r0, r1, c0, c1) rectangle of global indices per rank -- what every rank holds now and shall hold
        afterwards (SpParMat3D 2D -> 3D constructor / Convert2D, SpParMat3D.cpp:187, :441); returns (block, bytes moved)"""
        src = np.ascontiguousarray(np.asarray(source, dtype=np.int64).reshape(-1))
        dst = np.ascontiguousarray(np.asarray(target, dtype=np.int64).reshape(-1))
        world = len(src) // 4
        h = _P()
        moved = C.c_int64()
        self.ctx._check(self.ctx.lib.cbgpu_redistribute(self.ctx.handle, self.handle, D.handle, src.ctypes.data, dst.ctypes.data, world,
                                                        self.grid.rank, C.byref(h), C.byref(moved)))
        return DeviceMatrix(self.ctx, h), moved.value

    def summa_symbolic(self, sr, A: DeviceMatrix, B: DeviceMatrix):
        """(products, outputs) this rank produces in the distributed product (exact; EstPerProcessNnzSUMMA's role)"""
        f, z = C.c_int64(), C.c_int64()
        self.ctx._check(self.ctx.lib.cbgpu_summa_symbolic(self.ctx.handle, self.handle, sr, A.handle, B.handle, C.byref(f), C.byref(z)))
        return f.value, z.value

    def summa_phased(self, sr, A: DeviceMatrix, B: DeviceMatrix, phases: int, want_checksum=False, keep=False, global_offsets=None):
        """MemEfficientSpGEMM[3D]-style phased multiply; returns (slab results, kept slabs or None, stats).
        global_offsets = (first global row of this rank's A/C block, first global column of its B block): the slab checksums
        are then taken at their global positions and add up, over slabs and ranks, to the checksum of the whole product."""
        res = (SlabResult * phases)()
        st = DistStats()
        arr = (_P * phases)() if keep else None
        if global_offsets is not None:
            self.ctx._check(self.ctx.lib.cbgpu_summa_phased_global(self.ctx.handle, self.handle, sr, A.handle, B.handle, phases,
                                                                   int(global_offsets[0]), int(global_offsets[1]), arr, res, C.byref(st)))
        else:
            self.ctx._check(self.ctx.lib.cbgpu_summa_phased(self.ctx.handle, self.handle, sr, A.handle, B.handle, phases,
                                                            int(want_checksum), arr, res, C.byref(st)))
        kept = [DeviceMatrix(self.ctx, _P(arr[i])) for i in range(phases)] if keep else None
        return list(res), kept, st

    def destroy(self):
        if self.handle is not None:
            self.ctx.lib.cbgpu_comm_destroy(self.handle)
            self.handle = None
