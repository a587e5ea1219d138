/*
 * cbgpu.h -- C ABI of libcbgpu.so: B200-native (sm_100a) semiring SpGEMM hot path behind the
 * CombBLAS SpDCCols / SpTuples / semiring interface.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary. Every entry point returns
 * CBGPU_OK (0) or a negative cbgpu_status and never aborts the process; the C++ overlay
 * (include/combblas_b200/overlay) maps failures onto the reference's MPI_Abort codes (SpDefs.h:72-78).
 *
 * Each entry cites the reference interface (path:line under the CombBLAS tree) that it replaces.
 * The reference has no FFI of its own -- its seam is C++ templates -- so INTEGRATION.md shows the
 * constrained-overload binding a maintainer adds on the reference side.
 *
 * Threading: one host thread per context at a time. All work of a context is ordered on its stream.
 */
#ifndef CBGPU_H
#define CBGPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CBGPU_VERSION 100

typedef enum {
  CBGPU_OK = 0,
  CBGPU_ERR_INVALID = -1,     /* bad argument */
  CBGPU_ERR_CUDA = -2,        /* CUDA runtime failure; text in cbgpu_last_error */
  CBGPU_ERR_NOMEM = -3,       /* device allocation failed */
  CBGPU_ERR_UNSUPPORTED = -4, /* type/semiring/size combination outside the device path */
  CBGPU_ERR_DIMMISMATCH = -5, /* inner dimensions differ (reference: DIMMISMATCH 3002, SpDefs.h:73) */
  CBGPU_ERR_NCCL = -6,        /* NCCL failure or NCCL not loadable */
  CBGPU_ERR_GRID = -7         /* process grid not usable (reference: GRIDMISMATCH 3001 / NOTSQUARE 3003) */
} cbgpu_status;

/* value types of matrix entries; bool travels as one byte (C++ bool) */
typedef enum { CBGPU_F64 = 0, CBGPU_F32 = 1, CBGPU_I64 = 2, CBGPU_I32 = 3, CBGPU_BOOL = 4 } cbgpu_dtype;

/*
 * Library semirings (Semirings.h:143-255; OR-AND as in ReleaseTests/KTipsTest.cpp:12-20).
 * Each fixes the operand types (A, B) and the output type.
 */
typedef enum {
  CBGPU_SR_PLUS_TIMES_F64 = 0,      /* PlusTimesSRing<double,double>   f64 x f64 -> f64 */
  CBGPU_SR_PLUS_TIMES_F32 = 1,      /* PlusTimesSRing<float,float>     f32 x f32 -> f32 */
  CBGPU_SR_PLUS_TIMES_I64 = 2,      /* PlusTimesSRing<int64,int64>     i64 x i64 -> i64 */
  CBGPU_SR_SELECT_MAX_BOOL_I64 = 3, /* SelectMaxSRing<bool,int64_t>    bool x i64 -> i64 */
  CBGPU_SR_MIN_PLUS_F64 = 4,        /* MinPlusSRing<double,double>     f64 x f64 -> f64 */
  CBGPU_SR_OR_AND_BOOL = 5,         /* user OR-AND semiring on bool    bool x bool -> bool */
  CBGPU_SR_PLUS_TIMES_BOOL_F64 = 6, /* PlusTimesSRing<bool,double>     bool x f64 -> f64 */
  CBGPU_SR_PLUS_TIMES_I32 = 7,      /* PlusTimesSRing<int32,int32>     i32 x i32 -> i32 */
  CBGPU_SR_SELECT_MAX_I64 = 8,      /* SelectMaxSRing<int64,int64>     i64 x i64 -> i64 */
  /* BoolCopy2ndSRing<OUT> / BoolCopy1stSRing<OUT> (Semirings.h:51-138): the pair SpParMat::SubsRef_SR multiplies with
   * (SpParMat.cpp:2515-2566). multiply copies the non-boolean operand; add must not happen (the reference throws): a product
   * or merge in which an output would receive two values fails with CBGPU_ERR_INVALID. */
  CBGPU_SR_BOOL_COPY_2ND_F64 = 9,   /* bool x f64 -> f64 */
  CBGPU_SR_BOOL_COPY_1ST_F64 = 10,  /* f64 x bool -> f64 */
  CBGPU_SR_BOOL_COPY_2ND_I64 = 11,  /* bool x i64 -> i64 */
  CBGPU_SR_BOOL_COPY_1ST_I64 = 12,  /* i64 x bool -> i64 */
  CBGPU_SR_BOOL_COPY_2ND_BOOL = 13, /* bool x bool -> bool */
  CBGPU_SR_BOOL_COPY_1ST_BOOL = 14, /* bool x bool -> bool */
  CBGPU_SR_COUNT = 15,
  /* ids from here on belong to user-defined semirings: structs with the reference's semiring interface (static id / add /
   * multiply, Semirings.h:143-255) whose members are __host__ __device__, instantiated into the engine by a translation
   * unit of the application (include/combblas_b200/device_semiring.cuh) and registered when that unit is loaded */
  CBGPU_SR_USER_BASE = 64
} cbgpu_semiring;

typedef struct cbgpu_ctx cbgpu_ctx; /* one per (process, GPU): stream, workspace, tunables */
typedef struct cbgpu_mat cbgpu_mat; /* device-resident DCSC block (dcsc.h:125-132), rows ascending per column */

/*
 * Host- or device-side view of an SpDCCols<IT,NT> block: the four arrays of Dcsc (dcsc.h:125-132)
 * plus the "essentials" {nnz, m, n, nzc} (SpDCCols.cpp:788 GetEssentials). idx_bytes is sizeof(IT), 4 or 8.
 * A zero matrix has nnz == nzc == 0 and may pass NULL arrays (SpDCCols keeps dcsc == NULL then).
 */
typedef struct {
  int64_t m, n, nnz, nzc;
  const void *cp;   /* nzc+1 column pointers      (IT) */
  const void *jc;   /* nzc   non-empty column ids (IT) */
  const void *ir;   /* nnz   row ids              (IT) */
  const void *numx; /* nnz   values               (dtype) */
  int idx_bytes;
  int dtype; /* cbgpu_dtype */
} cbgpu_dcsc_view;

/* writable counterpart used for downloads: caller allocates arrays sized from cbgpu_mat_info */
typedef struct {
  void *cp, *jc, *ir, *numx;
  int idx_bytes;
} cbgpu_dcsc_out;

typedef struct {
  int64_t m, n, nnz, nzc;
  int dtype;
  int64_t device_bytes;
} cbgpu_mat_info_t;

/* counters and stream-timed phases of the last multiply / merge (milliseconds, CUDA events) */
typedef struct {
  int64_t flops;     /* products = sum_j sum_{k in B(:,j)} nnz(A(:,k)); reference: estimateFLOP mtSpGEMM.h:1058 */
  int64_t nnz_out;   /* nnz(C) */
  int64_t nzc_out;   /* non-empty columns of C */
  int64_t tasks;     /* (column, row-window) work items */
  int64_t kernel_launches;
  float ms_setup;    /* dense column index, row windows, flop count, binning */
  float ms_symbolic; /* distinct-row counting (estimateNNZ_Hash mtSpGEMM.h:807) + scan */
  float ms_numeric;  /* accumulation + sorted emission */
  float ms_total;
  int64_t tasks_hash_warp, tasks_hash_cta, tasks_bitmap_smem, tasks_bitmap_gmem; /* numeric path census */
  int64_t flops_hash_warp, flops_hash_cta, flops_bitmap_smem, flops_bitmap_gmem;
  int64_t nnz_hash_warp, nnz_hash_cta, nnz_bitmap_smem, nnz_bitmap_gmem; /* outputs written per numeric path */
  /* device time of each kernel class of this call (CUDA events on the context's stream), see CBGPU_K_* */
  float ms_kernel[16];
  int64_t flops_sym[5]; /* products walked by the symbolic classes CBGPU_K_SYM_* */
  /* per kernel class (CBGPU_K_*): tasks, products and outputs it handled in this call */
  int64_t class_tasks[16], class_flops[16], class_nnz[16];
} cbgpu_stats;

enum {
  CBGPU_K_SYM_BITMAP = 0, CBGPU_K_SYM_HASH_CTA_L = 1, CBGPU_K_SYM_HASH_CTA = 2, CBGPU_K_SYM_HASH_WARP = 3,
  CBGPU_K_SYM_HASH_WARP_S = 4, CBGPU_K_NUM_BITMAP_GMEM = 5, CBGPU_K_NUM_BITMAP_SMEM = 6, CBGPU_K_NUM_HASH_CTA = 7,
  CBGPU_K_NUM_HASH_WARP = 8, CBGPU_K_NUM_HASH_WARP_S = 9, CBGPU_K_FLOP = 10, CBGPU_K_NUM_HASH_WARP_M = 11,
  /* CBGPU_K_NUM_BITMAP_SMEM is the large (1024-thread) shape of the shared-accumulator kernel; medium and small: */
  CBGPU_K_NUM_SACC_M = 12, CBGPU_K_NUM_SACC_S = 13, CBGPU_K_SYM_BITMAP_S = 14,
  /* register-sort classes (tasks with <= 256 products and segments): symbolic reuses the retired slot 1 */
  CBGPU_K_SYM_REGSORT = 1, CBGPU_K_NUM_REGSORT = 15,
  CBGPU_K_COUNT = 16
};

/* ---------------------------------------------------------------- lifecycle */
int cbgpu_version(void);
int cbgpu_device_count(int *count);
/* stream may be NULL (the context creates its own non-blocking stream) or a cudaStream_t to adopt */
int cbgpu_create(int device, void *stream, cbgpu_ctx **ctx);
int cbgpu_destroy(cbgpu_ctx *ctx);
const char *cbgpu_last_error(const cbgpu_ctx *ctx);
int cbgpu_sync(cbgpu_ctx *ctx);
/* device memory the library holds for live objects right now (the device's stream-ordered pool minus the reusable blocks parked
 * in the large-block cache); equal before and after any call that fails */
int cbgpu_memory_in_use(cbgpu_ctx *ctx, int64_t *live_bytes);
/* tunables (all have working defaults; unknown names are rejected): "bitmap_window_log2" rows per window of the bitmap path,
 * "bitmap_min_nnz" smallest task the bitmap path takes, "light_max" products up to which a column stays one task,
 * "shared_acc" / "shared_acc_max" / "shared_acc_small_max" shared-memory accumulator classes, "bitmap_save_mb" /
 * "bitmap_save_min_flop" symbolic -> numeric hand-over, "bitmap_cta_threads", "bitmap_small_threads", "force_path" (tests:
 * 1 hash only, 2 bitmap only), "regsort" (1: tasks with <= 256 products and segments sorted in registers) / "regsort_packed"
 * (row and staging position sorted as one 32-bit key), "sacc_v2" (bit per CTA shape: second version of the shared-accumulator
 * kernels) / "sacc_overflow" (the large shape takes tasks of up to this many times its capacity, the rest of their outputs accumulate in C), "bitmap_small_minblocks",
 * "merge_engine", "merge_tma" (streaming merge with bulk tile copies), "validate_uploads", "summa_fused", "fiber_fused",
 * "fiber_pipeline" */
int cbgpu_set_option(cbgpu_ctx *ctx, const char *name, int64_t value);
int cbgpu_get_option(cbgpu_ctx *ctx, const char *name, int64_t *value);
/* kernels launched by this context so far (bench.py's gpu_launches) */
int64_t cbgpu_launch_count(const cbgpu_ctx *ctx);

/* ---------------------------------------------------------------- DCSC staging in HBM
 * replaces: SpDCCols::CreateImpl / GetArrays (SpDCCols.cpp:735,:827) as the thing BCastMatrix moves,
 * and the tuples->DCSC constructor SpDCCols(const SpTuples&, bool) (SpDCCols.cpp:110-189). */
int cbgpu_mat_upload(cbgpu_ctx *ctx, const cbgpu_dcsc_view *host, cbgpu_mat **out);
/* Structural check of a resident block on the device: column pointers ascending from 0 to nnz, listed columns non-empty,
 * column ids ascending and < n, row ids in [0, m) and strictly ascending inside every column -- what the engine's window
 * searches and the streaming merge assume, and what reference blocks from the sort=false paths (mtSpGEMM.h:434) or a
 * hand-built SpDCCols may violate. CBGPU_ERR_INVALID + a message naming what is wrong. Option "validate_uploads" = 1 runs
 * it inside every cbgpu_mat_upload (the overlay sets it when CBGPU_VALIDATE is in the environment). */
int cbgpu_mat_validate(cbgpu_ctx *ctx, const cbgpu_mat *M);
/* adopt DEVICE arrays laid out as plain CSC: colptr int64[n+1], rows int32[nnz] ascending per column,
 * vals dtype[nnz]. Arrays are copied into library-owned storage (the caller keeps its buffers). */
int cbgpu_mat_from_device_csc(cbgpu_ctx *ctx, int64_t m, int64_t n, int64_t nnz, const int64_t *colptr,
                              const int32_t *rows, const void *vals, int dtype, cbgpu_mat **out);
int cbgpu_mat_info(const cbgpu_mat *mat, cbgpu_mat_info_t *info);
int cbgpu_mat_download(cbgpu_ctx *ctx, const cbgpu_mat *mat, const cbgpu_dcsc_out *host);
/* column-major COO (rows ascending per column) as three separate host arrays; idx_bytes 4 or 8.
 * This is what the overlay turns into SpTuples (SpTuples.h:64; std::tuple layout is the overlay's business). */
int cbgpu_mat_download_coo(cbgpu_ctx *ctx, const cbgpu_mat *mat, void *rows, void *cols, void *vals, int idx_bytes);
/* raw device pointers of a resident block (jc int64[nzc], cp int64[nzc+1], ir int32[nnz], numx dtype[nnz]) */
int cbgpu_mat_device_arrays(const cbgpu_mat *mat, const int64_t **jc, const int64_t **cp, const int32_t **ir,
                            const void **numx);
int cbgpu_mat_free(cbgpu_ctx *ctx, cbgpu_mat *mat);
/* order-independent 64-bit checksum of (row, col, value bits) over all entries; used for slab-wise parity */
int cbgpu_mat_checksum(cbgpu_ctx *ctx, const cbgpu_mat *mat, uint64_t *pattern_sum, uint64_t *value_sum);
/* the same sums with the block placed at (row_offset, col_offset) of a larger matrix: the checksums of the blocks of a
 * distributed matrix, each taken at its global position, add up (mod 2^64) to the checksum of the whole matrix -- this is how
 * a distributed product is compared with the single-GPU product and with the reference (SpParMat::operator==,
 * SpParMat.cpp:2891, compares block by block on equal grids; the sum is grid independent). Global rows below 2^32. */
int cbgpu_mat_checksum_at(cbgpu_ctx *ctx, const cbgpu_mat *mat, int64_t row_offset, int64_t col_offset, uint64_t *pattern_sum,
                          uint64_t *value_sum);
/* ColSplit / ColConcatenate of B and C slabs (dcsc.cpp:1202-1277, :1317-1360; SpDCCols.cpp:1054-1263) */
int cbgpu_mat_colsplit(cbgpu_ctx *ctx, const cbgpu_mat *mat, int parts, cbgpu_mat **out /* parts */);
int cbgpu_mat_colslice(cbgpu_ctx *ctx, const cbgpu_mat *mat, int64_t col_begin, int64_t col_end, cbgpu_mat **out);
int cbgpu_mat_colconcat(cbgpu_ctx *ctx, int parts, cbgpu_mat *const *in, cbgpu_mat **out);
/* block [row_begin,row_end) x [col_begin,col_end) with local indices: what the 2D/3D distributions hand each rank
 * (SpParMat::Owner SpParMat.cpp:5081; SpParMat3D ctor SpParMat3D.cpp:187-283 does this through an all-to-all) */
int cbgpu_mat_submatrix(cbgpu_ctx *ctx, const cbgpu_mat *mat, int64_t row_begin, int64_t row_end, int64_t col_begin,
                        int64_t col_end, cbgpu_mat **out);

/* transpose on the device: replaces SpDCCols::Transpose / TransposeConst (SpDCCols.cpp:871-905) for resident operands, so that
 * chains like the Galerkin product R^T A R never stage through the host. The result has rows ascending in every column. */
int cbgpu_mat_transpose(cbgpu_ctx *ctx, const cbgpu_mat *mat, cbgpu_mat **out);

/* ---------------------------------------------------------------- local multiply (K1-K4)
 * replaces: LocalHybridSpGEMM (mtSpGEMM.h:213-460), LocalSpGEMMHash (:463-656), LocalSpGEMM (:74-202)
 * and, fused in, estimateFLOP (:1058), estimateNNZ_Hash (:807), prefixsum (:24) and the
 * tuples->DCSC conversion the SUMMA drivers do afterwards (ParFriends.h:1549).
 * C = A (x) B; output block is DCSC with rows ascending inside every column. */
int cbgpu_spgemm_local(cbgpu_ctx *ctx, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, cbgpu_mat **C,
                       cbgpu_stats *stats);
/* symbolic only: total products and nnz(C) (EstimateFLOP ParFriends.h:357; estimateNNZ_Hash) */
int cbgpu_spgemm_symbolic(cbgpu_ctx *ctx, const cbgpu_mat *A, const cbgpu_mat *B, int64_t *flops, int64_t *nnz_out);
/* the same with the per-column results the reference's symbolic functions return: col_flops[j] (estimateFLOP, mtSpGEMM.h:1058)
 * and col_nnz[j] (estimateNNZ_Hash, :807) for the j-th NON-EMPTY column of B (nzc(B) entries each, HOST arrays, either may be NULL) */
int cbgpu_spgemm_symbolic_columns(cbgpu_ctx *ctx, const cbgpu_mat *A, const cbgpu_mat *B, int64_t *flops, int64_t *nnz_out,
                                  int64_t *col_flops, int64_t *col_nnz);
/* host-buffers-in, host-buffers-out form of the same call (what a LocalHybridSpGEMM overlay invokes):
 * uploads both operands, multiplies, and hands back a resident result to download. */
int cbgpu_spgemm_local_host(cbgpu_ctx *ctx, int semiring, const cbgpu_dcsc_view *A, const cbgpu_dcsc_view *B,
                            cbgpu_mat **C, cbgpu_stats *stats);

/* operand and result value types (cbgpu_dtype) of a library or registered user semiring; CBGPU_ERR_INVALID for an unknown id */
int cbgpu_semiring_types(int semiring, int *a_dtype, int *b_dtype, int *c_dtype);

/* ---------------------------------------------------------------- k-way merge (K7/K8)
 * replaces: MultiwayMerge (MultiwayMerge.h:428-543) and MultiwayMergeHash (:553-701): column-wise union of k
 * blocks of identical shape with SR::add on equal (row, col). k == 1 returns a copy. */
int cbgpu_merge(cbgpu_ctx *ctx, int semiring, int k, const cbgpu_mat *const *lists, cbgpu_mat **out,
                cbgpu_stats *stats);

/* ---------------------------------------------------------------- MCL pruning epilogue of the phased multiply
 * replaces: MCLPruneRecoverySelect (ParFriends.h:186-354), which MemEfficientSpGEMM runs on every column slab of C
 * (:744), for a block that holds WHOLE columns (one rank per process column): A.Prune(val <= hardThreshold), the column
 * sums / counts (Reduce(Column)), Kselect1 (SpParMat.cpp:1413-1700) and PruneColumn(thresholds, std::less) fused into
 * one per-column pass. Column thresholds: hardThreshold; the recoverNum-th largest entry for columns left with fewer
 * than recoverNum entries summing below recoverPct; the selectNum-th largest for columns left with more than selectNum
 * entries (with the reference's second recovery check after selection). Entries below their column's threshold go.
 * Floating-point blocks only (the reference instantiates it for float / double). */
typedef struct {
  int64_t nnz_in, nnz_out, nzc_out;
  int64_t cols_recovered, cols_selected, cols_recovered_after_select;
  float ms;
} cbgpu_prune_stats;
int cbgpu_mcl_prune(cbgpu_ctx *ctx, const cbgpu_mat *A, double hardThreshold, int64_t selectNum, int64_t recoverNum,
                    double recoverPct, cbgpu_mat **out, cbgpu_prune_stats *stats);
/* The phased multiply with its pruning epilogue on one GPU. replaces: MemEfficientSpGEMM (ParFriends.h:452-777) at P = 1:
 * B is cut into `phases` column slabs (ColSplit rule, dcsc.cpp:1202), every slab of C = A (x) B(:, slab) is pruned in HBM by
 * cbgpu_mcl_prune before the next one is multiplied, the pruned slabs are concatenated (ColConcatenate, :772).
 * phases <= 0: chosen from the exact symbolic nnz(C) so that an unpruned slab fits a quarter of the free HBM
 * (CalculateNumberOfPhases, ParFriends.h:780-843). Results that are not floating point are multiplied but not pruned. */
typedef struct {
  int phases;
  int64_t flops, nnz_unpruned, nnz_out;
  int64_t cols_recovered, cols_selected, cols_recovered_after_select;
  float ms_multiply, ms_prune, ms_total;
} cbgpu_memeff_stats;
int cbgpu_memefficient_spgemm(cbgpu_ctx *ctx, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                              double hardThreshold, int64_t selectNum, int64_t recoverNum, double recoverPct, cbgpu_mat **C,
                              cbgpu_memeff_stats *stats);
/* Host arithmetic of CalculateNumberOfPhases (ParFriends.h:779-832), no GPU needed: phases = 1 + asquareMem / remainingMem
 * with asquareMem = nnz_product_per_process * (2 idx_bytes + out_val_bytes) * 2 and remainingMem = per_process_memory_gb
 * * 1e9 - max_local_nnz_a * (2 idx_bytes + in_val_bytes) * 4. The reference estimates nnz_product_per_process
 * (EstPerProcessNnzSUMMA, :1698); here cbgpu_spgemm_symbolic supplies the exact count. Returns the phase count (>= 1),
 * or CBGPU_ERR_INVALID when the inputs alone exceed the memory. */
int cbgpu_calculate_phases(int64_t max_local_nnz_a, int64_t nnz_product_per_process, int idx_bytes, int in_val_bytes,
                           int out_val_bytes, int64_t per_process_memory_gb);
/* in place: every column scaled to sum 1 (MakeColStochastic, Applications/MCL.cpp:389-394) */
int cbgpu_mat_make_col_stochastic(cbgpu_ctx *ctx, cbgpu_mat *A);
/* in place: v = pow(v, power), then MakeColStochastic (Inflate, Applications/MCL.cpp:431-437) */
int cbgpu_mat_inflate(cbgpu_ctx *ctx, cbgpu_mat *A, double power);

/* ---------------------------------------------------------------- process grids and distributed SpGEMM
 * Host-side arithmetic of the reference's distributions (no GPU needed):
 * 2D owner rule SpParMat::Owner (SpParMat.cpp:5081-5107), 3D split SpParMat3D::Owner/LocalDim
 * (SpParMat3D.cpp:337-436), layer column split CalculateColSplitDistributionOfLayer (:576-609),
 * rank maps CommGrid (CommGrid.h:106, src/CommGrid.cpp:57-58) and CommGrid3D (CommGrid3D.h:75-93). */
typedef struct {
  int world, rank;
  int layers;    /* c; 1 for a plain 2D grid */
  int grid_rows; /* pr == pc inside a layer */
  int grid_cols;
  int my_layer, my_row, my_col; /* position of `rank` */
} cbgpu_grid;
int cbgpu_grid_make(int world, int rank, int layers, cbgpu_grid *grid);
/* same grid shape with the rank map of the older 3D code path (3DSpGEMM/CCGrid.h:14-17: layer = rank % c,
 * rank in layer = rank / c); the communicators of cbgpu_comm_create follow whichever map built the grid */
int cbgpu_grid_make_ccgrid(int world, int rank, int layers, cbgpu_grid *grid);
/* half-open range [begin,end) of the global dimension `dim` owned by block `index` of `parts` (last takes remainder) */
int cbgpu_block_range(int64_t dim, int parts, int index, int64_t *begin, int64_t *end);
int cbgpu_block_owner(int64_t dim, int parts, int64_t global_index);
/* local ranges of the A (column-split) / B (row-split) / C (column-split) block of `grid`'s rank for an m x n matrix */
int cbgpu_grid_local_range(const cbgpu_grid *grid, int64_t m, int64_t n, int split_cols /*1: A,C  0: B*/,
                           int64_t *row_begin, int64_t *row_end, int64_t *col_begin, int64_t *col_end);

/* columns [begin,end) of a rank's block of B (n local columns) that phase `phase` of a phased multiply takes from the chunk of
 * fiber rank `layer`: one layer = the ColSplit slab (dcsc.cpp:1202); several = piece `phase` of chunk `layer`
 * (MemEfficientSpGEMM3D, ParFriends.h:3774-3811; chunks by CalculateColSplitDistributionOfLayer, SpParMat3D.cpp:576-609). Host only. */
int cbgpu_phase_columns(int64_t n, int phases, int layers, int phase, int layer, int64_t *begin, int64_t *end);

typedef struct cbgpu_comm cbgpu_comm; /* NCCL communicators of one rank: world, row, column, fiber */
int cbgpu_nccl_unique_id(void *id128 /* 128 bytes out */);
int cbgpu_comm_create(cbgpu_ctx *ctx, const cbgpu_grid *grid, const void *id128, cbgpu_comm **comm);
int cbgpu_comm_destroy(cbgpu_comm *comm);

typedef struct {
  cbgpu_stats local; /* summed over stages */
  float ms_bcast, ms_multiply, ms_merge, ms_fiber_exchange, ms_fiber_merge, ms_total;
  int64_t bytes_bcast, bytes_fiber;
  int stages;
} cbgpu_dist_stats;

/* 2D Sparse SUMMA: replaces Mult_AnXBn_Synch (ParFriends.h:1447-1556) incl. GetSetSizes/BCastMatrix
 * (SpParHelper.cpp:798,:583) and the final MultiwayMerge. A, B, C are this rank's blocks. */
int cbgpu_summa2d(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B,
                  cbgpu_mat **C, cbgpu_dist_stats *stats);
/* 3D SUMMA: replaces Mult_AnXBn_SUMMA3D (ParFriends.h:3374-3667): per-layer 2D SUMMA, fiber all-to-all of
 * column slabs (:3578-3612), fiber merge (:3642). A column-split, B row-split, C column-split across layers. */
int cbgpu_summa3d(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B,
                  cbgpu_mat **C, cbgpu_dist_stats *stats);

/* Redistribution between block layouts on the device: replaces the 2D -> 3D constructor SpParMat3D(const SpParMat&, nlayers,
 * colsplit, special) (SpParMat3D.cpp:187-283) and SpParMat3D::Convert2D (:441-570), both built on ExchangeData (:51, :97).
 * Every rank passes its block and the rectangles {row_begin, row_end, col_begin, col_end} (global indices) that EVERY rank holds
 * now (source[4 * r ...]) and shall hold afterwards (target[4 * r ...]); cbgpu_grid_make + cbgpu_grid_local_range give them for
 * the 2D and both 3D layouts. The pieces travel device to device (grouped ncclSend / ncclRecv over the world communicator). */
int cbgpu_redistribute(cbgpu_ctx *ctx, cbgpu_comm *comm, const cbgpu_mat *local, const int64_t *source, const int64_t *target,
                       int world, int rank, cbgpu_mat **out, int64_t *bytes_moved);

/* Distributed symbolic pass: products and outputs THIS rank produces in the distributed product (final distribution of C).
 * replaces: EstPerProcessNnzSUMMA (ParFriends.h:1698) and the estimate loop of CalculateNumberOfPhases (:780-843); exact. */
int cbgpu_summa_symbolic(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int64_t *flops,
                         int64_t *nnz_out);

/* Phased distributed multiply: replaces MemEfficientSpGEMM (ParFriends.h:453-777; without the pruning, which stays with
 * the caller) and MemEfficientSpGEMM3D (:3674-4170). B's local columns are cut into `phases` slabs (ColSplit rule), one
 * SUMMA per slab. With several layers the fiber exchange + merge of slab p overlaps the multiply of slab p+1 (second
 * host thread, second stream). results[p] receives the essentials (and, if asked, the order-independent checksums) of
 * slab p; slabs == NULL means every slab is consumed (freed) as soon as it is finished, otherwise slabs[p] is handed out. */
typedef struct {
  int64_t nnz, nzc;
  uint64_t pattern_sum, value_sum;
} cbgpu_slab_result;
int cbgpu_summa_phased(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                       int want_checksum, cbgpu_mat **slabs, cbgpu_slab_result *results, cbgpu_dist_stats *stats);
/* the same with every slab's checksums taken at its GLOBAL position: row_offset = first global row of this rank's block of
 * A / C, col_offset = first global column of this rank's block of B (before the cut into phases and, with layers, into the
 * fiber's column sub-slabs, whose offsets the library adds). Summed over slabs and ranks they equal the checksums of the
 * whole product on any grid. */
int cbgpu_summa_phased_global(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                              int64_t row_offset, int64_t col_offset, cbgpu_mat **slabs, cbgpu_slab_result *results,
                              cbgpu_dist_stats *stats);

/* The phased distributed multiply WITH its pruning epilogue: replaces MemEfficientSpGEMM (ParFriends.h:452-777) on a 2D grid and
 * MemEfficientSpGEMM3D (:3673-4170, pruning :4148) on a layered one. Every finished piece of C is pruned by
 * MCLPruneRecoverySelect (:186-354) over the WHOLE distributed columns -- the pieces of a process column are re-cut so that every
 * rank holds whole columns of a share, pruned by the single-GPU kernels, and sent back (this replaces the column reductions and the
 * distributed Kselect1 of SpParMat.cpp:1413-1700) -- before the next slab is multiplied; the pruned pieces are concatenated into
 * this rank's block of C in the layout of Mult_AnXBn_Synch / Mult_AnXBn_SUMMA3D. phases <= 0: from the distributed symbolic pass. */
int cbgpu_memefficient_spgemm_dist(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                                   double hardThreshold, int64_t selectNum, int64_t recoverNum, double recoverPct, cbgpu_mat **C,
                                   cbgpu_memeff_stats *stats, cbgpu_dist_stats *dist_stats);

/* ---------------------------------------------------------------- synthetic inputs (own seeded generators)
 * R-MAT (Graph500 initiator a,b,c,d as in 3DSpGEMM/mpipspgemm.cpp:126-133), duplicates summed into the value
 * (SpTuples.cpp:70-124 semantics), vertex scramble by a seeded bijection. Device generation; the identical
 * arithmetic is available on the host for parity inputs (cbgpu_rmat_edges_host). Not part of the timed path. */
int cbgpu_rmat_edges_host(int scale, int64_t nedges, uint64_t seed, double a, double b, double c, int scramble,
                          int64_t *rows, int64_t *cols);
/* value_mode: 0 = multiplicity of the edge (duplicates summed), 1 = one (duplicates collapsed), 2 = 1 + row id */
int cbgpu_gen_rmat(cbgpu_ctx *ctx, int scale, int64_t nedges, uint64_t seed, double a, double b, double c,
                   int scramble, int dtype, int value_mode, cbgpu_mat **out);
/* the block [row_begin,row_end) x [col_begin,col_end) of the same matrix with block-local indices, built without ever holding
 * the whole matrix: every rank of a grid generates the edge stream and keeps what its block owns (what the reference does with
 * DistEdgeList::GenGraph500Data + the SpParMat(DistEdgeList) all-to-all, DistEdgeList.cpp:223, SpParMat.cpp:3153; and for
 * 3D blocks the SpParMat3D constructor, SpParMat3D.cpp:187-283) */
int cbgpu_gen_rmat_block(cbgpu_ctx *ctx, int scale, int64_t nedges, uint64_t seed, double a, double b, double c, int scramble,
                         int dtype, int value_mode, int64_t row_begin, int64_t row_end, int64_t col_begin, int64_t col_end,
                         cbgpu_mat **out);

#ifdef __cplusplus
}
#endif
#endif /* CBGPU_H */
