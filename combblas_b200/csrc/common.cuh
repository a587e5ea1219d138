// Shared internals of libcbgpu.so (sm_100a). Not part of the public ABI (include/cbgpu.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/cbgpu.h"

namespace cbgpu {

constexpr int kSMs = 148; // B200: 2 dies x 74 SMs; grids are sized in multiples of this where it matters
constexpr unsigned kEmptyKey = 0xFFFFFFFFu;

struct Options {
  // symbolic (keys only) hash limits, in products per task
  int64_t sym_warp_small = 32;   // warp table 64
  int64_t sym_warp_max = 256;    // warp table 512
  int64_t sym_cta_small = 2048;  // CTA(256) table 4096
  int64_t sym_cta_max = 16384;   // CTA(512) table 32768
  // numeric (key+value) hash limits, in output entries per task
  int64_t num_warp_small = 32;   // warp table 64
  int64_t num_warp_max = 256;    // warp table 512
  int64_t num_cta_max = 2048;    // CTA(256) table 4096
  // bitmap path
  int64_t bitmap_window_log2 = 17; // rows per window (2^17 rows = 23 KiB of ranked 45-row cells)
  int64_t bitmap_min_nnz = 0;      // 0 = automatic: clamp(window_rows/2048, 32, num_cta_max)
  int64_t shared_acc = 1;          // 1 = bitmap tasks whose outputs fit accumulate in shared memory (exchange protocol), 0 = all in C (L2 reductions)
  int64_t shared_acc_max = 0;      // > 0: upper limit on the outputs of a shared-accumulator task (default: what the CTA shapes hold)
  int64_t shared_acc_small_max = -1; // >= 0: upper limit on the outputs of the small (256-thread) shape
  int64_t bitmap_cta_threads = 512; // CTA size of the large-task bitmap kernels (512 or 256)
  int64_t bitmap_small_threads = 128; // CTA size of the small-task bitmap kernels (128 or 256)
  int64_t bitmap_save_mb = 8192;   // HBM budget (MiB) for the presence words the symbolic pass hands to the numeric pass; 0 = off
  int64_t bitmap_save_min_flop = 2048; // tasks with at least this many products are handed over (4 bytes per 32 rows of the window; bulk copies both ways: 8192 -> 2048 is 20 ms per step at R-MAT scale 22)
  int64_t light_max = 256;         // with several row windows, columns up to this many products stay one task (<= 2048)
  int64_t bitmap_small_minblocks = 8; // resident CTAs per SM the 128-thread numeric bitmap kernel is compiled for (8 or 12)
  int64_t regsort = 1;             // tasks with <= 256 products and segments sorted in registers (regsort_kernel): 1 = numeric pass, 2 = symbolic pass too, 0 = per-warp hash classes
  int64_t regsort_packed = 1;      // numeric register sort on row << log2(capacity) | position where the block has fewer than 2^24 rows (values fetched after the sort); 0 = key + value through the network
  int64_t force_path = 0;          // debugging: 1 = hash only (where it fits), 2 = bitmap only
  int64_t summa_fused = 1;         // 1 = all SUMMA stages as one stacked local multiply, 0 = stage loop + merge
  int64_t fiber_fused = 1;         // 3D: replicate the inputs along the fiber instead of reducing partial results (see dist.cu); 0 = the reference's fiber reduction
  int64_t fiber_pipeline = 0;      // fiber_fused == 0 only: second host thread + stream overlaps the fiber reduction of slab p with the multiply of slab p+1
  int64_t merge_engine = 0;        // 1 = k-way merges through the accumulation engine instead of streaming 2-way rounds
  int64_t sacc_overflow = 4;       // > 1: the large shared-accumulator shape takes tasks of up to this many times its capacity (at most 65532 outputs); the outputs beyond the capacity accumulate in C with L2 reductions. 1 = such tasks go to num_bitmap_kernel whole
  int64_t validate_uploads = 0;    // 1 = every cbgpu_mat_upload checks the block on the device (cbgpu_mat_validate) before handing it out
  int64_t merge_tma = 1;           // streaming 2-way merge: 1 = persistent CTAs with double-buffered bulk (TMA) tile copies, 0 = one tile per CTA with load/store loops
  int64_t sacc_v2 = 15;            // shared-accumulator numeric classes, second version (16-bit ranks, vector scan, bulk hand-over): bit 0 small, 1 medium, 2 large shape; bit 3: the large shape keeps a row array too, bit 4: the medium shape does not (measured best at R-MAT scale 22: 31)
};

} // namespace cbgpu

// the opaque handles of include/cbgpu.h
struct cbgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string last_error;
  cbgpu::Options opt;
  int64_t launches = 0;
  int sm_count = cbgpu::kSMs;
  int max_smem_optin = 0;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t kev[2 * CBGPU_K_COUNT] = {};
};

struct cbgpu_mat {
  int64_t m = 0, n = 0, nnz = 0, nzc = 0;
  int dtype = CBGPU_F64;
  int64_t *jc = nullptr;     // [nzc]
  int64_t *cp = nullptr;     // [nzc+1]
  int32_t *ir = nullptr;     // [nnz]
  void *numx = nullptr;      // [nnz]
  int64_t *colptr = nullptr; // [n+1] dense column index, built on demand (replaces Dcsc::ConstructAux/FillColInds)
  // window-major copy for use as the A operand (built on demand, keyed by the window size)
  int win_log2 = 0, win_nwin = 0;
  int64_t *win_T2 = nullptr; // [nwin*n + 1]
  int32_t *win_ir = nullptr; // [nnz]
  void *win_val = nullptr;   // [nnz]
  int device = 0;
};

namespace cbgpu {
typedef ::cbgpu_ctx cbgpu_ctx_impl;
typedef ::cbgpu_mat cbgpu_mat_impl;

inline size_t dtype_size(int dt) {
  switch (dt) {
    case CBGPU_F64: return 8;
    case CBGPU_F32: return 4;
    case CBGPU_I64: return 8;
    case CBGPU_I32: return 4;
    case CBGPU_BOOL: return 1;
  }
  return 0;
}

int set_error(cbgpu_ctx_impl *ctx, int code, const char *fmt, ...);

#define CB_CUDA(ctx, expr)                                                                                             \
  do {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess) {                                                                                           \
      return ::cbgpu::set_error((ctx), _e == cudaErrorMemoryAllocation ? CBGPU_ERR_NOMEM : CBGPU_ERR_CUDA,             \
                                "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e));           \
    }                                                                                                                  \
  } while (0)

#define CB_TRY(expr)                                                                                                   \
  do {                                                                                                                 \
    int _rc = (expr);                                                                                                  \
    if (_rc != CBGPU_OK) return _rc;                                                                                   \
  } while (0)

#define CB_LAUNCH_CHECK(ctx)                                                                                           \
  do {                                                                                                                 \
    (ctx)->launches++;                                                                                                 \
    CB_CUDA(ctx, cudaGetLastError());                                                                                  \
  } while (0)

// stream-ordered allocation helpers (cudaMallocAsync pool; no implicit device sync)
int dev_alloc(cbgpu_ctx_impl *ctx, void **p, size_t bytes);
int dev_free(cbgpu_ctx_impl *ctx, void *p);
void release_cached_blocks(cbgpu_ctx_impl *ctx);
int pool_live_bytes(cbgpu_ctx_impl *ctx, int64_t *live);
template <class T>
inline int dev_alloc_t(cbgpu_ctx_impl *ctx, T **p, size_t count) {
  return dev_alloc(ctx, reinterpret_cast<void **>(p), count * sizeof(T));
}

// device-wide primitives (util.cu)
int exclusive_scan_i64(cbgpu_ctx_impl *ctx, const int64_t *in, int64_t *out, int64_t n); // out has n+1 entries
int fill_i64(cbgpu_ctx_impl *ctx, int64_t *p, int64_t n, int64_t v);
int ensure_dense_colptr(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *M);
int ensure_window_major(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *M, int nwin, int wlog2);
int mat_alloc(cbgpu_ctx_impl *ctx, int64_t m, int64_t n, int64_t nnz, int64_t nzc, int dtype, cbgpu_mat_impl **out);
int mat_release(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *M);
// builds DCSC (jc, cp) of the non-empty columns from per-column counts over `ncols_in` candidate columns
int compact_columns(cbgpu_ctx_impl *ctx, const int64_t *cand_ids /*may be null: identity*/, const int64_t *cand_ptr,
                    int64_t ncand, int64_t **jc, int64_t **cp, int64_t *nzc);

// ---- ownership guards: every early return gives device temporaries and half-built results back
struct MatGuard { // releases a result block unless the call succeeds
  cbgpu_ctx_impl *ctx;
  cbgpu_mat_impl **m;
  bool armed = true;
  MatGuard(cbgpu_ctx_impl *c, cbgpu_mat_impl **mm) : ctx(c), m(mm) {}
  ~MatGuard() {
    if (armed && *m) mat_release(ctx, *m);
  }
};
struct Scratch {
  cbgpu_ctx_impl *ctx;
  std::vector<void *> ptrs;
  explicit Scratch(cbgpu_ctx_impl *c) : ctx(c) {}
  ~Scratch() {
    for (void *p : ptrs) dev_free(ctx, p);
  }
  template <class T>
  int alloc(T **p, size_t count) {
    int rc = dev_alloc_t(ctx, p, count);
    if (rc == CBGPU_OK) ptrs.push_back(*p);
    return rc;
  }
  void detach(void *p) { // ownership moves elsewhere
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == p) {
        ptrs.erase(ptrs.begin() + i);
        return;
      }
  }
  void release(void *p) { // early release
    if (!p) return;
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == p) {
        ptrs.erase(ptrs.begin() + i);
        dev_free(ctx, p);
        return;
      }
  }
};


// per-semiring entry points generated from accumulate.cuh (one translation unit per semiring)
struct SpgemmArgs {
  cbgpu_ctx_impl *ctx;
  cbgpu_mat_impl *A, *B;
  cbgpu_mat_impl **C; // null: symbolic only
  cbgpu_stats *stats;
  int64_t *flops_out, *nnz_out;
  int64_t *col_flops_host = nullptr, *col_nnz_host = nullptr; // per non-empty column of B (symbolic only)
};
struct MergeArgs {
  cbgpu_ctx_impl *ctx;
  int k;
  cbgpu_mat_impl *const *lists;
  cbgpu_mat_impl **out;
  cbgpu_stats *stats;
};
typedef int (*spgemm_fn)(const SpgemmArgs &);
typedef int (*merge_fn)(const MergeArgs &);
spgemm_fn spgemm_entry(int semiring);
merge_fn merge_entry(int semiring);
int semiring_types(int semiring, int *a, int *b, int *c);
// user-defined semirings: returns the new id (>= CBGPU_SR_USER_BASE) or a negative status
int register_user_semiring(spgemm_fn spgemm, merge_fn merge, int a_dtype, int b_dtype, int c_dtype);

} // namespace cbgpu
