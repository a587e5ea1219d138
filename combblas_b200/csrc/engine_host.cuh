// Host-side orchestration of the accumulation engine (engine.cuh): task construction, size/window ordering,
// symbolic pass, scan, numeric pass, DCSC assembly. Instantiated once per semiring (sr_instance.cu).
#pragma once
#include <algorithm>
#include "engine.cuh"
#include "util.cuh"

namespace cbgpu {

__device__ __forceinline__ int size_bucket(int64_t v) {
  if (v <= 0) return 0;
  if (v == 1) return 1;
  int b = 1 + (64 - __clzll((unsigned long long)(v - 1)));
  return b > 39 ? 39 : b;
}

__device__ __forceinline__ bool task_is_one_window(const uint32_t *task_win, int nwin, int64_t i) {
  if (nwin <= 1 || !task_win) return true;
  unsigned w = task_win[i];
  return ((w >> 16) - (w & 0xFFFFu)) == 1;
}

// symbolic path choice per task: bucket = size bucket of the products (+0 hash, +40 bitmap)
// Register-sort classes (engine.cuh regsort_kernel): tasks with at most 8 G products and at most 8 G segments, G = 8 / 16 / 32
// lanes per task; buckets 240 / 241 / 242. Both passes classify by the same two numbers, so a task keeps its class.
struct TinyRule {
  const int32_t *task_col; // null: task i is column i
  const int64_t *Bcp;      // multiply: column pointers of B (segments of a task = entries of its column of B); null: merge
  int merge_k;             // merge: segments per task
  int enabled;
};
__device__ __forceinline__ int tiny_bucket(const TinyRule &r, int64_t i, int64_t flop) {
  if (!r.enabled || flop < 1 || flop > 256) return 0;
  int64_t nseg = r.merge_k;
  if (r.Bcp) {
    const int64_t c = r.task_col ? r.task_col[i] : i;
    nseg = r.Bcp[c + 1] - r.Bcp[c];
  }
  if (nseg > 256) return 0;
  const int64_t m = flop > nseg ? flop : nseg;
  return m <= 64 ? 240 : (m <= 128 ? 241 : 242);
}

static __global__ void sym_bucket_kernel(const int64_t *flop, const uint32_t *task_win, int nwin, int64_t ntask,
                                         int64_t bitmap_min_flop, int force_path, TinyRule tiny, uint8_t *bucket) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntask) return;
  int64_t v = flop[i];
  if (const int tb = tiny_bucket(tiny, i, v)) {
    bucket[i] = (uint8_t)tb;
    return;
  }
  int sb = size_bucket(v);
  bool can_bitmap = task_is_one_window(task_win, nwin, i);
  bool use_bitmap = can_bitmap && (v >= bitmap_min_flop || v > 2048);
  if (force_path == 1 && v <= 2048) use_bitmap = false;
  if (force_path == 2 && can_bitmap) use_bitmap = true;
  bucket[i] = (uint8_t)((sb == 0 || !use_bitmap) ? sb : sb + 40);
}

// numeric path choice per task: bucket = size bucket of the outputs (+0 hash, +80 bitmap with accumulators in C itself,
// +120 / +160 / +200 bitmap with shared-memory accumulators in the small / medium / large CTA shape)
static __global__ void num_bucket_kernel(const int64_t *nnz, const uint32_t *task_win, int nwin, int64_t ntask,
                                         int64_t bitmap_min_nnz, int64_t hash_max_nnz, int64_t cap_s, int64_t cap_m,
                                         int64_t cap_l, int force_path, TinyRule tiny, const int64_t *flop, uint8_t *bucket) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntask) return;
  int64_t v = nnz[i];
  int sb = size_bucket(v);
  if (sb == 0) {
    bucket[i] = 0;
    return;
  }
  if (const int tb = tiny_bucket(tiny, i, flop[i])) {
    bucket[i] = (uint8_t)tb;
    return;
  }
  bool can_bitmap = task_is_one_window(task_win, nwin, i);
  bool use_bitmap = can_bitmap && (v >= bitmap_min_nnz || v > hash_max_nnz);
  if (force_path == 1 && v <= hash_max_nnz) use_bitmap = false;
  if (force_path == 2 && can_bitmap) use_bitmap = true;
  if (!use_bitmap) bucket[i] = (uint8_t)sb;
  else if (v <= cap_s) bucket[i] = (uint8_t)(sb + 120);
  else if (v <= cap_m) bucket[i] = (uint8_t)(sb + 160);
  else if (v <= cap_l) bucket[i] = (uint8_t)(sb + 200);
  else bucket[i] = (uint8_t)(sb + 80);
}

static __global__ void col_task_count_kernel(const int64_t *colflop, int64_t ncol, int nwin, int64_t lightmax, int64_t *cnt) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ncol) cnt[j] = colflop[j] <= lightmax ? 1 : nwin;
}

static __global__ void fill_tasks_kernel(const int64_t *first, int64_t ncol, int nwin, int32_t *task_col, uint32_t *task_win) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncol) return;
  int64_t b = first[j], e = first[j + 1];
  if (e - b == 1) {
    task_col[b] = (int32_t)j;
    task_win[b] = ((unsigned)nwin << 16);
  } else {
    for (int w = 0; w < nwin; ++w) {
      task_col[b + w] = (int32_t)j;
      task_win[b + w] = ((unsigned)(w + 1) << 16) | (unsigned)w;
    }
  }
}

static __global__ void gather_ptr_kernel(const int64_t *taskptr, const int64_t *first, int64_t ncol, int64_t *colptr_out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j <= ncol) colptr_out[j] = taskptr[first[j]];
}

// static: an application's own semiring unit (device_semiring.cuh) carries its own copy of every host helper that takes a kernel
// pointer, whatever CUDA runtime instance it was linked with (a kernel handle is only valid in the runtime that registered it)
template <class K>
static inline int optin_smem(cbgpu_ctx_impl *ctx, K kernel, size_t bytes) {
  // static + dynamic shared memory above 48 KiB needs the opt-in; the kernels here also carry static workspaces
  if (bytes > 16 * 1024) CB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CBGPU_OK;
}

// every device temporary of a call is owned by a guard: whatever path leaves the function -- in particular a failed
// allocation of C, the case the phased multiply exists to recover from -- gives all of them back to the pool
#define CB_KBEGIN(id) do { used[id] = true; cudaEventRecord(ctx->kev[2 * (id)], st); } while (0)
#define CB_KEND(id) cudaEventRecord(ctx->kev[2 * (id) + 1], st)

// row windows of an m-row block under the context's options
inline void engine_windows(const cbgpu_ctx_impl *ctx, int64_t m, int *nwin, int *wlog2) {
  int wl = (int)ctx->opt.bitmap_window_log2;
  if (wl < 10) wl = 10;
  if (wl > 19) wl = 19;
  const int64_t W = (int64_t)1 << wl;
  *wlog2 = wl;
  *nwin = (int)std::max<int64_t>(1, (m + W - 1) / W);
}

struct EngineIO {
  // A side (or the concatenated merge lists)
  const int64_t *Acolptr;
  int64_t ncolA;
  int64_t m;
  // task space
  int64_t ncol;               // multiply: nzc(B); merge: n
  const int64_t *out_col_ids; // multiply: B.jc; merge: null (identity)
  int64_t n_out;              // columns of C
  int out_dtype;
  // results
  cbgpu_mat_impl **C; // null: symbolic only
  cbgpu_stats *stats;
  int64_t *flops_out, *nnz_out;
  int64_t *col_flops_host, *col_nnz_host; // optional host arrays [ncol]: products / outputs per non-empty column of B
};

// symbolic kernel classes (launch order) and numeric kernel classes
enum { SYM_BM_L = 0, SYM_BM_S = 1, SYM_H_CTA = 2, SYM_H_WARP = 3, SYM_H_WARP_S = 4, SYM_RS_8 = 5, SYM_RS_16 = 6, SYM_RS_32 = 7 };
enum { NUM_BM_G = 0, NUM_SA_L = 1, NUM_SA_M = 2, NUM_H_CTA = 3, NUM_H_WARP = 4, NUM_H_WARP_M2 = 5, NUM_H_WARP_M = 6, NUM_H_WARP_S = 7, NUM_SA_S = 8, NUM_RS_8 = 9, NUM_RS_16 = 10, NUM_RS_32 = 11 };

// shape of the shared-accumulator numeric classes: CTA size and resident CTAs per SM (32 warps per SM in every shape)
constexpr int kSaccThreadsS = 256, kSaccBlocksS = 4, kSaccThreadsM = 512, kSaccBlocksM = 2, kSaccThreadsL = 1024, kSaccBlocksL = 1;
// dynamic shared memory one CTA of a shape can have: the SM's 228 KiB minus 1 KiB reserved per resident CTA, split evenly,
// minus the static workspace of the kernel (flat-walk queue + scan scratch)
inline int64_t sacc_dynamic_smem(const cbgpu_ctx_impl *ctx, int threads, int blocks, size_t queue_bytes) {
  const int64_t per_sm = 228 * 1024, per_cta = std::min<int64_t>((per_sm - 1024 * blocks) / blocks, ctx->max_smem_optin > 0 ? ctx->max_smem_optin : 227 * 1024);
  (void)threads;
  return ((per_cta - (int64_t)queue_bytes - 256) / 16) * 16;
}

// src must arrive with Air/Aval (whole columns) and, when the block has several row windows, T2/Wir/Wval set.
template <class SR, bool MERGE>
int run_engine(cbgpu_ctx_impl *ctx, Source<SR, MERGE> src, const EngineIO &io) {
  typedef typename SR::acc_t acc_t;
  typedef typename SR::out_t out_t;
  cudaStream_t st = ctx->stream;
  Scratch scratch(ctx);
  const Options &opt = ctx->opt;
  const int64_t launches0 = ctx->launches;
  cbgpu_stats stats;
  memset(&stats, 0, sizeof(stats));
  bool used[CBGPU_K_COUNT] = {};
  CB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));

  const int64_t ncol = io.ncol;
  if (ncol >= (int64_t)1 << 31 || io.m >= ((int64_t)1 << 31) - 1)
    return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "local dimensions must stay below 2^31 (m=%lld, columns=%lld)",
                     (long long)io.m, (long long)ncol);
  int nwin, wlog2;
  engine_windows(ctx, io.m, &nwin, &wlog2);
  if (nwin > 65535) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "too many row windows (%d)", nwin);
  const int64_t W = (int64_t)1 << wlog2;
  src.T = io.Acolptr;
  src.nwin = nwin;
  src.wlog2 = wlog2;
  src.N = io.ncolA;
  src.task_col = nullptr;
  src.task_win = nullptr;

  // ---- K1: products per column, then tasks
  int64_t *colflop = nullptr;
  CB_TRY(scratch.alloc(&colflop, (size_t)ncol + 1));
  if (ncol > 0) {
    CB_KBEGIN(CBGPU_K_FLOP);
    task_flop_kernel<SR, MERGE><<<(unsigned)((ncol + 7) / 8), 256, 0, st>>>(src, ncol, colflop);
    CB_LAUNCH_CHECK(ctx);
    CB_KEND(CBGPU_K_FLOP);
  }
  int64_t ntask = ncol;
  int64_t *taskflop = colflop;
  int64_t *first = nullptr; // first task of every column (nwin > 1)
  int32_t *task_col = nullptr;
  uint32_t *task_win = nullptr;
  if (nwin > 1 && ncol > 0) {
    int64_t *cnt = nullptr;
    CB_TRY(scratch.alloc(&cnt, (size_t)ncol));
    CB_TRY(scratch.alloc(&first, (size_t)ncol + 1));
    col_task_count_kernel<<<(unsigned)((ncol + 255) / 256), 256, 0, st>>>(colflop, ncol, nwin, std::min<int64_t>(std::max<int64_t>(opt.light_max, 1), 2048), cnt);
    CB_LAUNCH_CHECK(ctx);
    CB_TRY(exclusive_scan_i64(ctx, cnt, first, ncol));
    CB_CUDA(ctx, cudaMemcpyAsync(&ntask, first + ncol, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    scratch.release(cnt);
    if (ntask >= (int64_t)1 << 31) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "too many tasks");
    CB_TRY(scratch.alloc(&task_col, (size_t)ntask));
    CB_TRY(scratch.alloc(&task_win, (size_t)ntask));
    fill_tasks_kernel<<<(unsigned)((ncol + 255) / 256), 256, 0, st>>>(first, ncol, nwin, task_col, task_win);
    CB_LAUNCH_CHECK(ctx);
    src.task_col = task_col;
    src.task_win = task_win;
    CB_TRY(scratch.alloc(&taskflop, (size_t)ntask + 1));
    task_flop_kernel<SR, MERGE><<<(unsigned)((ntask + 7) / 8), 256, 0, st>>>(src, ntask, taskflop);
    CB_LAUNCH_CHECK(ctx);
  }
  stats.tasks = ntask;

  // ---- symbolic ordering
  uint8_t *bucket = nullptr;
  int32_t *order = nullptr;
  int64_t *tasknnz = nullptr, *taskptr = nullptr;
  CB_TRY(scratch.alloc(&bucket, (size_t)ntask + 1));
  CB_TRY(scratch.alloc(&order, (size_t)ntask + 1));
  CB_TRY(scratch.alloc(&tasknnz, (size_t)ntask + 1));
  CB_TRY(scratch.alloc(&taskptr, (size_t)ntask + 2));
  TaskRec *recs = nullptr; // launch-order records of the bitmap classes
  CB_TRY(scratch.alloc(&recs, (size_t)ntask + 1));
  BinResult bins;
  ClassRanges sc;
  memset(&bins, 0, sizeof(bins));
  memset(&sc, 0, sizeof(sc));
  uint8_t sym_class[256];
  for (int b = 0; b < 256; ++b) {
    uint8_t c = 15;
    if (b >= 1 && b <= 6) c = SYM_H_WARP_S;
    else if (b >= 7 && b <= 9) c = SYM_H_WARP;
    else if (b >= 10 && b <= 39) c = SYM_H_CTA;
    else if (b >= 41 && b <= 54) c = SYM_BM_S;
    else if (b >= 55 && b <= 79) c = SYM_BM_L;
    else if (b >= 240 && b <= 242) c = (uint8_t)(SYM_RS_8 + (b - 240));
    sym_class[b] = c;
  }
  TinyRule tiny;
  tiny.task_col = task_col;
  tiny.Bcp = MERGE ? nullptr : src.Bcp;
  tiny.merge_k = src.k;
  tiny.enabled = (opt.regsort && opt.force_path == 0 && (MERGE ? src.k <= 256 : true)) ? 1 : 0;
  TinyRule tiny_sym = tiny; // the symbolic pass keeps its per-warp tables unless asked (they only hold rows; measured faster)
  tiny_sym.enabled = tiny.enabled && opt.regsort >= 2;
  if (ntask > 0) {
    CB_CUDA(ctx, cudaMemsetAsync(tasknnz, 0, sizeof(int64_t) * (size_t)ntask, st));
    int64_t sym_min = std::min<int64_t>(std::max<int64_t>(std::min<int64_t>(io.m, W) / 256, 64), 2048);
    sym_bucket_kernel<<<(unsigned)((ntask + 255) / 256), 256, 0, st>>>(taskflop, task_win, nwin, ntask, sym_min,
                                                                      (int)opt.force_path, tiny_sym, bucket);
    CB_LAUNCH_CHECK(ctx);
    CB_TRY(bin_tasks(ctx, bucket, taskflop, nullptr, ntask, task_win, sym_class, order, &bins, &sc));
  }
  for (int b = 1; b < 256; ++b) stats.flops += bins.weight[b];
  CB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));

  // ---- K2: symbolic kernels
  const int max_words = (words_of_rows(std::min<int64_t>(io.m, W)) + 3) & ~3; // multiple of 4: uint4 clears and copies
  const size_t sym_bytes = (size_t)max_words * 4; // symbolic pass: presence words
  const size_t bm_bytes = (size_t)max_words * 8;  // numeric pass: presence words + rank prefixes
  // Presence words of the large symbolic tasks are kept in HBM for the numeric pass (which then needs no mark walk of
  // its own): one slot of max_words words per task, as many tasks as fit the budget.
  unsigned *saved = nullptr;
  int32_t *slot_of_task = nullptr;
  int save_count = 0;
  int *save_counter = nullptr;
  if (io.C && opt.bitmap_save_mb > 0 && sc.count[SYM_BM_L] + sc.count[SYM_BM_S] > 0) {
    const int64_t fit = (opt.bitmap_save_mb << 20) / (int64_t)sym_bytes;
    save_count = (int)std::min<int64_t>(sc.count[SYM_BM_L] + sc.count[SYM_BM_S], fit);
    if (save_count > 0) {
      // an optimisation only: when HBM is too full for the hand-over buffer the numeric pass marks and ranks again
      if (scratch.alloc(&saved, (size_t)save_count * max_words) != CBGPU_OK ||
          scratch.alloc(&slot_of_task, (size_t)ntask) != CBGPU_OK || scratch.alloc(&save_counter, 4) != CBGPU_OK) {
        scratch.release(saved);
        scratch.release(slot_of_task);
        scratch.release(save_counter);
        saved = nullptr;
        slot_of_task = nullptr;
        save_counter = nullptr;
        save_count = 0;
        ctx->last_error.clear();
      } else {
        CB_CUDA(ctx, cudaMemsetAsync(slot_of_task, 0xFF, sizeof(int32_t) * (size_t)ntask, st));
        CB_CUDA(ctx, cudaMemsetAsync(save_counter, 0, 4 * sizeof(int), st));
      }
    }
  }
  // the records of the symbolic launch, now that the hand-over slots can be dealt out
  if (ntask > 0 && bins.listed > 0) {
    task_record_kernel<SR, MERGE><<<(unsigned)((bins.listed + 255) / 256), 256, 0, st>>>(
        src, order, bins.listed, nullptr, nullptr, recs, taskflop, std::max<int64_t>(opt.bitmap_save_min_flop, 1), save_count,
        save_counter);
    CB_LAUNCH_CHECK(ctx);
  }
  auto class_weight = [](const BinResult &r, const uint8_t *tab, int c, bool second) {
    int64_t s = 0;
    for (int b = 0; b < 256; ++b)
      if (tab[b] == c) s += second ? r.weight2[b] : r.weight[b];
    return s;
  };
  auto note_class = [&](int kid, int64_t tasks, int64_t flops, int64_t nnz) {
    stats.class_tasks[kid] += tasks;
    stats.class_flops[kid] += flops;
    stats.class_nnz[kid] += nnz;
  };
  if (sc.count[SYM_BM_L] > 0) {
    CB_KBEGIN(CBGPU_K_SYM_BITMAP);
    if (opt.bitmap_cta_threads == 256) {
      auto kern = sym_bitmap_kernel<SR, MERGE, 256>;
      CB_TRY(optin_smem(ctx, kern, sym_bytes));
      kern<<<(unsigned)sc.count[SYM_BM_L], 256, sym_bytes, st>>>(src, recs + sc.begin[SYM_BM_L], sc.count[SYM_BM_L], io.m, tasknnz,
                                                               saved, max_words, save_count, slot_of_task);
    } else {
      auto kern = sym_bitmap_kernel<SR, MERGE, 512>;
      CB_TRY(optin_smem(ctx, kern, sym_bytes));
      kern<<<(unsigned)sc.count[SYM_BM_L], 512, sym_bytes, st>>>(src, recs + sc.begin[SYM_BM_L], sc.count[SYM_BM_L], io.m, tasknnz,
                                                               saved, max_words, save_count, slot_of_task);
    }
    CB_LAUNCH_CHECK(ctx);
    CB_KEND(CBGPU_K_SYM_BITMAP);
    note_class(CBGPU_K_SYM_BITMAP, sc.count[SYM_BM_L], class_weight(bins, sym_class, SYM_BM_L, false), 0);
  }
  if (sc.count[SYM_BM_S] > 0) {
    CB_KBEGIN(CBGPU_K_SYM_BITMAP_S);
    if (opt.bitmap_small_threads == 256) {
      auto kern = sym_bitmap_kernel<SR, MERGE, 256>;
      CB_TRY(optin_smem(ctx, kern, sym_bytes));
      kern<<<(unsigned)sc.count[SYM_BM_S], 256, sym_bytes, st>>>(src, recs + sc.begin[SYM_BM_S], sc.count[SYM_BM_S], io.m, tasknnz,
                                                               saved, max_words, save_count, slot_of_task);
    } else {
      auto kern = sym_bitmap_kernel<SR, MERGE, 128>;
      CB_TRY(optin_smem(ctx, kern, sym_bytes));
      kern<<<(unsigned)sc.count[SYM_BM_S], 128, sym_bytes, st>>>(src, recs + sc.begin[SYM_BM_S], sc.count[SYM_BM_S], io.m, tasknnz,
                                                               saved, max_words, save_count, slot_of_task);
    }
    CB_LAUNCH_CHECK(ctx);
    CB_KEND(CBGPU_K_SYM_BITMAP_S);
    note_class(CBGPU_K_SYM_BITMAP_S, sc.count[SYM_BM_S], class_weight(bins, sym_class, SYM_BM_S, false), 0);
  }
  stats.flops_sym[0] = class_weight(bins, sym_class, SYM_BM_L, false) + class_weight(bins, sym_class, SYM_BM_S, false);
  if (sc.count[SYM_H_CTA] > 0) {
    auto kern = sym_hash_kernel<SR, MERGE, 8, 12>;
    size_t sm = sizeof(unsigned) << 12;
    CB_KBEGIN(CBGPU_K_SYM_HASH_CTA);
    kern<<<(unsigned)sc.count[SYM_H_CTA], 256, sm, st>>>(src, order + sc.begin[SYM_H_CTA], sc.count[SYM_H_CTA], tasknnz);
    CB_LAUNCH_CHECK(ctx);
    CB_KEND(CBGPU_K_SYM_HASH_CTA);
    stats.flops_sym[2] = class_weight(bins, sym_class, SYM_H_CTA, false);
  }
  if (sc.count[SYM_H_WARP] > 0) {
    auto kern = sym_hash_kernel<SR, MERGE, 1, 9>;
    size_t sm = 8 * (sizeof(unsigned) << 9);
    CB_KBEGIN(CBGPU_K_SYM_HASH_WARP);
    kern<<<(unsigned)((sc.count[SYM_H_WARP] + 7) / 8), 256, sm, st>>>(src, order + sc.begin[SYM_H_WARP], sc.count[SYM_H_WARP], tasknnz);
    CB_LAUNCH_CHECK(ctx);
    CB_KEND(CBGPU_K_SYM_HASH_WARP);
    stats.flops_sym[3] = class_weight(bins, sym_class, SYM_H_WARP, false);
  }
  if (sc.count[SYM_H_WARP_S] > 0) {
    auto kern = sym_hash_kernel<SR, MERGE, 1, 6>;
    size_t sm = 8 * (sizeof(unsigned) << 6);
    CB_KBEGIN(CBGPU_K_SYM_HASH_WARP_S);
    kern<<<(unsigned)((sc.count[SYM_H_WARP_S] + 7) / 8), 256, sm, st>>>(src, order + sc.begin[SYM_H_WARP_S], sc.count[SYM_H_WARP_S], tasknnz);
    CB_LAUNCH_CHECK(ctx);
    CB_KEND(CBGPU_K_SYM_HASH_WARP_S);
    stats.flops_sym[4] = class_weight(bins, sym_class, SYM_H_WARP_S, false);
  }
  // register sort: rows only, count the distinct ones
  {
    const size_t sm = 2048 * sizeof(unsigned);
    auto launch = [&](auto kern, int cls, int G) -> int {
      if (sc.count[cls] <= 0) return CBGPU_OK;
      const int64_t groups = 256 / G;
      kern<<<(unsigned)((sc.count[cls] + groups - 1) / groups), 256, sm, st>>>(src, recs + sc.begin[cls], sc.count[cls], tasknnz, nullptr,
                                                                             nullptr);
      CB_LAUNCH_CHECK(ctx);
      return CBGPU_OK;
    };
    if (sc.count[SYM_RS_8] + sc.count[SYM_RS_16] + sc.count[SYM_RS_32] > 0) {
      CB_KBEGIN(CBGPU_K_SYM_REGSORT);
      CB_TRY(launch(regsort_kernel<SR, MERGE, 8, false>, SYM_RS_8, 8));
      CB_TRY(launch(regsort_kernel<SR, MERGE, 16, false>, SYM_RS_16, 16));
      CB_TRY(launch(regsort_kernel<SR, MERGE, 32, false>, SYM_RS_32, 32));
      CB_KEND(CBGPU_K_SYM_REGSORT);
      for (int c = SYM_RS_8; c <= SYM_RS_32; ++c) note_class(CBGPU_K_SYM_REGSORT, sc.count[c], class_weight(bins, sym_class, c, false), 0);
    }
  }
  // ---- K3: scan -> output offsets
  CB_TRY(exclusive_scan_i64(ctx, tasknnz, taskptr, ntask));
  int64_t nnzC = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&nnzC, taskptr + ntask, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  CB_CUDA(ctx, cudaStreamSynchronize(st));
  stats.nnz_out = nnzC;
  if (io.flops_out) *io.flops_out = stats.flops;
  if (io.nnz_out) *io.nnz_out = nnzC;
  if (add_forbidden<SR>::value && io.C && nnzC != stats.flops) // BoolCopy1st/2ndSRing::add throws (Semirings.h:56-62)
    return set_error(ctx, CBGPU_ERR_INVALID, "Add should not happen (BoolCopy semiring): %lld products for %lld outputs",
                     (long long)stats.flops, (long long)nnzC);

  int rc = CBGPU_OK;
  cbgpu_mat_impl *Cm = nullptr;
  MatGuard cguard(ctx, &Cm);
  // per-column results of the symbolic pass for callers that want what estimateFLOP / estimateNNZ_Hash return
  if (io.col_flops_host && ncol > 0)
    CB_CUDA(ctx, cudaMemcpyAsync(io.col_flops_host, colflop, sizeof(int64_t) * (size_t)ncol, cudaMemcpyDeviceToHost, st));
  if (io.col_nnz_host && ncol > 0) {
    std::vector<int64_t> cp((size_t)ncol + 1);
    const int64_t *src_ptr = taskptr;
    int64_t *colptr_sym = nullptr;
    if (first) {
      CB_TRY(scratch.alloc(&colptr_sym, (size_t)ncol + 1));
      gather_ptr_kernel<<<(unsigned)((ncol + 1 + 255) / 256), 256, 0, st>>>(taskptr, first, ncol, colptr_sym);
      CB_LAUNCH_CHECK(ctx);
      src_ptr = colptr_sym;
    }
    CB_CUDA(ctx, cudaMemcpyAsync(cp.data(), src_ptr, sizeof(int64_t) * ((size_t)ncol + 1), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    for (int64_t j = 0; j < ncol; ++j) io.col_nnz_host[j] = cp[(size_t)j + 1] - cp[(size_t)j];
    scratch.release(colptr_sym);
  }
  if (io.col_flops_host) CB_CUDA(ctx, cudaStreamSynchronize(st));
  if (io.C) {
    // ---- output block
    CB_TRY(mat_alloc(ctx, io.m, io.n_out, nnzC, -1, io.out_dtype, &Cm));
    // ---- numeric ordering by exact output size
    BinResult nb;
    ClassRanges nc;
    memset(&nb, 0, sizeof(nb));
    memset(&nc, 0, sizeof(nc));
    uint8_t num_class[256];
    for (int b = 0; b < 256; ++b) {
      uint8_t c = 15;
      if (b >= 1 && b <= 6) c = NUM_H_WARP_S;
      else if (b == 7) c = NUM_H_WARP_M;
      else if (b == 8) c = NUM_H_WARP_M2;
      else if (b == 9) c = NUM_H_WARP;
      else if (b >= 10 && b <= 39) c = NUM_H_CTA;
      else if (b >= 81 && b <= 119) c = NUM_BM_G;
      else if (b >= 121 && b <= 159) c = NUM_SA_S;
      else if (b >= 161 && b <= 199) c = NUM_SA_M;
      else if (b >= 201 && b <= 239) c = NUM_SA_L;
      else if (b >= 240 && b <= 242) c = (uint8_t)(NUM_RS_8 + (b - 240));
      num_class[b] = c;
    }
    // capacity (outputs per task) of the three shared-accumulator shapes
    const int64_t dyn_s = sacc_dynamic_smem(ctx, kSaccThreadsS, kSaccBlocksS, sizeof(FlatQueueT<kSaccThreadsS>));
    const int64_t dyn_m = sacc_dynamic_smem(ctx, kSaccThreadsM, kSaccBlocksM, sizeof(FlatQueueT<kSaccThreadsM>));
    const int64_t dyn_l = sacc_dynamic_smem(ctx, kSaccThreadsL, kSaccBlocksL, sizeof(FlatQueueT<kSaccThreadsL>));
    // second version of the kernel (option sacc_v2, bit per shape): 16-bit ranks, and for the small and medium shape (bit 3: the
    // large one too, bit 4: not the medium one) a row array beside the accumulators; needs max_words <= 16 * THREADS (vector scan)
    const bool v2_ok = max_words <= 4096; // 16-bit row offsets + the window's half bit: windows of at most 2^17 rows
    const bool v2_s = (opt.sacc_v2 & 1) && v2_ok && max_words <= 16 * kSaccThreadsS, v2_m = (opt.sacc_v2 & 2) && v2_ok,
               v2_l = (opt.sacc_v2 & 4) && v2_ok;
    const bool rbw_s = true, rbw_m = !(opt.sacc_v2 & 16), rbw_l = (opt.sacc_v2 & 8) != 0;
    auto cap_of = [&](int64_t dyn, bool v2, bool rbw) -> int64_t {
      int64_t c = (dyn - (int64_t)bm_bytes) / (int64_t)sizeof(acc_t);
      if (v2) // 256 bytes of the budget go to the static workspace the second version adds (64-bit warp totals, the mbarrier)
        c = std::min<int64_t>(((dyn - 256 - (int64_t)max_words * 6) / (int64_t)(sizeof(acc_t) + (rbw ? 2 : 0))) & ~(int64_t)3, 65532);
      return (opt.shared_acc && c >= 64) ? c : 0;
    };
    const int64_t lay_s = cap_of(dyn_s, v2_s, rbw_s), lay_m = cap_of(dyn_m, v2_m, rbw_m), lay_l = cap_of(dyn_l, v2_l, rbw_l);
    int64_t cap_s = lay_s, cap_m = lay_m, cap_l = lay_l;
    int64_t pass_l = lay_l; // outputs per pass of the large shape (the accumulator array of the second version)
    if (opt.shared_acc_max > 0) { // tests and tuning: shrink the three capacities together so that small inputs reach every shape
      cap_l = std::min(cap_l, opt.shared_acc_max), cap_m = std::min(cap_m, opt.shared_acc_max / 2), cap_s = std::min(cap_s, opt.shared_acc_max / 4);
      pass_l = std::max<int64_t>(std::min(lay_l, opt.shared_acc_max) & ~(int64_t)3, 4);
    }
    // the large shape of the second version also takes tasks above its capacity (16-bit ranks: up to 65532 outputs, or
    // sacc_overflow times the capacity): what does not fit the accumulators goes to C with L2 reductions
    if (v2_l && lay_l > 0 && opt.sacc_overflow > 1) cap_l = std::min<int64_t>(pass_l * opt.sacc_overflow, 65532);
    if (opt.shared_acc_small_max >= 0) cap_s = std::min(cap_s, opt.shared_acc_small_max);
    if (ntask > 0) {
      // below this many outputs a task is cheaper in the per-warp hash classes than with a window-sized bitmap
      int64_t auto_min = std::min<int64_t>(std::max<int64_t>(std::min<int64_t>(io.m, W) / 512, 32), 256);
      int64_t bmin = opt.bitmap_min_nnz > 0 ? opt.bitmap_min_nnz : auto_min;
      num_bucket_kernel<<<(unsigned)((ntask + 255) / 256), 256, 0, st>>>(tasknnz, task_win, nwin, ntask, bmin, 2048, cap_s, cap_m,
                                                                        cap_l, (int)opt.force_path, tiny, taskflop, bucket);
      CB_LAUNCH_CHECK(ctx);
      CB_TRY(bin_tasks(ctx, bucket, taskflop, tasknnz, ntask, task_win, num_class, order, &nb, &nc));
      if (nb.listed > 0) {
        task_record_kernel<SR, MERGE><<<(unsigned)((nb.listed + 255) / 256), 256, 0, st>>>(src, order, nb.listed, taskptr, slot_of_task, recs);
        CB_LAUNCH_CHECK(ctx);
      }
    }
    out_t *Cval = reinterpret_cast<out_t *>(Cm->numx);
    // bitmap, accumulators in C itself: RED.ADD straight into the (L2-resident) output slice of the task.
    // Largest tasks by 512 threads; the many small ones (<= 2048 outputs) by 128 threads with a small shared-memory
    // footprint so that many of them are in flight per SM (their cost is dependent-load latency, not bandwidth).
    if (nc.count[NUM_BM_G] > 0) {
      size_t sm = bm_bytes + 16;
      CB_KBEGIN(CBGPU_K_NUM_BITMAP_GMEM);
      if (opt.bitmap_cta_threads == 256) {
        auto kern = num_bitmap_kernel<SR, MERGE, 256>;
        CB_TRY(optin_smem(ctx, kern, sm));
        kern<<<(unsigned)nc.count[NUM_BM_G], 256, sm, st>>>(src, recs + nc.begin[NUM_BM_G], nc.count[NUM_BM_G], io.m, max_words,
                                                           Cm->ir, Cval, saved, max_words);
      } else {
        auto kern = num_bitmap_kernel<SR, MERGE, 512>;
        CB_TRY(optin_smem(ctx, kern, sm));
        kern<<<(unsigned)nc.count[NUM_BM_G], 512, sm, st>>>(src, recs + nc.begin[NUM_BM_G], nc.count[NUM_BM_G], io.m, max_words,
                                                           Cm->ir, Cval, saved, max_words);
      }
      CB_LAUNCH_CHECK(ctx);
      CB_KEND(CBGPU_K_NUM_BITMAP_GMEM);
      stats.tasks_bitmap_gmem = nc.count[NUM_BM_G];
      stats.flops_bitmap_gmem = class_weight(nb, num_class, NUM_BM_G, false);
      stats.nnz_bitmap_gmem = class_weight(nb, num_class, NUM_BM_G, true);
    }
    // bitmap, accumulators in shared memory (exchange protocol): three CTA shapes by output count
    // second version: same task records and shared-memory budget, its own layout (capacity lay_*)
    auto launch_v2 = [&](auto kern, int cls, int threads, int64_t dyn, int64_t lay) -> int {
      CB_TRY(optin_smem(ctx, kern, (size_t)dyn - 256));
      kern<<<(unsigned)nc.count[cls], threads, (size_t)dyn - 256, st>>>(src, recs + nc.begin[cls], io.m, max_words, (int)lay, Cm->ir, Cval,
                                                                       saved, max_words);
      CB_LAUNCH_CHECK(ctx);
      return CBGPU_OK;
    };
    if (nc.count[NUM_SA_L] > 0) {
      CB_KBEGIN(CBGPU_K_NUM_BITMAP_SMEM);
      if (v2_l && rbw_l) {
        CB_TRY(launch_v2(num_sacc2_kernel<SR, MERGE, kSaccThreadsL, kSaccBlocksL, false, true>, NUM_SA_L, kSaccThreadsL, dyn_l, pass_l));
      } else if (v2_l) {
        CB_TRY(launch_v2(num_sacc2_kernel<SR, MERGE, kSaccThreadsL, kSaccBlocksL, false, false>, NUM_SA_L, kSaccThreadsL, dyn_l, pass_l));
      } else {
        auto kern = num_sacc_kernel<SR, MERGE, kSaccThreadsL, kSaccBlocksL, false>;
        CB_TRY(optin_smem(ctx, kern, (size_t)dyn_l));
        kern<<<(unsigned)nc.count[NUM_SA_L], kSaccThreadsL, (size_t)dyn_l, st>>>(src, recs + nc.begin[NUM_SA_L], io.m, max_words, Cm->ir,
                                                                                Cval, saved, max_words);
        CB_LAUNCH_CHECK(ctx);
      }
      CB_KEND(CBGPU_K_NUM_BITMAP_SMEM);
    }
    if (nc.count[NUM_SA_M] > 0) {
      CB_KBEGIN(CBGPU_K_NUM_SACC_M);
      if (v2_m && rbw_m) {
        CB_TRY(launch_v2(num_sacc2_kernel<SR, MERGE, kSaccThreadsM, kSaccBlocksM, true, true>, NUM_SA_M, kSaccThreadsM, dyn_m, lay_m));
      } else if (v2_m) {
        CB_TRY(launch_v2(num_sacc2_kernel<SR, MERGE, kSaccThreadsM, kSaccBlocksM, true, false>, NUM_SA_M, kSaccThreadsM, dyn_m, lay_m));
      } else {
        auto kern = num_sacc_kernel<SR, MERGE, kSaccThreadsM, kSaccBlocksM, true>; // compression 1.6: most slots see one product
        CB_TRY(optin_smem(ctx, kern, (size_t)dyn_m));
        kern<<<(unsigned)nc.count[NUM_SA_M], kSaccThreadsM, (size_t)dyn_m, st>>>(src, recs + nc.begin[NUM_SA_M], io.m, max_words, Cm->ir,
                                                                                Cval, saved, max_words);
        CB_LAUNCH_CHECK(ctx);
      }
      CB_KEND(CBGPU_K_NUM_SACC_M);
    }
    if (nc.count[NUM_SA_S] > 0) {
      CB_KBEGIN(CBGPU_K_NUM_SACC_S);
      if (v2_s) {
        CB_TRY(launch_v2(num_sacc2_kernel<SR, MERGE, kSaccThreadsS, kSaccBlocksS, true, rbw_s>, NUM_SA_S, kSaccThreadsS, dyn_s, lay_s));
      } else {
        auto kern = num_sacc_kernel<SR, MERGE, kSaccThreadsS, kSaccBlocksS, true>;
        CB_TRY(optin_smem(ctx, kern, (size_t)dyn_s));
        kern<<<(unsigned)nc.count[NUM_SA_S], kSaccThreadsS, (size_t)dyn_s, st>>>(src, recs + nc.begin[NUM_SA_S], io.m, max_words, Cm->ir,
                                                                                Cval, saved, max_words);
        CB_LAUNCH_CHECK(ctx);
      }
      CB_KEND(CBGPU_K_NUM_SACC_S);
    }
    {
      const int kid[3] = {CBGPU_K_NUM_BITMAP_SMEM, CBGPU_K_NUM_SACC_M, CBGPU_K_NUM_SACC_S};
      const int cls[3] = {NUM_SA_L, NUM_SA_M, NUM_SA_S};
      for (int i = 0; i < 3; ++i) {
        const int c = cls[i];
        stats.tasks_bitmap_smem += nc.count[c];
        stats.flops_bitmap_smem += class_weight(nb, num_class, c, false);
        stats.nnz_bitmap_smem += class_weight(nb, num_class, c, true);
        note_class(kid[i], nc.count[c], class_weight(nb, num_class, c, false), class_weight(nb, num_class, c, true));
      }
      note_class(CBGPU_K_NUM_BITMAP_GMEM, nc.count[NUM_BM_G], class_weight(nb, num_class, NUM_BM_G, false),
                 class_weight(nb, num_class, NUM_BM_G, true));
      note_class(CBGPU_K_NUM_HASH_CTA, nc.count[NUM_H_CTA], class_weight(nb, num_class, NUM_H_CTA, false),
                 class_weight(nb, num_class, NUM_H_CTA, true));
      note_class(CBGPU_K_NUM_HASH_WARP, nc.count[NUM_H_WARP], class_weight(nb, num_class, NUM_H_WARP, false),
                 class_weight(nb, num_class, NUM_H_WARP, true));
      note_class(CBGPU_K_NUM_HASH_WARP_M, nc.count[NUM_H_WARP_M] + nc.count[NUM_H_WARP_M2],
                 class_weight(nb, num_class, NUM_H_WARP_M, false) + class_weight(nb, num_class, NUM_H_WARP_M2, false),
                 class_weight(nb, num_class, NUM_H_WARP_M, true) + class_weight(nb, num_class, NUM_H_WARP_M2, true));
      note_class(CBGPU_K_NUM_HASH_WARP_S, nc.count[NUM_H_WARP_S], class_weight(nb, num_class, NUM_H_WARP_S, false),
                 class_weight(nb, num_class, NUM_H_WARP_S, true));
    }
    // hash per CTA: 257..2048 outputs
    if (nc.count[NUM_H_CTA] > 0) {
      auto kern = num_hash_kernel<SR, MERGE, 8, 12>;
      size_t sm = ((size_t)1 << 12) * (4 + sizeof(acc_t) + 4) + 16;
      CB_TRY(optin_smem(ctx, kern, sm));
      CB_KBEGIN(CBGPU_K_NUM_HASH_CTA);
      kern<<<(unsigned)nc.count[NUM_H_CTA], 256, sm, st>>>(src, order + nc.begin[NUM_H_CTA], nc.count[NUM_H_CTA], taskptr, Cm->ir, Cval);
      CB_LAUNCH_CHECK(ctx);
      CB_KEND(CBGPU_K_NUM_HASH_CTA);
      stats.tasks_hash_cta = nc.count[NUM_H_CTA];
      stats.flops_hash_cta = class_weight(nb, num_class, NUM_H_CTA, false);
      stats.nnz_hash_cta = class_weight(nb, num_class, NUM_H_CTA, true);
    }
    // hash per warp: 33..256 outputs
    if (nc.count[NUM_H_WARP] > 0) {
      auto kern = num_hash_kernel<SR, MERGE, 1, 9>;
      size_t sm = 8 * ((size_t)1 << 9) * (4 + sizeof(acc_t) + 4) + 64;
      CB_TRY(optin_smem(ctx, kern, sm));
      CB_KBEGIN(CBGPU_K_NUM_HASH_WARP);
      kern<<<(unsigned)((nc.count[NUM_H_WARP] + 7) / 8), 256, sm, st>>>(src, order + nc.begin[NUM_H_WARP], nc.count[NUM_H_WARP], taskptr,
                                                                       Cm->ir, Cval);
      CB_LAUNCH_CHECK(ctx);
      CB_KEND(CBGPU_K_NUM_HASH_WARP);
      stats.tasks_hash_warp += nc.count[NUM_H_WARP];
      stats.flops_hash_warp += class_weight(nb, num_class, NUM_H_WARP, false);
      stats.nnz_hash_warp += class_weight(nb, num_class, NUM_H_WARP, true);
    }
    // hash per warp: 33..64 outputs (128 slots: the Erdos-Renyi regime, d^2 = 64 products per column)
    if (nc.count[NUM_H_WARP_M] + nc.count[NUM_H_WARP_M2] > 0) {
      auto kern = num_hash_kernel<SR, MERGE, 1, 7>;
      size_t sm = 8 * ((size_t)1 << 7) * (4 + sizeof(acc_t) + 4) + 64;
      CB_TRY(optin_smem(ctx, kern, sm));
      CB_KBEGIN(CBGPU_K_NUM_HASH_WARP_M);
      if (nc.count[NUM_H_WARP_M] > 0) {
        kern<<<(unsigned)((nc.count[NUM_H_WARP_M] + 7) / 8), 256, sm, st>>>(src, order + nc.begin[NUM_H_WARP_M], nc.count[NUM_H_WARP_M],
                                                                           taskptr, Cm->ir, Cval);
        CB_LAUNCH_CHECK(ctx);
      }
      if (nc.count[NUM_H_WARP_M2] > 0) { // 65..128 outputs: 256 slots
        auto kern2 = num_hash_kernel<SR, MERGE, 1, 8>;
        size_t sm2 = 8 * ((size_t)1 << 8) * (4 + sizeof(acc_t) + 4) + 64;
        CB_TRY(optin_smem(ctx, kern2, sm2));
        kern2<<<(unsigned)((nc.count[NUM_H_WARP_M2] + 7) / 8), 256, sm2, st>>>(src, order + nc.begin[NUM_H_WARP_M2], nc.count[NUM_H_WARP_M2],
                                                                              taskptr, Cm->ir, Cval);
        CB_LAUNCH_CHECK(ctx);
        stats.tasks_hash_warp += nc.count[NUM_H_WARP_M2];
        stats.flops_hash_warp += class_weight(nb, num_class, NUM_H_WARP_M2, false);
        stats.nnz_hash_warp += class_weight(nb, num_class, NUM_H_WARP_M2, true);
      }
      CB_KEND(CBGPU_K_NUM_HASH_WARP_M);
      stats.tasks_hash_warp += nc.count[NUM_H_WARP_M];
      stats.flops_hash_warp += class_weight(nb, num_class, NUM_H_WARP_M, false);
      stats.nnz_hash_warp += class_weight(nb, num_class, NUM_H_WARP_M, true);
    }
    // hash per warp: <= 32 outputs
    if (nc.count[NUM_H_WARP_S] > 0) {
      auto kern = num_hash_kernel<SR, MERGE, 1, 6>;
      size_t sm = 8 * ((size_t)1 << 6) * (4 + sizeof(acc_t) + 4) + 64;
      CB_TRY(optin_smem(ctx, kern, sm));
      CB_KBEGIN(CBGPU_K_NUM_HASH_WARP_S);
      kern<<<(unsigned)((nc.count[NUM_H_WARP_S] + 7) / 8), 256, sm, st>>>(src, order + nc.begin[NUM_H_WARP_S], nc.count[NUM_H_WARP_S],
                                                                         taskptr, Cm->ir, Cval);
      CB_LAUNCH_CHECK(ctx);
      CB_KEND(CBGPU_K_NUM_HASH_WARP_S);
      stats.tasks_hash_warp += nc.count[NUM_H_WARP_S];
      stats.flops_hash_warp += class_weight(nb, num_class, NUM_H_WARP_S, false);
      stats.nnz_hash_warp += class_weight(nb, num_class, NUM_H_WARP_S, true);
    }
    // register sort: products sorted by row in registers, equal rows folded, no table
    {
      const size_t sm = 2048 * (sizeof(unsigned) + sizeof(acc_t));
      auto launch = [&](auto kern, int cls, int G) -> int {
        if (nc.count[cls] <= 0) return CBGPU_OK;
        const int64_t groups = 256 / G;
        kern<<<(unsigned)((nc.count[cls] + groups - 1) / groups), 256, sm, st>>>(src, recs + nc.begin[cls], nc.count[cls], nullptr, Cm->ir,
                                                                               Cval);
        CB_LAUNCH_CHECK(ctx);
        return CBGPU_OK;
      };
      if (nc.count[NUM_RS_8] + nc.count[NUM_RS_16] + nc.count[NUM_RS_32] > 0) {
        CB_KBEGIN(CBGPU_K_NUM_REGSORT);
        // row << log2(capacity) | position must fit 32 bits below the pad key: blocks of fewer than 2^24 rows sort packed keys
        if (io.m < ((int64_t)1 << 24) - 1 && opt.regsort_packed) {
          CB_TRY(launch(regsort_kernel<SR, MERGE, 8, true, true>, NUM_RS_8, 8));
          CB_TRY(launch(regsort_kernel<SR, MERGE, 16, true, true>, NUM_RS_16, 16));
          CB_TRY(launch(regsort_kernel<SR, MERGE, 32, true, true>, NUM_RS_32, 32));
        } else {
          CB_TRY(launch(regsort_kernel<SR, MERGE, 8, true>, NUM_RS_8, 8));
          CB_TRY(launch(regsort_kernel<SR, MERGE, 16, true>, NUM_RS_16, 16));
          CB_TRY(launch(regsort_kernel<SR, MERGE, 32, true>, NUM_RS_32, 32));
        }
        CB_KEND(CBGPU_K_NUM_REGSORT);
        for (int c = NUM_RS_8; c <= NUM_RS_32; ++c) {
          note_class(CBGPU_K_NUM_REGSORT, nc.count[c], class_weight(nb, num_class, c, false), class_weight(nb, num_class, c, true));
          stats.tasks_hash_warp += nc.count[c]; // census: the per-warp classes they replace
          stats.flops_hash_warp += class_weight(nb, num_class, c, false);
          stats.nnz_hash_warp += class_weight(nb, num_class, c, true);
        }
      }
    }
    // ---- DCSC assembly: column pointers of the candidate columns, then drop the empty ones
    const int64_t *cand_ptr = taskptr;
    int64_t *colptr_out = nullptr;
    if (first) {
      CB_TRY(scratch.alloc(&colptr_out, (size_t)ncol + 1));
      gather_ptr_kernel<<<(unsigned)((ncol + 1 + 255) / 256), 256, 0, st>>>(taskptr, first, ncol, colptr_out);
      CB_LAUNCH_CHECK(ctx);
      cand_ptr = colptr_out;
    }
    rc = compact_columns(ctx, io.out_col_ids, cand_ptr, ncol, &Cm->jc, &Cm->cp, &Cm->nzc);
      stats.nzc_out = Cm->nzc;
  }
  CB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));

  if (rc != CBGPU_OK) return rc; // the guards release C and the temporaries
  CB_CUDA(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&stats.ms_setup, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&stats.ms_symbolic, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&stats.ms_numeric, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&stats.ms_total, ctx->ev[0], ctx->ev[3]);
  for (int i = 0; i < CBGPU_K_COUNT; ++i)
    if (used[i]) cudaEventElapsedTime(&stats.ms_kernel[i], ctx->kev[2 * i], ctx->kev[2 * i + 1]);
  stats.kernel_launches = ctx->launches - launches0;
  if (io.stats) *io.stats = stats;
  if (io.C) *io.C = Cm;
  cguard.armed = false;
  return CBGPU_OK;
}

} // namespace cbgpu
