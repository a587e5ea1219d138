// Streaming 2-way merge of two DCSC blocks of equal shape: C = X (+) Y with SR::add on equal (row, col).
// Replaces the k = 2 case of MultiwayMerge / MultiwayMergeHash (MultiwayMerge.h:428-543, :553-701) -- the case every
// 2x2 SUMMA (two stages) and every 2-layer fiber reduction produces. Both inputs have ascending rows per column, so
// the merge is a pure stream: no hashing, no atomics, no sort. Work is cut into tiles of kTile merged elements of one
// column (merge-path diagonals found by binary search); a count pass (row ids only) gives the per-tile output sizes,
// one exclusive scan turns them into output offsets and column pointers, the write pass merges again and emits.
// Duplicate pairs are never split across tile or thread boundaries (the partition takes the Y twin along with its X).
#pragma once
#include "common.cuh"
#include "semiring.cuh"
#include "util.cuh"

namespace cbgpu {

constexpr int kMergeThreads = 128;
constexpr int kMergeItems = 16;
constexpr int kTile = kMergeThreads * kMergeItems; // 2048 merged elements per tile

// merge-path partition of diagonal d over sorted, duplicate-free X[0,nx) and Y[0,ny) with ties taking X first; when the
// last X taken equals the next Y, that Y is taken as well so that a duplicate pair always stays on one side.
template <class RowsX, class RowsY>
__device__ __forceinline__ void merge_partition(const RowsX &X, int nx, const RowsY &Y, int ny, int d, int &i, int &j) {
  int lo = max(0, d - ny), hi = min(d, nx);
  while (lo < hi) { // smallest i such that X[i] > Y[d-i-1] fails ... standard: find i with X[i-1] <= Y[d-i] and Y[d-i-1] < X[i]
    int mid = (lo + hi) >> 1;
    // take X[mid] before Y[d-1-mid] iff X[mid] <= Y[d-1-mid]
    if (X[mid] <= Y[d - 1 - mid]) lo = mid + 1;
    else hi = mid;
  }
  i = lo;
  j = d - lo;
  if (i > 0 && j < ny && X[i - 1] == Y[j]) ++j;
}

struct MergeCols {
  const int64_t *xcp, *ycp; // dense column pointers [n+1]
  const int32_t *xir, *yir;
};

static __global__ void merge_tiles_per_col(MergeCols m, int64_t n, int64_t *ntiles) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int64_t tot = (m.xcp[j + 1] - m.xcp[j]) + (m.ycp[j + 1] - m.ycp[j]);
  ntiles[j] = (tot + kTile - 1) / kTile;
}
static __global__ void merge_fill_tile_cols(const int64_t *first, int64_t n, int32_t *tile_col) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  for (int64_t t = first[j]; t < first[j + 1]; ++t) tile_col[t] = (int32_t)j;
}
static __global__ void merge_col_ptr(const int64_t *tile_base, const int64_t *first, int64_t n, int64_t *colptr) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j <= n) colptr[j] = tile_base[first[j]];
}

struct GlobalRows {
  const int32_t *p;
  __device__ __forceinline__ int operator[](int i) const { return p[i]; }
};
struct SharedRows {
  const int *p;
  __device__ __forceinline__ int operator[](int i) const { return p[i]; }
};

// merge-path start of every tile, one thread per tile: the dependent global-memory binary search is paid once, in
// parallel over all tiles, instead of as a serial latency chain at the head of every merge CTA (and it serves both passes)
static __global__ void merge_partition_kernel(MergeCols m, const int32_t *tile_col, const int64_t *first, int64_t ntiles,
                                              int2 *tile_start) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int col = tile_col[t];
  const int tl = (int)(t - first[col]);
  const int64_t xb = m.xcp[col], yb = m.ycp[col];
  const int nx = (int)(m.xcp[col + 1] - xb), ny = (int)(m.ycp[col + 1] - yb);
  int i, j;
  merge_partition(GlobalRows{m.xir + xb}, nx, GlobalRows{m.yir + yb}, ny, min(nx + ny, tl * kTile), i, j);
  tile_start[t] = make_int2(i, j);
}

// WRITE == false: count the distinct rows of every tile. WRITE == true: emit rows and values at tile_base[t].
template <class SR, bool WRITE>
__global__ void __launch_bounds__(kMergeThreads)
merge2_kernel(MergeCols m, const typename SR::out_t *xval, const typename SR::out_t *yval, const int32_t *tile_col,
              const int64_t *first, const int2 *tile_start, int64_t ntiles, int64_t *tile_count, const int64_t *tile_base,
              int32_t *cir, typename SR::out_t *cval) {
  typedef typename SR::out_t out_t;
  __shared__ int srow[kTile + 2]; // cx + cy <= kTile + 1 elements: X part first, then the Y part
  __shared__ out_t sval[WRITE ? kTile + 2 : 1];
  __shared__ int warp_sums[32];
  __shared__ int total_s;
  const int64_t t = blockIdx.x;
  const int col = tile_col[t];
  const int tl = (int)(t - first[col]);
  const int64_t xb = m.xcp[col], yb = m.ycp[col];
  const int nx = (int)(m.xcp[col + 1] - xb), ny = (int)(m.ycp[col + 1] - yb);
  const bool last_tile = (t + 1 == first[col + 1]);
  const int2 p0 = tile_start[t];
  const int2 p1 = last_tile ? make_int2(nx, ny) : tile_start[t + 1];
  (void)tl;
  const int i0 = p0.x, j0 = p0.y, cx = p1.x - i0, cy = p1.y - j0; // cx + cy <= kTile + 1
  int *sx = srow, *sy = srow + cx;
  for (int q = threadIdx.x; q < cx; q += kMergeThreads) {
    sx[q] = m.xir[xb + i0 + q];
    if (WRITE) sval[q] = xval[xb + i0 + q];
  }
  for (int q = threadIdx.x; q < cy; q += kMergeThreads) {
    sy[q] = m.yir[yb + j0 + q];
    if (WRITE) sval[cx + q] = yval[yb + j0 + q];
  }
  __syncthreads();
  // every thread merges its slice of the tile sequentially
  const int tot = cx + cy;
  int a0, b0, a1, b1;
  merge_partition(SharedRows{sx}, cx, SharedRows{sy}, cy, min(tot, (int)threadIdx.x * kMergeItems), a0, b0);
  merge_partition(SharedRows{sx}, cx, SharedRows{sy}, cy, min(tot, ((int)threadIdx.x + 1) * kMergeItems), a1, b1);
  int orow[kMergeItems + 1];
  out_t oval[WRITE ? kMergeItems + 1 : 1];
  int cnt = 0;
  {
    int a = a0, b = b0;
    while (a < a1 || b < b1) {
      bool takex = (b >= b1) || (a < a1 && sx[a] <= sy[b]);
      int r = takex ? sx[a] : sy[b];
      out_t v = out_t();
      if (WRITE) v = takex ? sval[a] : sval[cx + b];
      if (takex) {
        ++a;
        if (b < b1 && sy[b] == r) { // the Y twin of this row: SR::add(current, stored), MultiwayMerge.h:374
          if (WRITE) v = SR::add(sval[cx + b], v);
          ++b;
        }
      } else {
        ++b;
      }
      if (cnt <= kMergeItems) {
        orow[cnt] = r;
        if (WRITE) oval[cnt] = v;
      }
      ++cnt;
    }
  }
  // block-wide exclusive scan of cnt
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads(); // also: everybody is done reading srow/sval
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < kMergeThreads / 32; ++w) {
      int v = warp_sums[w];
      warp_sums[w] = run;
      run += v;
    }
    total_s = run;
  }
  __syncthreads();
  const int off = warp_sums[warp] + incl - cnt;
  if (!WRITE) {
    if (threadIdx.x == 0) tile_count[t] = total_s;
    return;
  }
  // stage the outputs in shared memory (inputs are dead now), then write them out coalesced
  int *orow_s = srow;
  out_t *oval_s = sval;
  for (int q = 0; q < cnt; ++q) {
    orow_s[off + q] = orow[q];
    oval_s[off + q] = oval[q];
  }
  __syncthreads();
  const int64_t base = tile_base[t];
  for (int q = threadIdx.x; q < total_s; q += kMergeThreads) {
    cir[base + q] = orow_s[q];
    cval[base + q] = oval_s[q];
  }
}

template <class SR>
int merge2_run(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *X, cbgpu_mat_impl *Y, cbgpu_mat_impl **out, cbgpu_stats *stats) {
  typedef typename SR::out_t out_t;
  cudaStream_t st = ctx->stream;
  const int64_t n = X->n;
  const int64_t launches0 = ctx->launches;
  Scratch scratch(ctx); // frees the temporaries on every return
  cudaEventRecord(ctx->ev[0], st);
  CB_TRY(ensure_dense_colptr(ctx, X));
  CB_TRY(ensure_dense_colptr(ctx, Y));
  MergeCols m{X->colptr, Y->colptr, X->ir, Y->ir};
  int64_t *ntiles_col = nullptr, *first = nullptr;
  CB_TRY(scratch.alloc(&ntiles_col, (size_t)n + 1));
  CB_TRY(scratch.alloc(&first, (size_t)n + 1));
  merge_tiles_per_col<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, n, ntiles_col);
  CB_LAUNCH_CHECK(ctx);
  CB_TRY(exclusive_scan_i64(ctx, ntiles_col, first, n));
  int64_t ntiles = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&ntiles, first + n, 8, cudaMemcpyDeviceToHost, st));
  CB_CUDA(ctx, cudaStreamSynchronize(st));
  if (ntiles >= ((int64_t)1 << 31)) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "too many merge tiles");
  int32_t *tile_col = nullptr;
  int64_t *tile_count = nullptr, *tile_base = nullptr;
  CB_TRY(scratch.alloc(&tile_col, (size_t)ntiles + 1));
  CB_TRY(scratch.alloc(&tile_count, (size_t)ntiles + 1));
  CB_TRY(scratch.alloc(&tile_base, (size_t)ntiles + 2));
  merge_fill_tile_cols<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(first, n, tile_col);
  CB_LAUNCH_CHECK(ctx);
  int2 *tile_start = nullptr;
  CB_TRY(scratch.alloc(&tile_start, (size_t)ntiles + 1));
  if (ntiles > 0) {
    merge_partition_kernel<<<(unsigned)((ntiles + 255) / 256), 256, 0, st>>>(m, tile_col, first, ntiles, tile_start);
    CB_LAUNCH_CHECK(ctx);
  }
  cudaEventRecord(ctx->ev[1], st);
  if (ntiles > 0) {
    merge2_kernel<SR, false><<<(unsigned)ntiles, kMergeThreads, 0, st>>>(m, nullptr, nullptr, tile_col, first, tile_start, ntiles,
                                                                        tile_count, nullptr, nullptr, nullptr);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(exclusive_scan_i64(ctx, tile_count, tile_base, ntiles));
  int64_t nnz = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&nnz, tile_base + ntiles, 8, cudaMemcpyDeviceToHost, st));
  cudaEventRecord(ctx->ev[2], st);
  CB_CUDA(ctx, cudaStreamSynchronize(st));
  cbgpu_mat_impl *C = nullptr;
  MatGuard cguard(ctx, &C);
  CB_TRY(mat_alloc(ctx, X->m, n, nnz, -1, X->dtype, &C));
  if (ntiles > 0) {
    merge2_kernel<SR, true><<<(unsigned)ntiles, kMergeThreads, 0, st>>>(
        m, reinterpret_cast<const out_t *>(X->numx), reinterpret_cast<const out_t *>(Y->numx), tile_col, first, tile_start, ntiles,
        nullptr, tile_base, C->ir, reinterpret_cast<out_t *>(C->numx));
    CB_LAUNCH_CHECK(ctx);
  }
  int64_t *colptr = nullptr;
  CB_TRY(scratch.alloc(&colptr, (size_t)n + 1));
  merge_col_ptr<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(tile_base, first, n, colptr);
  CB_LAUNCH_CHECK(ctx);
  int rc = compact_columns(ctx, nullptr, colptr, n, &C->jc, &C->cp, &C->nzc);
  scratch.detach(colptr);
  C->colptr = colptr; // the merged block already has its dense column index (a following merge round uses it)
  cudaEventRecord(ctx->ev[3], st);
  if (rc != CBGPU_OK) return rc; // the guards release C and the temporaries
  CB_CUDA(ctx, cudaStreamSynchronize(st));
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->flops = X->nnz + Y->nnz;
    stats->nnz_out = nnz;
    stats->nzc_out = C->nzc;
    stats->tasks = ntiles;
    cudaEventElapsedTime(&stats->ms_setup, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&stats->ms_symbolic, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&stats->ms_numeric, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&stats->ms_total, ctx->ev[0], ctx->ev[3]);
    stats->kernel_launches = ctx->launches - launches0;
  }
  *out = C;
  cguard.armed = false;
  return CBGPU_OK;
}

} // namespace cbgpu
