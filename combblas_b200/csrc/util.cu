// Device-wide helper kernels: scan, task binning, dense column index, row-window table, column compaction.
// Hand-written (no CUB/Thrust) so every launch on the hot path is ours and counted.
#include <stdarg.h>
#include <mutex>
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"
#include "util.cuh"

namespace cbgpu {

int set_error(cbgpu_ctx_impl *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->last_error = buf;
  else fprintf(stderr, "[cbgpu] %s\n", buf);
  return code;
}

constexpr size_t kBigBlock = (size_t)64 << 20; // blocks from 64 MiB are cached per device

// Large-block cache in front of the stream-ordered pool: result arrays of tens of GB are recycled between multiplies
// (and between the column slabs of a phased multiply) without going back to the driver. The cache is per device and
// shared by all contexts (the pipelined 3D driver runs two contexts on two host threads); every cached block carries
// an event recorded on the stream that released it, and the stream that takes it waits on that event.
struct BigCache {
  std::mutex mu;
  std::map<void *, size_t> live;
  struct Entry { void *p; cudaEvent_t ev; };
  std::multimap<size_t, Entry> free_blocks;
};
static BigCache &big_cache(int device) {
  static BigCache caches[64];
  return caches[device & 63];
}

static void flush_big_cache(cbgpu_ctx_impl *ctx) {
  BigCache &bc = big_cache(ctx->device);
  std::lock_guard<std::mutex> lock(bc.mu);
  for (auto &kv : bc.free_blocks) {
    cudaStreamWaitEvent(ctx->stream, kv.second.ev, 0);
    cudaFreeAsync(kv.second.p, ctx->stream);
    cudaEventDestroy(kv.second.ev);
  }
  bc.free_blocks.clear();
}

int dev_alloc(cbgpu_ctx_impl *ctx, void **p, size_t bytes) {
  *p = nullptr;
  // whole 16-byte granules: the bulk (TMA) copies of the streaming merge and of the hand-over read the granule that holds the
  // last element of an array to its end
  bytes = bytes == 0 ? 16 : ((bytes + 15) & ~(size_t)15);
  BigCache &bc = big_cache(ctx->device);
  if (bytes >= kBigBlock) {
    std::lock_guard<std::mutex> lock(bc.mu);
    auto it = bc.free_blocks.lower_bound(bytes); // smallest cached block that fits and wastes at most a quarter
    if (it != bc.free_blocks.end() && it->first <= bytes + bytes / 4) {
      *p = it->second.p;
      cudaStreamWaitEvent(ctx->stream, it->second.ev, 0);
      cudaEventDestroy(it->second.ev);
      bc.live[*p] = it->first;
      bc.free_blocks.erase(it);
      return CBGPU_OK;
    }
  }
  cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
  if (e != cudaSuccess) { // give the cached blocks back and try once more
    cudaGetLastError();
    flush_big_cache(ctx);
    cudaStreamSynchronize(ctx->stream);
    e = cudaMallocAsync(p, bytes, ctx->stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(ctx, e == cudaErrorMemoryAllocation ? CBGPU_ERR_NOMEM : CBGPU_ERR_CUDA,
                     "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  }
  if (bytes >= kBigBlock) {
    std::lock_guard<std::mutex> lock(bc.mu);
    bc.live[*p] = bytes;
  }
  return CBGPU_OK;
}

int dev_free(cbgpu_ctx_impl *ctx, void *p) {
  if (!p) return CBGPU_OK;
  BigCache &bc = big_cache(ctx->device);
  {
    std::lock_guard<std::mutex> lock(bc.mu);
    auto it = bc.live.find(p);
    if (it != bc.live.end()) {
      BigCache::Entry e{p, nullptr};
      cudaEventCreateWithFlags(&e.ev, cudaEventDisableTiming);
      cudaEventRecord(e.ev, ctx->stream); // every consumer of this block was enqueued on ctx->stream before now
      bc.free_blocks.emplace(it->second, e);
      bc.live.erase(it);
      return CBGPU_OK;
    }
  }
  CB_CUDA(ctx, cudaFreeAsync(p, ctx->stream));
  return CBGPU_OK;
}

void release_cached_blocks(cbgpu_ctx_impl *ctx) { flush_big_cache(ctx); }

// bytes of the device's stream-ordered pool that are handed out right now, not counting the blocks parked in the large-block
// cache (they are reusable): what a leak check compares before and after a failed call
int pool_live_bytes(cbgpu_ctx_impl *ctx, int64_t *live) {
  cudaMemPool_t pool;
  CB_CUDA(ctx, cudaDeviceGetDefaultMemPool(&pool, ctx->device));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t used = 0;
  CB_CUDA(ctx, cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
  BigCache &bc = big_cache(ctx->device);
  std::lock_guard<std::mutex> lock(bc.mu);
  uint64_t parked = 0;
  for (auto &kv : bc.free_blocks) parked += kv.first;
  *live = (int64_t)used - (int64_t)parked;
  return CBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ scan
// three-phase exclusive scan of int64: per-tile sums, scan of the sums by one block, per-tile scan + offset
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t block_scan_i64(int64_t v, int64_t *warp_sums, int64_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int64_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int64_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += x;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int64_t w = lane < nwarp ? warp_sums[lane] : 0;
    int64_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int64_t x = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (lane >= d) wi += x;
    }
    warp_sums[lane] = wi - w;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  int64_t r = warp_sums[warp] + incl - v;
  __syncthreads(); // warp_sums is reused by the caller's next round
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int64_t *in, int64_t n, int64_t *tile_sums) {
  __shared__ int64_t ws[32];
  __shared__ int64_t tot;
  int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t s = 0;
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  block_scan_i64(s, ws, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan in place over ntiles entries; writes the grand total to sums[ntiles]
__global__ void __launch_bounds__(1024) scan_sums_inplace(int64_t *sums, int64_t ntiles) {
  __shared__ int64_t ws[32];
  __shared__ int64_t tot;
  int64_t carry = 0;
  for (int64_t base = 0; base < ntiles; base += blockDim.x) {
    int64_t idx = base + threadIdx.x;
    int64_t v = idx < ntiles ? sums[idx] : 0;
    int64_t ex = block_scan_i64(v, ws, &tot);
    if (idx < ntiles) sums[idx] = carry + ex;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[ntiles] = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply(const int64_t *in, int64_t n, const int64_t *tile_offsets, int64_t *out) {
  __shared__ int64_t ws[32];
  __shared__ int64_t tot;
  // each thread owns kScanItems consecutive items so that the scan order equals the index order
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  int64_t ex = block_scan_i64(s, ws, &tot) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_offsets[gridDim.x];
}

__global__ void set_one_i64(int64_t *p, int64_t v) { *p = v; }

int exclusive_scan_i64(cbgpu_ctx_impl *ctx, const int64_t *in, int64_t *out, int64_t n) {
  if (n <= 0) {
    set_one_i64<<<1, 1, 0, ctx->stream>>>(out, 0);
    CB_LAUNCH_CHECK(ctx);
    return CBGPU_OK;
  }
  int64_t ntiles = (n + kScanTile - 1) / kScanTile;
  int64_t *sums = nullptr;
  CB_TRY(dev_alloc_t(ctx, &sums, (size_t)ntiles + 1));
  scan_tile_sums<<<(unsigned)ntiles, kScanThreads, 0, ctx->stream>>>(in, n, sums);
  CB_LAUNCH_CHECK(ctx);
  scan_sums_inplace<<<1, 1024, 0, ctx->stream>>>(sums, ntiles);
  CB_LAUNCH_CHECK(ctx);
  scan_apply<<<(unsigned)ntiles, kScanThreads, 0, ctx->stream>>>(in, n, sums, out);
  CB_LAUNCH_CHECK(ctx);
  CB_TRY(dev_free(ctx, sums));
  return CBGPU_OK;
}

__global__ void fill_i64_kernel(int64_t *p, int64_t n, int64_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_i64(cbgpu_ctx_impl *ctx, int64_t *p, int64_t n, int64_t v) {
  if (n <= 0) return CBGPU_OK;
  fill_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, n, v);
  CB_LAUNCH_CHECK(ctx);
  return CBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ dense column index
__global__ void scatter_col_counts(const int64_t *jc, const int64_t *cp, int64_t nzc, int64_t *counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nzc) counts[jc[i]] = cp[i + 1] - cp[i];
}

static std::mutex &mat_cache_mutex();
static int ensure_dense_colptr_locked(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *M) {
  if (M->colptr) return CBGPU_OK;
  int64_t *counts = nullptr, *colptr = nullptr;
  CB_TRY(dev_alloc_t(ctx, &counts, (size_t)M->n + 1));
  CB_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int64_t) * ((size_t)M->n + 1), ctx->stream));
  if (M->nzc > 0) {
    scatter_col_counts<<<(unsigned)((M->nzc + 255) / 256), 256, 0, ctx->stream>>>(M->jc, M->cp, M->nzc, counts);
    CB_LAUNCH_CHECK(ctx);
  }
  int rc = dev_alloc_t(ctx, &colptr, (size_t)M->n + 1);
  if (rc == CBGPU_OK) rc = exclusive_scan_i64(ctx, counts, colptr, M->n);
  dev_free(ctx, counts);
  if (rc != CBGPU_OK) {
    dev_free(ctx, colptr);
    return rc;
  }
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // visible to every stream once published
  M->colptr = colptr;
  return CBGPU_OK;
}
int ensure_dense_colptr(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *M) {
  if (M->colptr) return CBGPU_OK;
  std::lock_guard<std::mutex> lock(mat_cache_mutex());
  return ensure_dense_colptr_locked(ctx, M);
}

// ------------------------------------------------------------------------------------------------ row-window table
// T[c*nwin + w] = first position in column c whose row >= w << wlog2 ; T[ncols*nwin] = colptr[ncols]
__global__ void window_table_kernel(const int64_t *colptr, const int32_t *rows, int64_t ncols, int nwin, int wlog2,
                                    int64_t *T) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncols) return;
  if (c == ncols) {
    T[ncols * nwin] = colptr[ncols];
    return;
  }
  int64_t lo = colptr[c], end = colptr[c + 1];
  T[c * nwin] = lo;
  for (int w = 1; w < nwin; ++w) {
    int64_t bound = (int64_t)w << wlog2;
    int64_t a = lo, b = end;
    while (a < b) {
      int64_t mid = (a + b) >> 1;
      if ((int64_t)rows[mid] < bound) a = mid + 1;
      else b = mid;
    }
    lo = a;
    T[c * nwin + w] = lo;
  }
}

int build_window_table(cbgpu_ctx_impl *ctx, const int64_t *colptr, const int32_t *rows, int64_t ncols, int nwin,
                       int wlog2, int64_t **T) {
  CB_TRY(dev_alloc_t(ctx, T, (size_t)ncols * nwin + 1));
  window_table_kernel<<<(unsigned)((ncols + 1 + 255) / 256), 256, 0, ctx->stream>>>(colptr, rows, ncols, nwin, wlog2, *T);
  CB_LAUNCH_CHECK(ctx);
  return CBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ window-major copy
// Layout of the A operand for the accumulation engine. Piece (window w, column c) of the copy starts at a multiple of 4
// elements (16 bytes of row ids, 32 bytes of f64 values) and is padded to a multiple of 4 with row id -1, so that a lane
// reads 4 products with one 16-byte load of rows and aligned vector loads of values. T2[w*ncols + c] = start | npad: the
// start is a multiple of 4, the low two bits hold the number of pad entries at the end of the piece, so
//     real length = (T2[i+1] & ~3) - (T2[i] & ~3) - (T2[i] & 3).
// On R-MAT scale 20 the padding adds 28 % to the copy and 0.9 % to the products walked (most products come from long
// pieces); it cuts the load instructions and the segment lookups of the product walk by four.
__global__ void piece_len_kernel(const int64_t *T, int64_t ncols, int nwin, int64_t *len2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // i = w * ncols + c
  if (i >= ncols * nwin) return;
  int64_t w = i / ncols, c = i - w * ncols;
  const int64_t *t = T + c * nwin + w;
  len2[i] = (t[1] - t[0] + 3) & ~(int64_t)3;
}
__global__ void piece_tag_kernel(const int64_t *T, int64_t ncols, int nwin, int64_t *T2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncols * nwin) return;
  int64_t w = i / ncols, c = i - w * ncols;
  const int64_t *t = T + c * nwin + w;
  const int64_t len = t[1] - t[0];
  T2[i] |= ((len + 3) & ~(int64_t)3) - len; // starts are multiples of 4: the pad count fits below them
}
template <int VB>
__global__ void piece_copy_kernel(const int64_t *T, const int64_t *T2, const int32_t *rows, const unsigned char *vals,
                                  int64_t ncols, int nwin, int32_t *wrows, unsigned char *wvals) {
  int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= ncols) return;
  const int lane = threadIdx.x & 31;
  for (int w = 0; w < nwin; ++w) {
    const int64_t src = T[c * nwin + w], n = T[c * nwin + w + 1] - src, tag = T2[(int64_t)w * ncols + c];
    const int64_t dst = tag & ~(int64_t)3, npad = tag & 3;
    for (int64_t i = lane; i < n + npad; i += 32) {
      const bool real = i < n;
      wrows[dst + i] = real ? rows[src + i] : -1;
      if (VB == 8) reinterpret_cast<uint64_t *>(wvals)[dst + i] = real ? reinterpret_cast<const uint64_t *>(vals)[src + i] : 0ull;
      else if (VB == 4) reinterpret_cast<uint32_t *>(wvals)[dst + i] = real ? reinterpret_cast<const uint32_t *>(vals)[src + i] : 0u;
      else wvals[dst + i] = real ? vals[src + i] : (unsigned char)0;
    }
  }
}

int build_window_major(cbgpu_ctx_impl *ctx, const int64_t *colptr, const int32_t *rows, const void *vals, int vbytes,
                       int64_t ncols, int64_t nnz, int nwin, int wlog2, int64_t **T2, int32_t **Wir, void **Wval) {
  (void)nnz;
  int64_t *T = nullptr, *len2 = nullptr;
  CB_TRY(build_window_table(ctx, colptr, rows, ncols, nwin, wlog2, &T));
  const int64_t np = ncols * nwin;
  CB_TRY(dev_alloc_t(ctx, &len2, (size_t)np + 1));
  CB_TRY(dev_alloc_t(ctx, T2, (size_t)np + 1));
  if (np > 0) {
    piece_len_kernel<<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>(T, ncols, nwin, len2);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(exclusive_scan_i64(ctx, len2, *T2, np));
  int64_t padded = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&padded, *T2 + np, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_alloc_t(ctx, Wir, (size_t)padded + 4));
  CB_TRY(dev_alloc(ctx, Wval, ((size_t)padded + 4) * vbytes));
  if (np > 0) {
    piece_tag_kernel<<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>(T, ncols, nwin, *T2);
    CB_LAUNCH_CHECK(ctx);
  }
  if (ncols > 0) {
    unsigned nb = (unsigned)((ncols * 32 + 255) / 256);
    const unsigned char *v = (const unsigned char *)vals;
    unsigned char *wv = (unsigned char *)*Wval;
    if (vbytes == 8) piece_copy_kernel<8><<<nb, 256, 0, ctx->stream>>>(T, *T2, rows, v, ncols, nwin, *Wir, wv);
    else if (vbytes == 4) piece_copy_kernel<4><<<nb, 256, 0, ctx->stream>>>(T, *T2, rows, v, ncols, nwin, *Wir, wv);
    else piece_copy_kernel<1><<<nb, 256, 0, ctx->stream>>>(T, *T2, rows, v, ncols, nwin, *Wir, wv);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(dev_free(ctx, T));
  CB_TRY(dev_free(ctx, len2));
  return CBGPU_OK;
}

// The per-matrix caches (dense column index, window-major copy) are built lazily by whichever multiply first uses the
// matrix as its A operand; several contexts / host threads may multiply with the same A (SlabPipeline, the pipelined 3D
// driver), so building is serialised per process and a cache is never freed while it may be in use: a copy for another
// window size replaces the old one only after the device has drained.
static std::mutex &mat_cache_mutex() {
  static std::mutex m;
  return m;
}

int ensure_window_major(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *M, int nwin, int wlog2) {
  std::lock_guard<std::mutex> lock(mat_cache_mutex());
  if (M->win_T2 && M->win_log2 == wlog2 && M->win_nwin == nwin) return CBGPU_OK;
  if (M->win_T2) {
    CB_CUDA(ctx, cudaDeviceSynchronize()); // kernels of any stream may still read the old copy
    dev_free(ctx, M->win_T2); dev_free(ctx, M->win_ir); dev_free(ctx, M->win_val);
    M->win_T2 = nullptr; M->win_ir = nullptr; M->win_val = nullptr;
  }
  CB_TRY(ensure_dense_colptr_locked(ctx, M));
  CB_TRY(build_window_major(ctx, M->colptr, M->ir, M->numx, (int)dtype_size(M->dtype), M->n, M->nnz, nwin, wlog2, &M->win_T2,
                            &M->win_ir, &M->win_val));
  // other streams may use the copy as soon as this call returns
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  M->win_log2 = wlog2;
  M->win_nwin = nwin;
  return CBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ binning
__global__ void __launch_bounds__(256)
bucket_hist_kernel(const uint8_t *bucket, const int64_t *weight, const int64_t *weight2, int64_t n,
                   unsigned long long *hist, unsigned long long *whist, unsigned long long *whist2) {
  __shared__ unsigned int h[256];
  __shared__ unsigned long long wh[256];
  __shared__ unsigned long long wh2[256];
  h[threadIdx.x] = 0;
  wh[threadIdx.x] = 0;
  wh2[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * 4096;
  for (int i = 0; i < 16; ++i) {
    int64_t idx = base + (int64_t)i * 256 + threadIdx.x;
    if (idx < n) {
      int b = bucket[idx];
      atomicAdd(&h[b], 1u);
      if (weight) atomicAdd(&wh[b], (unsigned long long)weight[idx]);
      if (weight2) atomicAdd(&wh2[b], (unsigned long long)weight2[idx]);
    }
  }
  __syncthreads();
  if (h[threadIdx.x]) {
    atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
    if (weight) atomicAdd(&whist[threadIdx.x], wh[threadIdx.x]);
    if (weight2) atomicAdd(&whist2[threadIdx.x], wh2[threadIdx.x]);
  }
}

struct ClassTable { uint8_t c[256]; };
__global__ void task_key_kernel(const uint8_t *bucket, const uint32_t *task_win, ClassTable tab, int64_t n, unsigned *keys,
                                int32_t *ids) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned b = bucket[i];
  unsigned w = 0;
  if (task_win) {
    unsigned tw = task_win[i];
    if (((tw >> 16) - (tw & 0xFFFFu)) == 1) w = tw & 0xFFFFu;
  }
  keys[i] = ((unsigned)tab.c[b] << 28) | ((w & 0xFFFFu) << 8) | (255u - b);
  ids[i] = (int32_t)i;
}

int bin_tasks(cbgpu_ctx_impl *ctx, const uint8_t *bucket, const int64_t *weight, const int64_t *weight2, int64_t n,
              const uint32_t *task_win, const uint8_t class_of_bucket[256], int32_t *order, BinResult *res,
              ClassRanges *classes) {
  memset(res, 0, sizeof(*res));
  memset(classes, 0, sizeof(*classes));
  if (n <= 0) return CBGPU_OK;
  unsigned long long *dh = nullptr;
  CB_TRY(dev_alloc_t(ctx, &dh, 768));
  CB_CUDA(ctx, cudaMemsetAsync(dh, 0, 768 * sizeof(unsigned long long), ctx->stream));
  unsigned nblk = (unsigned)((n + 4095) / 4096);
  bucket_hist_kernel<<<nblk, 256, 0, ctx->stream>>>(bucket, weight, weight2, n, dh, dh + 256, dh + 512);
  CB_LAUNCH_CHECK(ctx);
  unsigned long long hh[768];
  CB_CUDA(ctx, cudaMemcpyAsync(hh, dh, sizeof(hh), cudaMemcpyDeviceToHost, ctx->stream));
  // ordering: sort task ids by (class, window, descending bucket) -- CUB radix sort (library call on a few MB)
  ClassTable tab;
  memcpy(tab.c, class_of_bucket, 256);
  unsigned *keys = nullptr, *keys2 = nullptr;
  int32_t *ids = nullptr;
  CB_TRY(dev_alloc_t(ctx, &keys, (size_t)n));
  CB_TRY(dev_alloc_t(ctx, &keys2, (size_t)n));
  CB_TRY(dev_alloc_t(ctx, &ids, (size_t)n));
  task_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(bucket, task_win, tab, n, keys, ids);
  CB_LAUNCH_CHECK(ctx);
  size_t tmp_bytes = 0;
  CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, order, n, 0, 32, ctx->stream));
  void *tmp = nullptr;
  CB_TRY(dev_alloc(ctx, &tmp, tmp_bytes));
  CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, order, n, 0, 32, ctx->stream));
  ctx->launches += 4;
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_free(ctx, tmp));
  CB_TRY(dev_free(ctx, keys));
  CB_TRY(dev_free(ctx, keys2));
  CB_TRY(dev_free(ctx, ids));
  CB_TRY(dev_free(ctx, dh));
  for (int b = 0; b < 256; ++b) {
    res->count[b] = (int64_t)hh[b];
    res->weight[b] = (int64_t)hh[256 + b];
    res->weight2[b] = (int64_t)hh[512 + b];
    classes->count[class_of_bucket[b] & 15] += (int64_t)hh[b];
  }
  int64_t off = 0;
  for (int c = 0; c < 16; ++c) {
    classes->begin[c] = off;
    off += classes->count[c];
  }
  res->listed = off - classes->count[15];
  return CBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ column compaction
__global__ void nonempty_flags(const int64_t *ptr, int64_t n, int64_t *flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = ptr[i + 1] > ptr[i] ? 1 : 0;
}
__global__ void compact_cols_kernel(const int64_t *ids, const int64_t *ptr, const int64_t *pos, int64_t n, int64_t *jc,
                                    int64_t *cp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ptr[i + 1] > ptr[i]) {
    int64_t o = pos[i];
    jc[o] = ids ? ids[i] : i;
    cp[o] = ptr[i];
  }
  if (i == n) cp[pos[n]] = ptr[n];
}

int compact_columns(cbgpu_ctx_impl *ctx, const int64_t *cand_ids, const int64_t *cand_ptr, int64_t ncand, int64_t **jc,
                    int64_t **cp, int64_t *nzc) {
  *jc = nullptr;
  *cp = nullptr;
  *nzc = 0;
  if (ncand <= 0) {
    CB_TRY(dev_alloc_t(ctx, jc, 1));
    CB_TRY(dev_alloc_t(ctx, cp, 1));
    CB_CUDA(ctx, cudaMemsetAsync(*cp, 0, sizeof(int64_t), ctx->stream));
    return CBGPU_OK;
  }
  int64_t *flags = nullptr, *pos = nullptr;
  CB_TRY(dev_alloc_t(ctx, &flags, (size_t)ncand));
  CB_TRY(dev_alloc_t(ctx, &pos, (size_t)ncand + 1));
  unsigned nblk = (unsigned)((ncand + 1 + 255) / 256);
  nonempty_flags<<<nblk, 256, 0, ctx->stream>>>(cand_ptr, ncand, flags);
  CB_LAUNCH_CHECK(ctx);
  CB_TRY(exclusive_scan_i64(ctx, flags, pos, ncand));
  int64_t cnt = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&cnt, pos + ncand, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_alloc_t(ctx, jc, (size_t)cnt));
  CB_TRY(dev_alloc_t(ctx, cp, (size_t)cnt + 1));
  compact_cols_kernel<<<nblk, 256, 0, ctx->stream>>>(cand_ids, cand_ptr, pos, ncand, *jc, *cp);
  CB_LAUNCH_CHECK(ctx);
  *nzc = cnt;
  CB_TRY(dev_free(ctx, flags));
  CB_TRY(dev_free(ctx, pos));
  return CBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ matrices
int mat_alloc(cbgpu_ctx_impl *ctx, int64_t m, int64_t n, int64_t nnz, int64_t nzc, int dtype, cbgpu_mat_impl **out) {
  cbgpu_mat_impl *M = new cbgpu_mat_impl();
  M->m = m; M->n = n; M->nnz = nnz; M->nzc = nzc; M->dtype = dtype; M->device = ctx->device;
  int rc = CBGPU_OK;
  if (nzc >= 0) {
    if ((rc = dev_alloc_t(ctx, &M->jc, (size_t)nzc)) != CBGPU_OK) goto fail;
    if ((rc = dev_alloc_t(ctx, &M->cp, (size_t)nzc + 1)) != CBGPU_OK) goto fail;
  }
  if ((rc = dev_alloc_t(ctx, &M->ir, (size_t)nnz)) != CBGPU_OK) goto fail;
  if ((rc = dev_alloc(ctx, &M->numx, (size_t)nnz * dtype_size(dtype))) != CBGPU_OK) goto fail;
  *out = M;
  return CBGPU_OK;
fail:
  mat_release(ctx, M);
  return rc;
}

int mat_release(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *M) {
  if (!M) return CBGPU_OK;
  dev_free(ctx, M->jc);
  dev_free(ctx, M->cp);
  dev_free(ctx, M->ir);
  dev_free(ctx, M->numx);
  dev_free(ctx, M->colptr);
  dev_free(ctx, M->win_T2);
  dev_free(ctx, M->win_ir);
  dev_free(ctx, M->win_val);
  delete M;
  return CBGPU_OK;
}

} // namespace cbgpu
