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
#include "bulk.cuh"

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


// ---- second version: persistent CTAs, tiles brought in by the TMA unit (cp.async.bulk), double buffered -------------
// What ncu showed for the kernel above (profiles/r2_ncu_merge2_before.txt): the count pass issues 77 % of the time (69
// thread instructions per element: two merge-path searches per thread, three shared loads per merged element, scalar
// load/store loops for the tile) and the write pass stalls on its chain of global loads at the head of every CTA. Here
//  * a CTA walks tiles t = blockIdx.x, + gridDim.x, ...; while it merges tile i, the rows (and values) of tile i + 1
//    arrive in the other buffer as two (four) bulk copies issued by one thread and tracked by an mbarrier; the tile
//    descriptor (where its X and Y parts start, how long they are) was prepared once by merge_desc_kernel;
//  * every thread searches ONE diagonal (its end is its neighbour's start) and merges with the heads of both runs in
//    registers: one shared load per consumed element, static register indices for the outputs;
//  * the merged tile leaves as bulk stores from the staging area (16-byte aligned middle; the few elements before and
//    after the aligned part by ordinary stores).
struct MergeTile {
  long long xoff, yoff; // first element of the X / Y part of the tile in the row and value arrays
  int cx, cy;           // their lengths; cx + cy <= kTile + 1
  int pad0, pad1;
};

static __global__ void merge_desc_kernel(MergeCols m, const int32_t *tile_col, const int64_t *first, const int2 *tile_start,
                                         int64_t ntiles, MergeTile *desc) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int col = tile_col[t];
  const int64_t xb = m.xcp[col], yb = m.ycp[col];
  const bool last_tile = (t + 1 == first[col + 1]);
  const int2 p0 = tile_start[t];
  const int2 p1 = last_tile ? make_int2((int)(m.xcp[col + 1] - xb), (int)(m.ycp[col + 1] - yb)) : tile_start[t + 1];
  MergeTile d;
  d.xoff = xb + p0.x;
  d.yoff = yb + p0.y;
  d.cx = p1.x - p0.x;
  d.cy = p1.y - p0.y;
  d.pad0 = d.pad1 = 0;
  desc[t] = d;
}

constexpr int kMergeRowBytes = ((kTile + 1) * 4 + 64 + 15) & ~15; // X part + Y part, each starting on a 16-byte boundary with up to 15 bytes of lead
template <class V>
constexpr int merge_val_bytes() { return ((kTile + 1) * (int)sizeof(V) + 64 + 15) & ~15; }
constexpr int kMergeInf = 0x7FFFFFFF; // sorts behind every row id (local dimensions stay below 2^31 - 1)

// where a part of `count` elements of `esize` bytes starting at global address `addr` lands in a buffer: the copy starts
// at the 16-byte boundary below addr, so the first element sits `lead` bytes into the buffer part
struct MergePart {
  unsigned lead, bytes; // bytes: whole copy, a multiple of 16 (0 for an empty part)
};
__device__ __forceinline__ MergePart merge_part(unsigned long long addr, int count, int esize) {
  MergePart p;
  p.lead = (unsigned)(addr & 15ull);
  p.bytes = count > 0 ? ((p.lead + (unsigned)count * (unsigned)esize + 15u) & ~15u) : 0u;
  return p;
}

template <class SR, bool WRITE>
__global__ void __launch_bounds__(kMergeThreads)
merge2_tma_kernel(MergeCols m, const typename SR::out_t *xval, const typename SR::out_t *yval, const MergeTile *desc, int64_t ntiles,
                  int64_t *tile_count, const int64_t *tile_base, int32_t *cir, typename SR::out_t *cval) {
  typedef typename SR::out_t out_t;
  constexpr int VB = WRITE ? merge_val_bytes<out_t>() : 0;
  extern __shared__ __align__(128) unsigned char msm[];
  // two stages of: rows [kMergeRowBytes] | values [VB]
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ MergeTile cur[2];
  __shared__ int2 part[kMergeThreads + 1];
  __shared__ int warp_sums[kMergeThreads / 32];
  __shared__ long long base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((int64_t)blockIdx.x >= ntiles) return;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  // thread 0 issues the copies of a tile into stage b
  auto issue = [&](const MergeTile &d, int b) {
    unsigned char *rows = msm + (size_t)b * (kMergeRowBytes + VB);
    const unsigned long long xa = (unsigned long long)(m.xir + d.xoff), ya = (unsigned long long)(m.yir + d.yoff);
    const MergePart px = merge_part(xa, d.cx, 4), py = merge_part(ya, d.cy, 4);
    unsigned tx = px.bytes + py.bytes;
    MergePart vx{0, 0}, vy{0, 0};
    unsigned long long xv = 0, yv = 0;
    if (WRITE) {
      xv = (unsigned long long)(xval + d.xoff);
      yv = (unsigned long long)(yval + d.yoff);
      vx = merge_part(xv, d.cx, (int)sizeof(out_t));
      vy = merge_part(yv, d.cy, (int)sizeof(out_t));
      tx += vx.bytes + vy.bytes;
    }
    cur[b] = d;
    fence_proxy_async(); // the stage was read and written by ordinary accesses while it held the tile before
    mbar_arrive_expect_tx(&bar[b], tx);
    if (px.bytes) bulk_load(rows, reinterpret_cast<const void *>(xa & ~15ull), px.bytes, &bar[b]);
    if (py.bytes) bulk_load(rows + px.bytes, reinterpret_cast<const void *>(ya & ~15ull), py.bytes, &bar[b]);
    if (WRITE) {
      unsigned char *vals = rows + kMergeRowBytes;
      if (vx.bytes) bulk_load(vals, reinterpret_cast<const void *>(xv & ~15ull), vx.bytes, &bar[b]);
      if (vy.bytes) bulk_load(vals + vx.bytes, reinterpret_cast<const void *>(yv & ~15ull), vy.bytes, &bar[b]);
    }
  };
  MergeTile nextd; // thread 0: descriptor of the tile after the one in flight
  int64_t nextbase = 0;
  if (tid == 0) {
    issue(desc[blockIdx.x], 0);
    const int64_t t1 = (int64_t)blockIdx.x + gridDim.x;
    if (t1 < ntiles) nextd = desc[t1];
  }
  __syncthreads();
  int it = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int b = it & 1;
    const int64_t tn = t + gridDim.x;
    if (tid == 0) {
      if (WRITE) {
        bulk_wait_read(); // the stores of the tile that used the other stage have read it
        nextbase = tile_base[t];
      }
      if (tn < ntiles) {
        issue(nextd, b ^ 1);
        if (tn + gridDim.x < ntiles) nextd = desc[tn + gridDim.x];
      }
    }
    mbar_wait(&bar[b], (unsigned)((it >> 1) & 1));
    const MergeTile d = cur[b];
    unsigned char *rows = msm + (size_t)b * (kMergeRowBytes + VB);
    const int cx = d.cx, cy = d.cy, tot = cx + cy;
    const unsigned long long xa = (unsigned long long)(m.xir + d.xoff), ya = (unsigned long long)(m.yir + d.yoff);
    const MergePart px = merge_part(xa, cx, 4), py = merge_part(ya, cy, 4);
    const int *sx = reinterpret_cast<const int *>(rows + px.lead);
    const int *sy = reinterpret_cast<const int *>(rows + px.bytes + py.lead);
    const out_t *svx = nullptr, *svy = nullptr;
    if (WRITE) {
      const unsigned long long xv = (unsigned long long)(xval + d.xoff), yv = (unsigned long long)(yval + d.yoff);
      const MergePart vx = merge_part(xv, cx, (int)sizeof(out_t)), vy = merge_part(yv, cy, (int)sizeof(out_t));
      svx = reinterpret_cast<const out_t *>(rows + kMergeRowBytes + vx.lead);
      svy = reinterpret_cast<const out_t *>(rows + kMergeRowBytes + vx.bytes + vy.lead);
    }
    // one merge-path search per thread; its slice ends where the next thread's begins
    int a0, b0;
    merge_partition(SharedRows{sx}, cx, SharedRows{sy}, cy, min(tot, tid * kMergeItems), a0, b0);
    part[tid] = make_int2(a0, b0);
    if (tid == 0) part[kMergeThreads] = make_int2(cx, cy);
    __syncthreads();
    const int2 pe = part[tid + 1];
    const int a1 = pe.x, b1 = pe.y;
    int orow[kMergeItems + 1];
    out_t oval[WRITE ? kMergeItems + 1 : 1];
    int cnt = 0;
    {
      int a = a0, bq = b0;
      int ka = a < a1 ? sx[a] : kMergeInf, kb = bq < b1 ? sy[bq] : kMergeInf;
#pragma unroll
      for (int q = 0; q <= kMergeItems; ++q) {
        const bool takex = ka <= kb; // ties: X first, and its Y twin goes with it
        const bool twin = ka == kb;
        const int r = takex ? ka : kb;
        const bool valid = r != kMergeInf;
        orow[q] = r;
        if (WRITE) {
          out_t v = out_t();
          if (valid) {
            if (twin) v = SR::add(svy[bq], svx[a]); // SR::add(current, stored) with X stored first, MultiwayMerge.h:374
            else v = takex ? svx[a] : svy[bq];
          }
          oval[q] = v;
        }
        cnt += valid ? 1 : 0;
        if (valid && takex) {
          ++a;
          ka = a < a1 ? sx[a] : kMergeInf;
        }
        if (valid && (!takex || twin)) {
          ++bq;
          kb = bq < b1 ? sy[bq] : kMergeInf;
        }
      }
    }
    // block-wide exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const int v = __shfl_up_sync(0xFFFFFFFFu, incl, dd);
      if (lane >= dd) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads(); // also: everybody is done reading the tile
    int off = incl - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < kMergeThreads / 32; ++w) {
      const int v = warp_sums[w];
      if (w < warp) off += v;
      total += v;
    }
    if (!WRITE) {
      if (tid == 0) tile_count[t] = total;
      __syncthreads(); // part[] and warp_sums[] are reused by the next tile
      continue;
    }
    if (tid == 0) base_s = nextbase; // tile_base[t] for everybody
    // stage the outputs where their 16-byte phase equals that of their place in C, then store
    __syncthreads();
    const int64_t base = base_s;
    const unsigned rshift = (unsigned)(((unsigned long long)(cir + base) & 15ull) >> 2);
    const unsigned vshift = (unsigned)(((unsigned long long)(cval + base) & 15ull) / sizeof(out_t));
    int *orow_s = reinterpret_cast<int *>(rows) + rshift;
    out_t *oval_s = reinterpret_cast<out_t *>(rows + kMergeRowBytes) + vshift;
#pragma unroll
    for (int q = 0; q <= kMergeItems; ++q)
      if (q < cnt) {
        orow_s[off + q] = orow[q];
        oval_s[off + q] = oval[q];
      }
    __syncthreads();
    {
      // rows: elements [0, rh) before the first 16-byte boundary of C, [rh, rh + rmid) by bulk store, the rest after
      const int rh = min(total, (int)((4u - rshift) & 3u));
      const int rmid = (total - rh) & ~3;
      constexpr int VPER = 16 / (int)sizeof(out_t); // values per 16 bytes
      const int vh = min(total, (int)(((unsigned)VPER - vshift) & (unsigned)(VPER - 1)));
      const int vmid = (total - vh) & ~(VPER - 1);
      if (tid == 0) {
        fence_proxy_async(); // the staged outputs were written by ordinary stores
        if (rmid > 0) bulk_store(cir + base + rh, orow_s + rh, (unsigned)rmid * 4u);
        if (vmid > 0) bulk_store(cval + base + vh, oval_s + vh, (unsigned)vmid * (unsigned)sizeof(out_t));
        bulk_commit();
      }
      for (int q = tid; q < rh; q += kMergeThreads) cir[base + q] = orow_s[q];
      for (int q = rh + rmid + tid; q < total; q += kMergeThreads) cir[base + q] = orow_s[q];
      for (int q = tid; q < vh; q += kMergeThreads) cval[base + q] = oval_s[q];
      for (int q = vh + vmid + tid; q < total; q += kMergeThreads) cval[base + q] = oval_s[q];
    }
    __syncthreads(); // the stage may be refilled (after thread 0 has waited for the bulk stores to read it)
  }
  if (WRITE && tid == 0) bulk_wait_all();
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
  const bool tma = ctx->opt.merge_tma != 0;
  MergeTile *desc = nullptr;
  if (ntiles > 0 && tma) {
    CB_TRY(scratch.alloc(&desc, (size_t)ntiles));
    merge_desc_kernel<<<(unsigned)((ntiles + 255) / 256), 256, 0, st>>>(m, tile_col, first, tile_start, ntiles, desc);
    CB_LAUNCH_CHECK(ctx);
    auto kern = merge2_tma_kernel<SR, false>;
    const size_t sm = 2 * (size_t)kMergeRowBytes;
    CB_TRY(optin_smem(ctx, kern, sm));
    int per_sm = 0;
    CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMergeThreads, sm));
    const int64_t grid = std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * std::max(per_sm, 1));
    kern<<<(unsigned)grid, kMergeThreads, sm, st>>>(m, nullptr, nullptr, desc, ntiles, tile_count, nullptr, nullptr, nullptr);
    CB_LAUNCH_CHECK(ctx);
  } else if (ntiles > 0) {
    merge2_kernel<SR, false><<<(unsigned)ntiles, kMergeThreads, 0, st>>>(m, nullptr, nullptr, tile_col, first, tile_start, ntiles,
                                                                        tile_count, nullptr, nullptr, nullptr);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(exclusive_scan_i64(ctx, tile_count, tile_base, ntiles));
  int64_t nnz = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&nnz, tile_base + ntiles, 8, cudaMemcpyDeviceToHost, st));
  cudaEventRecord(ctx->ev[2], st);
  CB_CUDA(ctx, cudaStreamSynchronize(st));
  if (add_forbidden<SR>::value && nnz != X->nnz + Y->nnz) // BoolCopy1st/2ndSRing::add throws (Semirings.h:56-62)
    return set_error(ctx, CBGPU_ERR_INVALID, "Add should not happen (BoolCopy semiring): the merged lists share %lld entries",
                     (long long)(X->nnz + Y->nnz - nnz));
  cbgpu_mat_impl *C = nullptr;
  MatGuard cguard(ctx, &C);
  CB_TRY(mat_alloc(ctx, X->m, n, nnz, -1, X->dtype, &C));
  if (ntiles > 0 && tma) {
    auto kern = merge2_tma_kernel<SR, true>;
    const size_t sm = 2 * ((size_t)kMergeRowBytes + (size_t)merge_val_bytes<out_t>());
    CB_TRY(optin_smem(ctx, kern, sm));
    int per_sm = 0;
    CB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMergeThreads, sm));
    const int64_t grid = std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * std::max(per_sm, 1));
    kern<<<(unsigned)grid, kMergeThreads, sm, st>>>(m, reinterpret_cast<const out_t *>(X->numx), reinterpret_cast<const out_t *>(Y->numx),
                                                    desc, ntiles, nullptr, tile_base, C->ir, reinterpret_cast<out_t *>(C->numx));
    CB_LAUNCH_CHECK(ctx);
  } else if (ntiles > 0) {
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
