// MCL pruning epilogue of the phased multiply, on the device, for blocks that hold whole columns.
//
// Replaces (reference lines):
//   MCLPruneRecoverySelect                  ParFriends.h:186-354   (called per phase by MemEfficientSpGEMM, :744)
//     SpParMat::Prune(val <= hardThreshold) SpParMat.cpp (via Dcsc::Prune, dcsc.cpp:759)
//     Reduce(Column, plus) of values / ones ParFriends.h:199-201
//     SpParMat::Kselect1 (sparse)           SpParMat.cpp:1413-1700: k-th largest entry of a column; a column with fewer
//                                           than k entries yields its smallest entry, an empty one numeric_limits::min()
//     PruneColumn(pruneCols, std::less)     ParFriends.h:283,:339 (dcsc.cpp:900): drops val < threshold(column)
//   MakeColStochastic / Inflate             Applications/MCL.cpp:389-394, :431-437
//
// Every decision of the reference is per column, so one warp owns one stored column: statistics of the entries above the
// hard threshold, the choice recover / select / recover-after-select, a radix select for the k-th largest value where
// one is needed, and the count of survivors. A second kernel compacts the survivors in place order.
#include <float.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "util.cuh"

namespace cbgpu {

template <class T>
struct OrdKey;
template <>
struct OrdKey<double> {
  typedef unsigned long long U;
  static constexpr int BITS = 64;
  __device__ static __forceinline__ U key(double v) {
    U b = (U)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
  }
  __device__ static __forceinline__ double val(U k) {
    U b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
  }
  __device__ static __forceinline__ double tiny() { return DBL_MIN; }
};
template <>
struct OrdKey<float> {
  typedef unsigned int U;
  static constexpr int BITS = 32;
  __device__ static __forceinline__ U key(float v) {
    U b = (U)__float_as_int(v);
    return (b >> 31) ? ~b : (b | 0x80000000u);
  }
  __device__ static __forceinline__ float val(U k) {
    U b = (k >> 31) ? (k & 0x7FFFFFFFu) : ~k;
    return __int_as_float((int)b);
  }
  __device__ static __forceinline__ float tiny() { return FLT_MIN; }
};

__device__ __forceinline__ int prune_lane() { return threadIdx.x & 31; }

constexpr int kLongCol = 4096;      // columns with more entries than this are handled by a whole CTA
constexpr int kPruneThreads = 512;  // CTA size of the long-column kernels (4 histogram bins per thread)
constexpr int kDigitBits = 11;      // radix-select digit of the long-column kernel
constexpr int kBins = 1 << kDigitBits;
constexpr int kCandCap = 4096;      // candidates kept in shared memory once few enough share the selected prefix
static_assert(kBins == 4 * kPruneThreads, "find_digit gives every thread four bins");

// k-th largest of v[0..n), 1 <= k <= n, by one warp: radix select from the most significant byte of the ordered key
template <class T>
__device__ T warp_kth_largest(const T *v, int n, int k, int *hist /* 256 ints of this warp */) {
  typedef typename OrdKey<T>::U U;
  const int lane = prune_lane();
  U prefix = 0;
  int need = k;
  for (int shift = OrdKey<T>::BITS - 8; shift >= 0; shift -= 8) {
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    const U himask = (shift + 8 >= OrdKey<T>::BITS) ? (U)0 : (U)(~(U)0 << (shift + 8));
    for (int i = lane; i < n; i += 32) {
      const U key = OrdKey<T>::key(v[i]);
      if ((key & himask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255)], 1);
    }
    __syncwarp();
    // lane L owns the bins 255-8L .. 248-8L (descending); the digit is where the running count from the top reaches `need`
    int c[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = hist[255 - (8 * lane + j)];
      s += c[j];
    }
    int incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += x;
    }
    const int excl = incl - s;
    const bool mine = excl < need && need <= incl;
    const unsigned who = __ballot_sync(0xFFFFFFFFu, mine);
    const int src = __ffs(who) - 1; // exists: the total over all lanes is >= need
    int digit = 0, rest = 0;
    if (mine) {
      int r = need - excl;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (r > 0 && r <= c[j]) {
          digit = 255 - (8 * lane + j);
          rest = r;
          r = 0;
        } else if (r > 0) {
          r -= c[j];
        }
      }
    }
    digit = __shfl_sync(0xFFFFFFFFu, digit, src);
    need = __shfl_sync(0xFFFFFFFFu, rest, src);
    prefix |= (U)digit << shift;
    __syncwarp();
  }
  return OrdKey<T>::val(prefix);
}

// Kselect1's answer for one column (SpParMat.cpp:1672-1684)
template <class T>
__device__ T warp_kselect(const T *v, int n, long long k, int *hist) {
  if (n == 0) return OrdKey<T>::tiny();
  if ((long long)n >= k && k >= 1) return warp_kth_largest<T>(v, n, (int)k, hist);
  T m = v[0];
  for (int i = prune_lane(); i < n; i += 32) m = v[i] < m ? v[i] : m;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    T o = __shfl_xor_sync(0xFFFFFFFFu, m, d);
    m = o < m ? o : m;
  }
  return m;
}

template <class T>
__device__ __forceinline__ void warp_count_sum(const T *v, int n, T bound, bool strictly_above, int &cnt, T &sum) {
  int c = 0;
  T s = 0;
  for (int i = prune_lane(); i < n; i += 32) {
    const T x = v[i];
    const bool in = strictly_above ? (x > bound) : !(x < bound);
    if (in) {
      ++c;
      s += x;
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
  }
  cnt = c;
  sum = s;
}

template <class T>
__global__ void __launch_bounds__(256)
mcl_threshold_kernel(const int64_t *cp, const T *vals, int64_t nzc, T hard, long long selectNum, long long recoverNum,
                     T recoverPct, T *thr, int64_t *keep, unsigned long long *counters) {
  __shared__ int hist_all[8][256];
  const int warp = threadIdx.x >> 5, lane = prune_lane();
  const int64_t col = (int64_t)blockIdx.x * 8 + warp;
  if (col >= nzc) return;
  int *hist = hist_all[warp];
  const int64_t b = cp[col];
  if (cp[col + 1] - b > kLongCol) return; // a whole CTA takes the long columns (mcl_threshold_long_kernel)
  const int n = (int)(cp[col + 1] - b);
  const T *v = vals + b;
  int npr;
  T spr;
  warp_count_sum<T>(v, n, hard, true, npr, spr); // statistics of A.Prune(val <= hardThreshold)
  T t = hard;
  const bool recover = (long long)npr < recoverNum && n > npr && spr < recoverPct;
  if (recover) {
    t = warp_kselect<T>(v, n, recoverNum, hist);
    if (lane == 0) atomicAdd(&counters[0], 1ull);
  } else if (selectNum > 0 && (long long)npr > selectNum) {
    t = warp_kselect<T>(v, n, selectNum, hist);
    if (lane == 0) atomicAdd(&counters[1], 1ull);
    if (recoverNum > 0) { // recovery can be attempted after selection (ParFriends.h:288-331)
      int n1;
      T s1;
      warp_count_sum<T>(v, n, t, false, n1, s1);
      if ((long long)n1 < recoverNum && s1 < recoverPct) {
        t = warp_kselect<T>(v, n, recoverNum, hist);
        if (lane == 0) atomicAdd(&counters[2], 1ull);
      }
    }
  }
  int kept;
  T dummy;
  warp_count_sum<T>(v, n, t, false, kept, dummy); // PruneColumn(pruneCols, std::less): val < t goes
  if (lane == 0) {
    thr[col] = t;
    keep[col] = kept;
  }
}

template <class T>
__global__ void __launch_bounds__(256)
mcl_compact_kernel(const int64_t *cp, const int32_t *rows, const T *vals, int64_t nzc, const T *thr, const int64_t *optr,
                   int32_t *orows, T *ovals) {
  const int warp = threadIdx.x >> 5, lane = prune_lane();
  const int64_t col = (int64_t)blockIdx.x * 8 + warp;
  if (col >= nzc) return;
  const int64_t b = cp[col];
  if (cp[col + 1] - b > kLongCol) return;
  const int n = (int)(cp[col + 1] - b);
  const T t = thr[col];
  int64_t o = optr[col];
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    T x = T();
    int r = 0;
    bool in = false;
    if (i < n) {
      x = vals[b + i];
      r = rows[b + i];
      in = !(x < t);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
    if (in) {
      const int64_t p = o + __popc(m & ((1u << lane) - 1u));
      orows[p] = r;
      ovals[p] = x;
    }
    o += __popc(m);
  }
}

// ------------------------------------------------------------------------------------------------ long columns
// One CTA per column. The passes over the column are what costs (an expansion slab is far larger than L2), so the
// statistics of the hard threshold come with the first pass, the k-th largest value is found with 11-bit digits
// (two passes narrow a column of 1e6 values in (0,1] to a few hundred candidates, which then live in shared memory), and
// the survivor count falls out of the post-selection statistics.
template <class V>
__device__ __forceinline__ V block_sum(V v, V *scratch /* 32 */) {
  const int lane = prune_lane(), warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  __syncthreads(); // scratch free
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  V s = 0;
  for (int w = 0; w < nwarp; ++w) s += scratch[w];
  return s;
}

// digit (bin index) where the count accumulated from the highest bin reaches `need`; out = {digit, rank inside the bin,
// size of the bin}. Thread t owns the bins kBins-1-4t .. kBins-4-4t.
__device__ __forceinline__ void find_digit(const int *hist, int need, int *scratch /* 32 */, int *out /* 3 */) {
  const int lane = prune_lane(), warp = threadIdx.x >> 5;
  int c[4], s = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    c[j] = hist[kBins - 1 - (4 * (int)threadIdx.x + j)];
    s += c[j];
  }
  int incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += x;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  for (int w = 0; w < warp; ++w) incl += scratch[w];
  const int excl = incl - s;
  if (excl < need && need <= incl) {
    int r = need - excl;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (r > 0 && r <= c[j]) {
        out[0] = kBins - 1 - (4 * (int)threadIdx.x + j);
        out[1] = r;
        out[2] = c[j];
        r = 0;
      } else if (r > 0) {
        r -= c[j];
      }
    }
  }
  __syncthreads();
}

template <class T>
struct PruneShared {
  int hist[kBins];
  typename OrdKey<T>::U cand[kCandCap];
  int iscratch[32];
  T tscratch[32];
  int out[3];
  int ncand;
};

// Kselect1's answer for one long column, by the whole CTA
template <class T>
__device__ T block_kselect(const T *v, int n, long long k, PruneShared<T> &sh) {
  typedef typename OrdKey<T>::U U;
  const int tid = threadIdx.x;
  if (k < 1 || (long long)n < k) { // fewer than k entries: the smallest one
    T m = v[0];
    for (int i = tid; i < n; i += kPruneThreads) m = v[i] < m ? v[i] : m;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      T o = __shfl_xor_sync(0xFFFFFFFFu, m, d);
      m = o < m ? o : m;
    }
    __syncthreads();
    if (prune_lane() == 0) sh.tscratch[tid >> 5] = m;
    __syncthreads();
    for (int w = 0; w < (kPruneThreads >> 5); ++w) m = sh.tscratch[w] < m ? sh.tscratch[w] : m;
    __syncthreads();
    return m;
  }
  U prefix = 0;
  int need = (int)k, shift = OrdKey<T>::BITS, ncand = 0;
  bool in_smem = false;
  while (shift > 0) {
    const int w = shift >= kDigitBits ? kDigitBits : shift; // 64 = 5 x 11 + 9, 32 = 2 x 11 + 10
    const U himask = (shift >= OrdKey<T>::BITS) ? (U)0 : (U)(~(U)0 << shift); // bits fixed so far
    shift -= w;
    for (int i = tid; i < kBins; i += kPruneThreads) sh.hist[i] = 0;
    __syncthreads();
    const U dmask = (U)((1u << w) - 1u);
    if (!in_smem) {
      for (int i = tid; i < n; i += kPruneThreads) {
        const U key = OrdKey<T>::key(v[i]);
        if ((key & himask) == prefix) atomicAdd(&sh.hist[(int)((key >> shift) & dmask)], 1);
      }
    } else {
      for (int i = tid; i < ncand; i += kPruneThreads) {
        const U key = sh.cand[i];
        if ((key & himask) == prefix) atomicAdd(&sh.hist[(int)((key >> shift) & dmask)], 1);
      }
    }
    __syncthreads();
    find_digit(sh.hist, need, sh.iscratch, sh.out); // ends with a barrier
    const int digit = sh.out[0], cnt = sh.out[2];
    need = sh.out[1];
    prefix |= (U)digit << shift;
    if (shift == 0) break;
    if (!in_smem && cnt <= kCandCap) { // few enough values share the prefix: finish in shared memory
      if (tid == 0) sh.ncand = 0;
      __syncthreads();
      const U fixed = (U)(~(U)0 << shift);
      for (int i = tid; i < n; i += kPruneThreads) {
        const U key = OrdKey<T>::key(v[i]);
        if ((key & fixed) == prefix) sh.cand[atomicAdd(&sh.ncand, 1)] = key;
      }
      __syncthreads();
      ncand = sh.ncand;
      in_smem = true;
    }
    __syncthreads(); // sh.out / sh.hist are rewritten by the next level
  }
  __syncthreads();
  return OrdKey<T>::val(prefix);
}

template <class T>
__global__ void __launch_bounds__(kPruneThreads)
mcl_threshold_long_kernel(const int64_t *cp, const T *vals, const int32_t *list, T hard, long long selectNum,
                          long long recoverNum, T recoverPct, T *thr, int64_t *keep, unsigned long long *counters) {
  __shared__ PruneShared<T> sh;
  const int tid = threadIdx.x;
  const int64_t col = list[blockIdx.x];
  const int64_t b = cp[col];
  const int n = (int)(cp[col + 1] - b);
  const T *v = vals + b;
  // pass 1: statistics of A.Prune(val <= hardThreshold), and the survivors of the plain hard threshold
  int npr = 0, nge = 0;
  T spr = 0;
  for (int i = tid; i < n; i += kPruneThreads) {
    const T x = v[i];
    if (x > hard) {
      ++npr;
      spr += x;
    }
    if (!(x < hard)) ++nge;
  }
  npr = block_sum<int>(npr, sh.iscratch);
  nge = block_sum<int>(nge, sh.iscratch);
  spr = block_sum<T>(spr, sh.tscratch);
  T t = hard;
  long long kept = nge;
  const bool recover = (long long)npr < recoverNum && n > npr && spr < recoverPct;
  const int rule = recover ? 1 : ((selectNum > 0 && (long long)npr > selectNum) ? 2 : 0);
  if (rule) {
    t = block_kselect<T>(v, n, rule == 1 ? recoverNum : selectNum, sh);
    int again = 0;
    do {
      int n1 = 0;
      T s1 = 0;
      for (int i = tid; i < n; i += kPruneThreads) {
        const T x = v[i];
        if (!(x < t)) {
          ++n1;
          s1 += x;
        }
      }
      n1 = block_sum<int>(n1, sh.iscratch);
      s1 = block_sum<T>(s1, sh.tscratch);
      kept = n1;
      // recovery can be attempted after selection (ParFriends.h:288-331)
      if (rule == 2 && !again && recoverNum > 0 && (long long)n1 < recoverNum && s1 < recoverPct) {
        t = block_kselect<T>(v, n, recoverNum, sh);
        again = 1;
        if (tid == 0) atomicAdd(&counters[2], 1ull);
      } else {
        again = 0;
      }
    } while (again);
    if (tid == 0) atomicAdd(&counters[rule - 1], 1ull);
  }
  if (tid == 0) {
    thr[col] = t;
    keep[col] = kept;
  }
}

template <class T>
__global__ void __launch_bounds__(kPruneThreads)
mcl_compact_long_kernel(const int64_t *cp, const int32_t *rows, const T *vals, const int32_t *list, const T *thr,
                        const int64_t *optr, int32_t *orows, T *ovals) {
  __shared__ int wcnt[kPruneThreads / 32];
  const int tid = threadIdx.x, lane = prune_lane(), warp = tid >> 5;
  const int64_t col = list[blockIdx.x];
  const int64_t b = cp[col];
  const int n = (int)(cp[col + 1] - b);
  const T t = thr[col];
  int64_t o = optr[col];
  for (int base = 0; base < n; base += kPruneThreads) {
    const int i = base + tid;
    T x = T();
    int r = 0;
    bool in = false;
    if (i < n) {
      x = vals[b + i];
      r = rows[b + i];
      in = !(x < t);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
    if (lane == 0) wcnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kPruneThreads / 32; ++w) {
      const int c = wcnt[w];
      if (w < warp) before += c;
      total += c;
    }
    if (in) {
      const int64_t p = o + before + __popc(m & ((1u << lane) - 1u));
      orows[p] = r;
      ovals[p] = x;
    }
    o += total;
    __syncthreads();
  }
}

__global__ void long_cols_kernel(const int64_t *cp, int64_t nzc, int32_t *list, int *count) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nzc && cp[j + 1] - cp[j] > kLongCol) list[atomicAdd(count, 1)] = (int32_t)j;
}

template <class T>
static int mcl_prune_typed(cbgpu_ctx_impl *ctx, const cbgpu_mat_impl *A, double hard, int64_t selectNum, int64_t recoverNum,
                           double recoverPct, cbgpu_mat_impl **out, cbgpu_prune_stats *stats) {
  cudaStream_t st = ctx->stream;
  const int64_t nzc = A->nzc;
  cbgpu_prune_stats ps;
  memset(&ps, 0, sizeof(ps));
  ps.nnz_in = A->nnz;
  CB_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
  Scratch scratch(ctx); // frees the temporaries on every return
  T *thr = nullptr;
  int64_t *keep = nullptr, *optr = nullptr;
  unsigned long long *counters = nullptr;
  CB_TRY(scratch.alloc(&thr, (size_t)nzc + 1));
  CB_TRY(scratch.alloc(&keep, (size_t)nzc + 1));
  CB_TRY(scratch.alloc(&optr, (size_t)nzc + 2));
  CB_TRY(scratch.alloc(&counters, 4));
  CB_CUDA(ctx, cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned long long), st));
  int64_t nnz_out = 0;
  unsigned long long hc[4] = {0, 0, 0, 0};
  int32_t *list = nullptr;
  int *nlong_dev = nullptr;
  int nlong = 0;
  if (nzc >= (int64_t)1 << 31) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "too many columns");
  if (nzc > 0) {
    CB_TRY(scratch.alloc(&list, (size_t)nzc));
    CB_TRY(scratch.alloc(&nlong_dev, 1));
    CB_CUDA(ctx, cudaMemsetAsync(nlong_dev, 0, sizeof(int), st));
    long_cols_kernel<<<(unsigned)((nzc + 255) / 256), 256, 0, st>>>(A->cp, nzc, list, nlong_dev);
    CB_LAUNCH_CHECK(ctx);
    CB_CUDA(ctx, cudaMemcpyAsync(&nlong, nlong_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    mcl_threshold_kernel<T><<<(unsigned)((nzc + 7) / 8), 256, 0, st>>>(A->cp, reinterpret_cast<const T *>(A->numx), nzc, (T)hard,
                                                                      (long long)selectNum, (long long)recoverNum, (T)recoverPct,
                                                                      thr, keep, counters);
    CB_LAUNCH_CHECK(ctx);
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    if (nlong > 0) {
      mcl_threshold_long_kernel<T><<<(unsigned)nlong, kPruneThreads, 0, st>>>(A->cp, reinterpret_cast<const T *>(A->numx), list,
                                                                              (T)hard, (long long)selectNum, (long long)recoverNum,
                                                                              (T)recoverPct, thr, keep, counters);
      CB_LAUNCH_CHECK(ctx);
    }
    CB_TRY(exclusive_scan_i64(ctx, keep, optr, nzc));
    CB_CUDA(ctx, cudaMemcpyAsync(&nnz_out, optr + nzc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
  } else {
    CB_CUDA(ctx, cudaMemsetAsync(optr, 0, sizeof(int64_t), st));
  }
  cbgpu_mat_impl *C = nullptr;
  MatGuard cguard(ctx, &C);
  CB_TRY(mat_alloc(ctx, A->m, A->n, nnz_out, -1, A->dtype, &C));
  int rc = CBGPU_OK;
  if (nzc > 0 && nnz_out > 0) {
    mcl_compact_kernel<T><<<(unsigned)((nzc + 7) / 8), 256, 0, st>>>(A->cp, A->ir, reinterpret_cast<const T *>(A->numx), nzc, thr,
                                                                    optr, C->ir, reinterpret_cast<T *>(C->numx));
    ctx->launches++;
    if (nlong > 0) {
      mcl_compact_long_kernel<T><<<(unsigned)nlong, kPruneThreads, 0, st>>>(A->cp, A->ir, reinterpret_cast<const T *>(A->numx), list,
                                                                            thr, optr, C->ir, reinterpret_cast<T *>(C->numx));
      ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error(ctx, CBGPU_ERR_CUDA, "mcl_compact_kernel: %s", cudaGetErrorString(e));
  }
  if (rc == CBGPU_OK) rc = compact_columns(ctx, A->jc, optr, nzc, &C->jc, &C->cp, &C->nzc);
  if (rc != CBGPU_OK) return rc; // the guards release C and the temporaries
  CB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
  CB_CUDA(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ps.ms, ctx->ev[4], ctx->ev[5]);
  ps.nnz_out = nnz_out;
  ps.nzc_out = C->nzc;
  ps.cols_recovered = (int64_t)hc[0];
  ps.cols_selected = (int64_t)hc[1];
  ps.cols_recovered_after_select = (int64_t)hc[2];
  if (stats) *stats = ps;
  *out = C;
  cguard.armed = false;
  return CBGPU_OK;
}

// ------------------------------------------------------------------------------------------------ column scaling
// mode 0: v *= 1 / sum(column)                       MakeColStochastic (MCL.cpp:389-394)
// mode 1: v = pow(v, power), then mode 0             Inflate           (MCL.cpp:431-437)
template <class T>
__global__ void __launch_bounds__(256) col_scale_kernel(const int64_t *cp, T *vals, int64_t nzc, int mode, double power) {
  const int warp = threadIdx.x >> 5, lane = prune_lane();
  const int64_t col = (int64_t)blockIdx.x * 8 + warp;
  if (col >= nzc) return;
  const int64_t b = cp[col];
  const int n = (int)(cp[col + 1] - b);
  T *v = vals + b;
  T s = 0;
  for (int i = lane; i < n; i += 32) {
    T x = v[i];
    if (mode == 1) {
      x = (T)pow((double)x, power);
      v[i] = x;
    }
    s += x;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
  const T inv = (s == (T)0) ? (sizeof(T) == 8 ? (T)DBL_MAX : (T)FLT_MAX) : (T)1 / s; // safemultinv (Operations.h)
  __syncwarp();
  for (int i = lane; i < n; i += 32) v[i] = v[i] * inv;
}

static int col_scale(cbgpu_ctx_impl *ctx, cbgpu_mat_impl *A, int mode, double power) {
  if (A->dtype != CBGPU_F64 && A->dtype != CBGPU_F32)
    return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "column scaling needs a floating-point block");
  if (A->nzc > 0) {
    unsigned grid = (unsigned)((A->nzc + 7) / 8);
    if (A->dtype == CBGPU_F64)
      col_scale_kernel<double><<<grid, 256, 0, ctx->stream>>>(A->cp, reinterpret_cast<double *>(A->numx), A->nzc, mode, power);
    else
      col_scale_kernel<float><<<grid, 256, 0, ctx->stream>>>(A->cp, reinterpret_cast<float *>(A->numx), A->nzc, mode, power);
    CB_LAUNCH_CHECK(ctx);
  }
  // the cached window-major copy holds the old values
  dev_free(ctx, A->win_T2);
  dev_free(ctx, A->win_ir);
  dev_free(ctx, A->win_val);
  A->win_T2 = nullptr;
  A->win_ir = nullptr;
  A->win_val = nullptr;
  A->win_log2 = 0;
  A->win_nwin = 0;
  return CBGPU_OK;
}

} // namespace cbgpu

using namespace cbgpu;

extern "C" {

int cbgpu_mcl_prune(cbgpu_ctx *ctx, const cbgpu_mat *A, double hardThreshold, int64_t selectNum, int64_t recoverNum,
                    double recoverPct, cbgpu_mat **out, cbgpu_prune_stats *stats) {
  if (!ctx || !A || !out) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (A->dtype == CBGPU_F64) return mcl_prune_typed<double>(ctx, A, hardThreshold, selectNum, recoverNum, recoverPct, out, stats);
  if (A->dtype == CBGPU_F32) return mcl_prune_typed<float>(ctx, A, hardThreshold, selectNum, recoverNum, recoverPct, out, stats);
  return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "MCL pruning needs a floating-point block (dtype %d)", A->dtype);
}

/* MemEfficientSpGEMM at P = 1 (ParFriends.h:452-777): column slabs of B (ColSplit rule), one multiply per slab, every
 * finished slab pruned in HBM by MCLPruneRecoverySelect (:744) before the next one is multiplied, ColConcatenate (:772). */
int cbgpu_memefficient_spgemm(cbgpu_ctx *ctx, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                              double hardThreshold, int64_t selectNum, int64_t recoverNum, double recoverPct, cbgpu_mat **C,
                              cbgpu_memeff_stats *stats) {
  if (!ctx || !A || !B || !C) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_memeff_stats ms;
  memset(&ms, 0, sizeof(ms));
  struct Events { // destroyed on every return
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } evs;
  CB_CUDA(ctx, cudaEventCreate(&evs.a));
  CB_CUDA(ctx, cudaEventCreate(&evs.b));
  cudaEvent_t e0 = evs.a, e1 = evs.b;
  cudaEventRecord(e0, ctx->stream);
  if (phases <= 0) {
    // CalculateNumberOfPhases (ParFriends.h:780-843) with the exact symbolic count instead of an estimate: an unpruned
    // slab (12 B per entry) may take a quarter of the HBM that is free right now
    int64_t flops = 0, nnz = 0;
    CB_TRY(cbgpu_spgemm_symbolic(ctx, A, B, &flops, &nnz));
    size_t freeb = 0, totalb = 0;
    CB_CUDA(ctx, cudaMemGetInfo(&freeb, &totalb));
    const double budget = (double)std::max<size_t>(freeb / 4, (size_t)1 << 30);
    phases = (int)std::min<double>(std::max<double>(1.0, ceil((double)nnz * 12.0 / budget)), (double)std::max<int64_t>(1, B->n));
  }
  phases = (int)std::min<int64_t>(std::max<int64_t>(1, phases), std::max<int64_t>(1, B->n));
  ms.phases = phases;
  std::vector<cbgpu_mat *> slabs(phases, nullptr), pieces(phases, nullptr);
  int rc = CBGPU_OK;
  if (phases > 1) rc = cbgpu_mat_colsplit(ctx, B, phases, slabs.data());
  for (int p = 0; p < phases && rc == CBGPU_OK; ++p) {
    const cbgpu_mat *Bs = phases > 1 ? slabs[p] : B;
    cbgpu_mat *Cs = nullptr;
    cbgpu_stats st;
    memset(&st, 0, sizeof(st));
    rc = cbgpu_spgemm_local(ctx, semiring, A, Bs, &Cs, &st);
    if (rc != CBGPU_OK) break;
    ms.flops += st.flops;
    ms.nnz_unpruned += st.nnz_out;
    ms.ms_multiply += st.ms_total;
    if (Cs->dtype == CBGPU_F64 || Cs->dtype == CBGPU_F32) {
      cbgpu_prune_stats ps;
      rc = cbgpu_mcl_prune(ctx, Cs, hardThreshold, selectNum, recoverNum, recoverPct, &pieces[p], &ps);
      mat_release(ctx, Cs);
      if (rc != CBGPU_OK) break;
      ms.ms_prune += ps.ms;
      ms.cols_recovered += ps.cols_recovered;
      ms.cols_selected += ps.cols_selected;
      ms.cols_recovered_after_select += ps.cols_recovered_after_select;
    } else {
      pieces[p] = Cs; // the reference instantiates the pruning for floating-point results only
    }
  }
  cbgpu_mat *out = nullptr;
  if (rc == CBGPU_OK) {
    if (phases > 1) rc = cbgpu_mat_colconcat(ctx, phases, pieces.data(), &out);
    else {
      out = pieces[0];
      pieces[0] = nullptr;
    }
  }
  for (int p = 0; p < phases; ++p) {
    mat_release(ctx, pieces[p]);
    if (phases > 1) mat_release(ctx, slabs[p]);
  }
  if (rc == CBGPU_OK) {
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms.ms_total, e0, e1);
    ms.nnz_out = out->nnz;
    if (stats) *stats = ms;
    *C = out;
  }
  return rc;
}

int cbgpu_mat_make_col_stochastic(cbgpu_ctx *ctx, cbgpu_mat *A) {
  if (!ctx || !A) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  return col_scale(ctx, A, 0, 1.0);
}

int cbgpu_mat_inflate(cbgpu_ctx *ctx, cbgpu_mat *A, double power) {
  if (!ctx || !A) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  return col_scale(ctx, A, 1, power);
}

} // extern "C"
