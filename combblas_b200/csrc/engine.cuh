// Column-accumulation engine: the sm_100a replacement for the reference's per-column loops
//   estimateFLOP (mtSpGEMM.h:1058), estimateNNZ_Hash (:807), the numeric hash column of
//   LocalHybridSpGEMM / LocalSpGEMMHash (:362-440, :552-634) and the merge column of
//   MultiwayMerge[Hash] (MultiwayMerge.h:194-248, :338-422).
//
// One engine serves both the multiply and the k-way merge: a TASK is (output column, row window) and a
// task's input is a list of SEGMENTS -- contiguous runs of (row, value) in one array pair:
//   multiply: one segment per nonzero B(k,j): the part of A(:,k) inside the row window, scaled by B(k,j)
//   merge   : one segment per input list: the part of L_s(:,j) inside the row window
// Segment bounds come from a monotone table T[col*nwin + w] (for one window this is the dense column
// pointer array itself; it replaces Dcsc::ConstructAux/FillColInds, dcsc.cpp:1043,:1363).
//
// Per task the engine runs a symbolic pass (count distinct rows) and, after an exclusive scan over the
// counts, a numeric pass that writes the column's rows ASCENDING together with the accumulated values:
//   hash path   (small tasks) : shared-memory open addressing per warp or per CTA + bitonic sort of the hits
//   bitmap path (large tasks) : shared-memory presence bitmap of the row window in 32-row words + per-word rank prefix:
//                               two loads and a popcount give every row its sorted output slot, so there is no sort
//                               and no probing; values accumulate in shared memory (exchange protocol) for every task
//                               whose outputs fit, straight into C in HBM (L2 reductions) for the few that do not.
// No tensor cores: the work is irregular integer/atomic traffic; the levers are coalesced segment reads,
// shared-memory atomics, and keeping every SM busy with size-ordered tasks.
#pragma once
#include "common.cuh"
#include "semiring.cuh"
#include "bulk.cuh"

namespace cbgpu {

constexpr int kWarpLong = 128;   // segments are walked by a whole warp in groups of this many products (4 loads per lane)
constexpr int kBitmapThreads = 512;

// ------------------------------------------------------------------------------------------------ source
template <class SR, bool MERGE>
struct Source {
  typedef typename std::conditional<MERGE, typename SR::out_t, typename SR::a_t>::type aval_t;
  typedef typename SR::b_t mult_t;
  const int64_t *T;   // dense column pointers of A (or of the concatenated merge lists): whole columns
  int nwin;
  int wlog2;
  const int32_t *Air;
  const aval_t *Aval;
  // window-major copy of A: piece (window w, column c) starts at T2[w*N+c] & ~3 of Wir/Wval (16-byte aligned, padded to a
  // multiple of 4 entries with row -1; see util.cu), so that all tasks of one row window gather from one contiguous,
  // L2-sized slice instead of striding through every column of A, four products per load
  const int64_t *T2;
  int64_t N;
  const int32_t *Wir;
  const aval_t *Wval;
  // multiply
  const int64_t *Bcp;
  const int32_t *Bir;
  const mult_t *Bval;
  // merge
  int k;
  int64_t n;
  // tasks (null task_col: task t is column t over all windows)
  const int32_t *task_col;
  const uint32_t *task_win;
};

struct Task {
  int col;      // multiply: index into B's non-empty columns; merge: column id
  int wlo, whi; // row windows [wlo, whi): either one window or all of them
  int64_t seg_begin, seg_end;
  const int32_t *rows; // where this task's segments live: A itself (all rows) or its window-major copy
  const void *vals;
};

template <class Src>
__device__ __forceinline__ Task load_task(const Src &s, int t) {
  Task k;
  if (s.task_col) {
    k.col = s.task_col[t];
    unsigned w = s.task_win[t];
    k.wlo = (int)(w & 0xFFFFu);
    k.whi = (int)(w >> 16);
  } else {
    k.col = t;
    k.wlo = 0;
    k.whi = s.nwin;
  }
  return k;
}

template <class SR, bool MERGE>
__device__ __forceinline__ void task_segments(const Source<SR, MERGE> &s, Task &k) {
  if (MERGE) {
    k.seg_begin = 0;
    k.seg_end = s.k;
  } else {
    k.seg_begin = s.Bcp[k.col];
    k.seg_end = s.Bcp[k.col + 1];
  }
  const bool whole = (k.whi - k.wlo) == s.nwin || s.T2 == nullptr;
  k.rows = whole ? s.Air : s.Wir;
  k.vals = whole ? (const void *)s.Aval : (const void *)s.Wval;
}

// Launch-order task record: everything a bitmap kernel needs to start a task, in one 32-byte load (the chain order[] ->
// task_col/task_win -> Bcp -> taskptr -> slot_of_task was five dependent loads; small tasks are latency bound).
struct TaskRec {
  long long seg_begin; // multiply: first entry of the column of B; merge: the column id
  long long obase;     // numeric: first output position; symbolic: the task id (tasknnz index)
  int nseg;            // multiply: entries of the column of B
  int nnz;             // numeric: outputs of the task
  int slot;            // hand-over slot of the symbolic pass, -1 = none
  unsigned win;        // wlo | whi << 16
};

template <class SR, bool MERGE>
__global__ void task_record_kernel(Source<SR, MERGE> s, const int32_t *order, int64_t count, const int64_t *taskptr,
                                   const int32_t *slot_of_task, TaskRec *recs, const int64_t *taskflop = nullptr,
                                   int64_t save_min_flop = 0, int save_cap = 0, int *save_counter = nullptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int t = order[i];
  Task k = load_task(s, t);
  TaskRec r;
  r.win = (unsigned)k.wlo | ((unsigned)k.whi << 16);
  if (MERGE) {
    r.seg_begin = k.col;
    r.nseg = s.k;
  } else {
    r.seg_begin = s.Bcp[k.col];
    r.nseg = (int)(s.Bcp[k.col + 1] - r.seg_begin);
  }
  if (taskptr) {
    r.obase = taskptr[t];
    r.nnz = (int)(taskptr[t + 1] - r.obase);
  } else {
    r.obase = t;
    r.nnz = 0;
  }
  r.slot = slot_of_task ? slot_of_task[t] : -1;
  // symbolic records: tasks with enough products get a hand-over slot for their presence words, while slots last
  if (save_counter && taskflop[t] >= save_min_flop && ((r.win >> 16) - (r.win & 0xFFFFu)) == 1u) {
    const int sl = atomicAdd(save_counter, 1);
    r.slot = sl < save_cap ? sl : -1;
  }
  recs[i] = r;
}

template <class SR, bool MERGE>
__device__ __forceinline__ Task task_from_record(const Source<SR, MERGE> &s, const TaskRec &r) {
  Task k;
  k.wlo = (int)(r.win & 0xFFFFu);
  k.whi = (int)(r.win >> 16);
  if (MERGE) {
    k.col = (int)r.seg_begin;
    k.seg_begin = 0;
    k.seg_end = s.k;
  } else {
    k.col = -1;
    k.seg_begin = r.seg_begin;
    k.seg_end = r.seg_begin + r.nseg;
  }
  const bool whole = (k.whi - k.wlo) == s.nwin || s.T2 == nullptr;
  k.rows = whole ? s.Air : s.Wir;
  k.vals = whole ? (const void *)s.Aval : (const void *)s.Wval;
  return k;
}

// segment p of task k, first half: the column of A (or of the stacked merge lists) it reads, and its multiplier
template <class SR, bool MERGE, bool NEED_MULT>
__device__ __forceinline__ int64_t segment_column(const Source<SR, MERGE> &s, const Task &k, int64_t p, typename SR::b_t &mult) {
  if (MERGE) return p * s.n + k.col;
  if (NEED_MULT) mult = s.Bval[p];
  return s.Bir[p];
}

// second half: the run [beg, beg+len) of that column inside the task's row window(s)
template <class SR, bool MERGE>
__device__ __forceinline__ void segment_range(const Source<SR, MERGE> &s, const Task &k, int64_t col, int64_t &beg, int &len) {
  if ((k.whi - k.wlo) == s.nwin) {
    beg = s.T[col];
    len = (int)(s.T[col + 1] - beg);
  } else if (s.T2 != nullptr) {
    // window-major copy: starts are multiples of 4, the low two bits count the pad entries of the piece (util.cu)
    const int64_t *t = s.T2 + (int64_t)k.wlo * s.N + col;
    const int64_t a = t[0];
    beg = a & ~(int64_t)3;
    len = (int)((t[1] & ~(int64_t)3) - beg - (a & 3));
  } else {
    // no window-major copy (merge: few, long segments): cut the row-sorted column by binary search
    int64_t b0 = s.T[col], e0 = s.T[col + 1];
    const int64_t lo = (int64_t)k.wlo << s.wlog2, hi = (int64_t)k.whi << s.wlog2;
    int64_t a = b0, b = e0;
    while (a < b) {
      int64_t mid = (a + b) >> 1;
      if ((int64_t)s.Air[mid] < lo) a = mid + 1;
      else b = mid;
    }
    beg = a;
    b = e0;
    while (a < b) {
      int64_t mid = (a + b) >> 1;
      if ((int64_t)s.Air[mid] < hi) a = mid + 1;
      else b = mid;
    }
    len = (int)(a - beg);
  }
}

// segment p of task k: [beg, beg+len) in Air/Aval, with its multiplier
template <class SR, bool MERGE, bool NEED_MULT>
__device__ __forceinline__ void load_segment(const Source<SR, MERGE> &s, const Task &k, int64_t p, int64_t &beg, int &len,
                                             typename SR::b_t &mult) {
  const int64_t col = segment_column<SR, MERGE, NEED_MULT>(s, k, p, mult);
  segment_range<SR, MERGE>(s, k, col, beg, len);
}

// ------------------------------------------------------------------------------------------------ product walk
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <class M>
__device__ __forceinline__ M shfl_mult(M v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
template <>
__device__ __forceinline__ uint8_t shfl_mult<uint8_t>(uint8_t v, int src) {
  return (uint8_t)__shfl_sync(0xFFFFFFFFu, (int)v, src);
}

// shared-memory workspace of the balanced CTA walk (one chunk of up to 512 segments at a time)
template <int CHUNK>
struct CtaQueueT {
  long long pre[CHUNK + 1];         // exclusive prefix of the segment lengths of the chunk (in products)
  long long beg[CHUNK];             // first position of every segment
  unsigned long long mult[CHUNK];   // raw bits of the multiplier
  long long warp_sums[32];
};
struct NoQueue {};

template <class M>
__device__ __forceinline__ unsigned long long mult_bits(M v) {
  unsigned long long r = 0;
  memcpy(&r, &v, sizeof(M));
  return r;
}
template <class M>
__device__ __forceinline__ M bits_mult(unsigned long long r) {
  M v;
  memcpy(&v, &r, sizeof(M));
  return v;
}

template <class V>
struct RowVal {
  int row;
  V val;
};

// One warp processes up to 32 segments, one held per lane as (beg, len, mult). Each product is loaded by ld(pos) and
// consumed by use(loaded, mult) by exactly one lane. Long segments are strided by the whole warp (coalesced) with four
// independent loads in flight per lane; the rest are flattened so that all 32 lanes stay busy on short A-columns (two
// batches in flight). Splitting load from use is what lets the loads of several products overlap: the consumers are
// atomics, which the compiler will not reorder loads across.
template <class mult_t, bool NEED_MULT, class LD, class USE>
__device__ __forceinline__ void warp_process32(int64_t beg, int len, mult_t mult, LD &&ld, USE &&use) {
  const int lane = lane_id();
  // long segments: whole groups of kWarpLong products, strided by the warp (coalesced), four loads in flight per lane
  unsigned longmask = __ballot_sync(0xFFFFFFFFu, len >= kWarpLong);
  while (longmask) {
    int src = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    int64_t b = __shfl_sync(0xFFFFFFFFu, beg, src);
    int l = __shfl_sync(0xFFFFFFFFu, len, src) & ~(kWarpLong - 1);
    mult_t mu = mult_t();
    if (NEED_MULT) mu = shfl_mult<mult_t>(mult, src);
#pragma unroll 1
    for (int i = lane; i < l; i += 128) {
      auto x0 = ld(b + i);
      auto x1 = ld(b + i + 32);
      auto x2 = ld(b + i + 64);
      auto x3 = ld(b + i + 96);
      use(x0, mu);
      use(x1, mu);
      use(x2, mu);
      use(x3, mu);
    }
  }
  if (len >= kWarpLong) {
    const int done = len & ~(kWarpLong - 1);
    beg += done;
    len -= done;
  }
  // what is left (< kWarpLong products per segment): whole groups of 32 are still strided by the warp, one to three
  // loads in flight; this costs a third of the instructions of the flattened path below
  unsigned midmask = __ballot_sync(0xFFFFFFFFu, len >= 32);
  while (midmask) {
    int src = __ffs(midmask) - 1;
    midmask &= midmask - 1;
    const int64_t b = __shfl_sync(0xFFFFFFFFu, beg, src) + lane;
    const int n = __shfl_sync(0xFFFFFFFFu, len, src) >> 5; // 1..3, warp-uniform
    mult_t mu = mult_t();
    if (NEED_MULT) mu = shfl_mult<mult_t>(mult, src);
    auto x0 = ld(b);
    decltype(x0) x1{}, x2{};
    if (n > 1) x1 = ld(b + 32);
    if (n > 2) x2 = ld(b + 64);
    use(x0, mu);
    if (n > 1) use(x1, mu);
    if (n > 2) use(x2, mu);
  }
  if (len >= 32) {
    const int done = len & ~31;
    beg += done;
    len -= done;
  }
  // the tails (< 32 products per segment) are flattened over the lanes
  int slen = len;
  int incl = slen;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += v;
  }
  int excl = incl - slen;
  int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  for (int t0 = 0; t0 < total; t0 += 64) {
    int ta = t0 + lane, tb = t0 + 32 + lane;
    int ia = 0, ib = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      int ea = __shfl_sync(0xFFFFFFFFu, excl, ia + step);
      int eb = __shfl_sync(0xFFFFFFFFu, excl, ib + step);
      if (ea <= ta) ia += step;
      if (eb <= tb) ib += step;
    }
    int e0a = __shfl_sync(0xFFFFFFFFu, excl, ia), e0b = __shfl_sync(0xFFFFFFFFu, excl, ib);
    int64_t ba = __shfl_sync(0xFFFFFFFFu, beg, ia), bb = __shfl_sync(0xFFFFFFFFu, beg, ib);
    mult_t ma = mult_t(), mb = mult_t();
    if (NEED_MULT) {
      ma = shfl_mult<mult_t>(mult, ia);
      mb = shfl_mult<mult_t>(mult, ib);
    }
    const bool va = ta < total, vb = tb < total;
    decltype(ld(0)) xa{}, xb{};
    if (va) xa = ld(ba + (ta - e0a));
    if (vb) xb = ld(bb + (tb - e0b));
    if (va) use(xa, ma);
    if (vb) use(xb, mb);
  }
}

// One warp walks all segments of a task, 32 at a time (small tasks: one task per warp).
template <class SR, bool MERGE, bool NEED_MULT, class LD, class USE>
__device__ __forceinline__ void warp_walk(const Source<SR, MERGE> &s, const Task &k, LD &&ld, USE &&use) {
  typedef typename SR::b_t mult_t;
  const int lane = lane_id();
  for (int64_t base = k.seg_begin; base < k.seg_end; base += 32) {
    int64_t p = base + lane;
    int64_t beg = 0;
    int len = 0;
    mult_t mult = mult_t();
    if (p < k.seg_end) load_segment<SR, MERGE, NEED_MULT>(s, k, p, beg, len, mult);
    warp_process32<mult_t, NEED_MULT>(beg, len, mult, ld, use);
  }
}

// Whole-CTA walk with equal shares of PRODUCTS per warp: the CTA stages a chunk of segments in shared memory with the
// running sum of their lengths; warp w then owns products [w*P/nw, (w+1)*P/nw) of the chunk, wherever the segment
// boundaries fall, so one very long A-column can no longer stall the other warps at the barrier. All threads must call.
template <class SR, bool MERGE, bool NEED_MULT, int CH, class LD, class USE>
__device__ __forceinline__ void cta_walk(const Source<SR, MERGE> &s, const Task &k, CtaQueueT<CH> *q, LD &&ld, USE &&use) {
  typedef typename SR::b_t mult_t;
  constexpr int nwarp = CH >> 5; // blockDim.x == CH
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (int64_t cbase = k.seg_begin; cbase < k.seg_end; cbase += CH) {
    __syncthreads(); // previous chunk fully consumed
    const int nseg = (int)min((int64_t)CH, k.seg_end - cbase);
    int64_t beg = 0;
    int len = 0;
    mult_t mult = mult_t();
    if ((int)threadIdx.x < nseg) load_segment<SR, MERGE, NEED_MULT>(s, k, cbase + threadIdx.x, beg, len, mult);
    // block-wide exclusive scan of len (64-bit)
    long long incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      long long v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) q->warp_sums[warp] = incl;
    __syncthreads();
    // offsets of the warps' segment groups: inclusive scan of the per-warp sums by the first nwarp lanes of every warp
    long long wv = lane < nwarp ? q->warp_sums[lane] : 0, winc = wv;
#pragma unroll
    for (int d = 1; d < nwarp; d <<= 1) {
      long long x = __shfl_up_sync(0xFFFFFFFFu, winc, d);
      if (lane >= d) winc += x;
    }
    const long long total = __shfl_sync(0xFFFFFFFFu, winc, nwarp - 1);
    const long long woff = __shfl_sync(0xFFFFFFFFu, winc - wv, warp);
    if ((int)threadIdx.x < CH) {
      q->pre[threadIdx.x] = woff + incl - len;
      q->beg[threadIdx.x] = beg;
      if (NEED_MULT) q->mult[threadIdx.x] = mult_bits<mult_t>(mult);
    }
    if (threadIdx.x == 0) q->pre[CH] = total;
    __syncthreads();
    if (total == 0) continue;
    const long long lo = total * warp / nwarp, hi = total * (warp + 1) / nwarp;
    if (hi <= lo) continue;
    // first segment whose range contains product `lo`: largest i with pre[i] <= lo
    int a = 0, b = nseg - 1;
    while (a < b) {
      int mid = (a + b + 1) >> 1;
      if (q->pre[mid] <= lo) a = mid;
      else b = mid - 1;
    }
    for (int sb = a; sb < nseg; sb += 32) {
      if (q->pre[sb] >= hi) break; // warp-uniform
      int si = sb + lane;
      int64_t mybeg = 0;
      int mylen = 0;
      mult_t mymult = mult_t();
      if (si < nseg) {
        long long pb = q->pre[si], pe = (si + 1 < nseg) ? q->pre[si + 1] : total;
        long long x = pb > lo ? pb : lo, y = pe < hi ? pe : hi;
        if (y > x) {
          mybeg = q->beg[si] + (x - pb);
          mylen = (int)(y - x);
          if (NEED_MULT) mymult = bits_mult<mult_t>(q->mult[si]);
        }
      }
      warp_process32<mult_t, NEED_MULT>(mybeg, mylen, mymult, ld, use);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ flat walk
// Whole-CTA walk over the PRODUCTS of a task, one product per lane and sub-step, whatever the segment lengths are.
// A chunk of up to CH segments is staged in shared memory: the empty ones are dropped, every other one leaves the
// exclusive prefix of the lengths (pre), its multiplier, and the addresses rowp / valp of where product 0 of the CHUNK
// would sit if the segment started there, so that product q of the chunk is read at rowp[seg] + 4 q (one IMAD.WIDE).
// Warp w owns the products [w*T/nw, (w+1)*T/nw) of the chunk and keeps a cursor: the segment of its next product. Per
// step of up to 128 products the lanes read the next 32 boundaries pre[cursor+1+lane] once; a boundary that falls L
// products ahead sets bit L & 31 of mask L >> 5; four REDUX.OR merge the bits (segments are non-empty, so boundaries are
// distinct) and popcounts below the lane give every lane the segment of each of its four products: no search, no
// per-segment loop, loads coalesced inside every segment, four independent loads in flight per lane.
// The average segment of the large R-MAT tasks is 46..80 products (30..40 with 2^16-row windows); the segment-by-segment
// walk above (cta_walk) spends about 50 instructions per segment, the first version of this walk (one sub-step per
// boundary load) spent 68 per 32 products (profiles/r2_ncu_flatwalk_v1.txt).
template <int CH>
struct FlatQueueT {
  long long rowp[CH];
  long long valp[CH];
  unsigned long long mult[CH];
  unsigned pre[CH + 34];
  unsigned long long warp_tot[32];
  unsigned total; // products (groups) of the staged chunk
  int nseg;       // its non-empty segments
};

constexpr int kFlatSub = 4; // sub-steps of 32 products per boundary-window load

template <class SR, bool MERGE, bool NEED_VAL, int CH, class USE>
__device__ __forceinline__ void flat_walk(const Source<SR, MERGE> &s, const Task &k, FlatQueueT<CH> *q, USE &&use) {
  typedef typename SR::b_t mult_t;
  typedef typename Source<SR, MERGE>::aval_t aval_t;
  constexpr int nwarp = CH >> 5; // blockDim.x == CH
  constexpr unsigned long long kLenMask = (1ull << 40) - 1ull;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned le_mask = 0xFFFFFFFFu >> (31 - lane); // bits 0..lane
  for (int64_t cbase = k.seg_begin; cbase < k.seg_end; cbase += CH) {
    __syncthreads(); // previous chunk fully consumed
    int64_t beg = 0;
    int len = 0;
    mult_t mult = mult_t();
    if (cbase + threadIdx.x < k.seg_end) load_segment<SR, MERGE, NEED_VAL>(s, k, cbase + threadIdx.x, beg, len, mult);
    // one block-wide exclusive scan of (non-empty ? 1 : 0, len) packed as count << 40 | products
    const unsigned long long x = (len > 0 ? (1ull << 40) : 0ull) | (unsigned long long)(unsigned)len;
    unsigned long long incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) q->warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long wv = lane < nwarp ? q->warp_tot[lane] : 0ull, winc = wv;
#pragma unroll
    for (int d = 1; d < nwarp; d <<= 1) {
      unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, winc, d);
      if (lane >= d) winc += v;
    }
    const unsigned long long total = __shfl_sync(0xFFFFFFFFu, winc, nwarp - 1);
    const unsigned long long excl = __shfl_sync(0xFFFFFFFFu, winc - wv, warp) + incl - x;
    const int nseg = (int)(total >> 40);
    const unsigned T = (unsigned)(total & kLenMask); // products of the chunk: CH segments of at most 2^19 rows
    if (len > 0) {
      const int idx = (int)(excl >> 40);
      const long long pre = (long long)(excl & kLenMask);
      q->pre[idx] = (unsigned)pre;
      q->rowp[idx] = (long long)(k.rows + beg) - 4ll * pre;
      if (NEED_VAL) {
        q->valp[idx] = (long long)(reinterpret_cast<const aval_t *>(k.vals) + beg) - (long long)sizeof(aval_t) * pre;
        q->mult[idx] = mult_bits<mult_t>(mult);
      }
    }
    if (threadIdx.x < 34) q->pre[nseg + threadIdx.x] = threadIdx.x == 0 ? T : 0xFFFFFFFFu;
    __syncthreads();
    if (T == 0) continue;
    const unsigned lo = (unsigned)((unsigned long long)T * (unsigned)warp / (unsigned)nwarp);
    const unsigned hi = (unsigned)((unsigned long long)T * (unsigned)(warp + 1) / (unsigned)nwarp);
    if (hi <= lo) continue;
    int cur; // segment of product lo: the largest s with pre[s] <= lo
    {
      int a = 0, b = nseg - 1;
      while (a < b) {
        const int mid = (a + b + 1) >> 1;
        if (q->pre[mid] <= lo) a = mid;
        else b = mid - 1;
      }
      cur = a;
    }
    for (unsigned q0 = lo; q0 < hi;) {
      // distance of the next 32 boundaries; >= 1 because the segments are non-empty and `cur` holds product q0
      const unsigned L = q->pre[cur + 1 + lane] - q0;
      // sub-steps of this step: all their boundaries must lie inside the 32 just read
      const unsigned L31 = __shfl_sync(0xFFFFFFFFu, L, 31);
      unsigned nsub = min((unsigned)kFlatSub, L31 >> 5);
      nsub = min(nsub, (hi - q0 + 31u) >> 5);
      const unsigned Lj = L >> 5, Lb = 1u << (L & 31u);
      int row[kFlatSub];
      aval_t val[kFlatSub];
      mult_t mu[kFlatSub];
      int seg0 = cur;
#pragma unroll
      for (int j = 0; j < kFlatSub; ++j) {
        row[j] = -1;
        if ((unsigned)j < nsub) { // warp-uniform
          const unsigned M = __reduce_or_sync(0xFFFFFFFFu, Lj == (unsigned)j ? Lb : 0u);
          const int seg = seg0 + __popc(M & le_mask);
          seg0 += __popc(M);
          const unsigned qq = q0 + 32u * j + (unsigned)lane;
          if (qq < hi) {
            row[j] = *reinterpret_cast<const int32_t *>(q->rowp[seg] + 4ll * qq);
            if (NEED_VAL) {
              val[j] = *reinterpret_cast<const aval_t *>(q->valp[seg] + (long long)sizeof(aval_t) * qq);
              mu[j] = bits_mult<mult_t>(q->mult[seg]);
            }
          }
        }
      }
      // a boundary exactly at the first product of the next step moves the cursor once more
      const unsigned span = nsub << 5;
      cur = seg0 + (__any_sync(0xFFFFFFFFu, L == span) ? 1 : 0);
      q0 += span;
#pragma unroll
      for (int j = 0; j < kFlatSub; ++j)
        if (row[j] >= 0) use(row[j], val[j], mu[j]);
    }
  }
  __syncthreads();
}

// The multiply's walk: the same scheme in units of GROUPS of 4 products. Pieces of the window-major copy start at
// multiples of 4 entries and are padded with row -1 (util.cu), so a lane takes one group per sub-step with a 16-byte load
// of row ids and aligned vector loads of values, and one segment lookup serves four products.
template <class T>
struct alignas(sizeof(T) * 4 > 16 ? 16 : sizeof(T) * 4) Quad {
  T v[4];
};
constexpr int kFlatSub4 = 2; // sub-steps of 32 groups (128 products) per boundary-window load

template <class SR, bool NEED_VAL, int CH, class USE>
__device__ __forceinline__ void flat_walk4(const Source<SR, false> &s, const Task &k, FlatQueueT<CH> *q, USE &&use,
                                           bool stage_vals = false, bool reuse = false) {
  // stage_vals: also stage value pointers and multipliers although this walk reads rows only, so that the next walk of
  // the same single-chunk task can skip the staging (reuse): small tasks are bound by their chain of dependent loads
  typedef typename SR::b_t mult_t;
  typedef typename SR::a_t aval_t;
  constexpr int nwarp = CH >> 5; // blockDim.x == CH
  constexpr unsigned long long kLenMask = (1ull << 40) - 1ull;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned le_mask = 0xFFFFFFFFu >> (31 - lane); // bits 0..lane
  const int64_t *__restrict__ Tw = s.T2 + (int64_t)k.wlo * s.N;
  for (int64_t cbase = k.seg_begin; cbase < k.seg_end; cbase += CH) {
    int nseg;
    unsigned T;
    if (!reuse) {
    __syncthreads(); // previous chunk fully consumed
    int64_t beg = 0;
    unsigned ng = 0; // groups of this thread's segment
    mult_t mult = mult_t();
    if (cbase + threadIdx.x < k.seg_end) {
      const int64_t p = cbase + threadIdx.x;
      const int64_t *t = Tw + s.Bir[p];
      beg = t[0] & ~(int64_t)3;
      ng = (unsigned)(((t[1] & ~(int64_t)3) - beg) >> 2);
      if (NEED_VAL || stage_vals) mult = s.Bval[p];
    }
    const unsigned long long x = (ng > 0 ? (1ull << 40) : 0ull) | (unsigned long long)ng;
    unsigned long long incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) q->warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long wv = lane < nwarp ? q->warp_tot[lane] : 0ull, winc = wv;
#pragma unroll
    for (int d = 1; d < nwarp; d <<= 1) {
      unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, winc, d);
      if (lane >= d) winc += v;
    }
    const unsigned long long total = __shfl_sync(0xFFFFFFFFu, winc, nwarp - 1);
    const unsigned long long excl = __shfl_sync(0xFFFFFFFFu, winc - wv, warp) + incl - x;
    nseg = (int)(total >> 40);
    T = (unsigned)(total & kLenMask); // groups of the chunk
    if (ng > 0) {
      const int idx = (int)(excl >> 40);
      const long long pre = (long long)(excl & kLenMask);
      q->pre[idx] = (unsigned)pre;
      q->rowp[idx] = (long long)(s.Wir + beg) - 16ll * pre;
      if (NEED_VAL || stage_vals) {
        q->valp[idx] = (long long)(s.Wval + beg) - 4ll * (long long)sizeof(aval_t) * pre;
        q->mult[idx] = mult_bits<mult_t>(mult);
      }
    }
    if (threadIdx.x < 34) q->pre[nseg + threadIdx.x] = threadIdx.x == 0 ? T : 0xFFFFFFFFu;
    if (threadIdx.x == 0) {
      q->total = T;
      q->nseg = nseg;
    }
    __syncthreads();
    } else {
      nseg = q->nseg;
      T = q->total;
    }
    if (T == 0) continue;
    const unsigned lo = (unsigned)((unsigned long long)T * (unsigned)warp / (unsigned)nwarp);
    const unsigned hi = (unsigned)((unsigned long long)T * (unsigned)(warp + 1) / (unsigned)nwarp);
    if (hi <= lo) continue;
    int cur; // segment of group lo: the largest s with pre[s] <= lo
    {
      int a = 0, b = nseg - 1;
      while (a < b) {
        const int mid = (a + b + 1) >> 1;
        if (q->pre[mid] <= lo) a = mid;
        else b = mid - 1;
      }
      cur = a;
    }
    for (unsigned q0 = lo; q0 < hi;) {
      const unsigned L = q->pre[cur + 1 + lane] - q0; // >= 1
      const unsigned L31 = __shfl_sync(0xFFFFFFFFu, L, 31);
      unsigned nsub = min((unsigned)kFlatSub4, L31 >> 5);
      nsub = min(nsub, (hi - q0 + 31u) >> 5);
      const unsigned Lj = L >> 5, Lb = 1u << (L & 31u);
      int4 row[kFlatSub4];
      Quad<aval_t> val[kFlatSub4];
      mult_t mu[kFlatSub4];
      int seg0 = cur;
#pragma unroll
      for (int j = 0; j < kFlatSub4; ++j) {
        row[j] = make_int4(-1, -1, -1, -1);
        if ((unsigned)j < nsub) { // warp-uniform
          const unsigned M = __reduce_or_sync(0xFFFFFFFFu, Lj == (unsigned)j ? Lb : 0u);
          const int seg = seg0 + __popc(M & le_mask);
          seg0 += __popc(M);
          const unsigned g = q0 + 32u * j + (unsigned)lane;
          if (g < hi) {
            row[j] = *reinterpret_cast<const int4 *>(q->rowp[seg] + 16ll * g);
            if (NEED_VAL) {
              val[j] = *reinterpret_cast<const Quad<aval_t> *>(q->valp[seg] + 4ll * (long long)sizeof(aval_t) * g);
              mu[j] = bits_mult<mult_t>(q->mult[seg]);
            }
          }
        }
      }
      const unsigned span = nsub << 5;
      cur = seg0 + (__any_sync(0xFFFFFFFFu, L == span) ? 1 : 0);
      q0 += span;
#pragma unroll
      for (int j = 0; j < kFlatSub4; ++j) {
        if (row[j].x >= 0) use(row[j].x, val[j].v[0], mu[j]);
        if (row[j].y >= 0) use(row[j].y, val[j].v[1], mu[j]);
        if (row[j].z >= 0) use(row[j].z, val[j].v[2], mu[j]);
        if (row[j].w >= 0) use(row[j].w, val[j].v[3], mu[j]);
      }
    }
  }
  __syncthreads();
}

// the walk of the bitmap kernels: groups of 4 from the padded window-major copy (multiply), single products (merge)
template <class SR, bool MERGE, bool NEED_VAL, int CH, class USE>
__device__ __forceinline__ void bitmap_walk(const Source<SR, MERGE> &s, const Task &k, FlatQueueT<CH> *q, USE &&use,
                                            bool stage_vals = false, bool reuse = false) {
  if constexpr (MERGE) flat_walk<SR, MERGE, NEED_VAL>(s, k, q, use);
  else flat_walk4<SR, NEED_VAL>(s, k, q, use, stage_vals, reuse);
}

// ------------------------------------------------------------------------------------------------ K1: products per task
template <class SR, bool MERGE>
__global__ void __launch_bounds__(256) task_flop_kernel(Source<SR, MERGE> s, int64_t ntask, int64_t *flop) {
  int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= ntask) return; // whole warp exits together
  Task k = load_task(s, (int)t);
  task_segments(s, k);
  int64_t sum = 0;
  typename SR::b_t dummy;
  for (int64_t p = k.seg_begin + lane_id(); p < k.seg_end; p += 32) {
    int64_t beg;
    int len;
    load_segment<SR, MERGE, false>(s, k, p, beg, len, dummy);
    sum += len;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
  if (lane_id() == 0) flop[t] = sum;
}

// ------------------------------------------------------------------------------------------------ hash tables
template <int LOG2T>
__device__ __forceinline__ int hash_row(unsigned row) { return (int)((row * 2654435761u) >> (32 - LOG2T)); }

// returns the slot of `row`; `fresh` is set when this call inserted it
template <int LOG2T>
__device__ __forceinline__ int table_insert(unsigned *keys, unsigned row, bool &fresh) {
  constexpr int MASK = (1 << LOG2T) - 1;
  int h = hash_row<LOG2T>(row);
  fresh = false;
  while (true) {
    unsigned cur = ((volatile unsigned *)keys)[h];
    if (cur == row) return h;
    if (cur == kEmptyKey) {
      unsigned old = atomicCAS(&keys[h], kEmptyKey, row);
      if (old == kEmptyKey) {
        fresh = true;
        return h;
      }
      if (old == row) return h;
    }
    h = (h + 1) & MASK;
  }
}

template <int GROUP_WARPS>
__device__ __forceinline__ void group_sync() {
  if (GROUP_WARPS == 1) __syncwarp();
  else __syncthreads();
}

// K2 (hash): distinct rows per task. GROUP_WARPS == 1: one task per warp, 8 tasks per CTA;
// otherwise one task per CTA of GROUP_WARPS warps.
template <class SR, bool MERGE, int GROUP_WARPS, int LOG2T>
__global__ void __launch_bounds__(GROUP_WARPS == 1 ? 256 : GROUP_WARPS * 32)
sym_hash_kernel(Source<SR, MERGE> s, const int32_t *order, int64_t count, int64_t *tasknnz) {
  constexpr int T = 1 << LOG2T;
  constexpr int GROUPS = GROUP_WARPS == 1 ? 8 : 1;
  constexpr int GT = GROUP_WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned *keys_all = reinterpret_cast<unsigned *>(smem_raw);
  __shared__ typename std::conditional<GROUP_WARPS == 1, NoQueue, CtaQueueT<GROUP_WARPS * 32>>::type queue;
  __shared__ int cta_count;
  const int group = GROUP_WARPS == 1 ? (threadIdx.x >> 5) : 0;
  const int gtid = GROUP_WARPS == 1 ? lane_id() : threadIdx.x;
  int64_t ti = (int64_t)blockIdx.x * GROUPS + group;
  const bool valid = ti < count; // uniform per group
  if (GROUP_WARPS == 1 && !valid) return;
  unsigned *keys = keys_all + group * T;
  for (int i = gtid; i < T; i += GT) keys[i] = kEmptyKey;
  if (GROUP_WARPS > 1 && threadIdx.x == 0) cta_count = 0;
  group_sync<GROUP_WARPS>();
  int t = order[ti];
  Task k = load_task(s, t);
  task_segments(s, k);
  int mine = 0;
  auto ld = [&](int64_t pos) { return k.rows[pos]; };
  auto use = [&](int row, typename SR::b_t) {
    bool fresh;
    table_insert<LOG2T>(keys, (unsigned)row, fresh);
    mine += fresh ? 1 : 0;
  };
  if (GROUP_WARPS == 1) warp_walk<SR, MERGE, false>(s, k, ld, use);
  else cta_walk<SR, MERGE, false>(s, k, reinterpret_cast<CtaQueueT<GROUP_WARPS * 32> *>(&queue), ld, use);
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
  if (GROUP_WARPS == 1) {
    if (lane_id() == 0) tasknnz[t] = mine;
  } else {
    if (lane_id() == 0 && mine) atomicAdd(&cta_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) tasknnz[t] = cta_count;
  }
}

// bitonic sort of P (power of two) 64-bit keys in shared memory by one group
template <int GROUP_WARPS>
__device__ __forceinline__ void group_bitonic(unsigned long long *a, int P, int gtid, int GT) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = gtid; i < P; i += GT) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long x = a[i], y = a[ixj];
          bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      group_sync<GROUP_WARPS>();
    }
  }
}

// K4 (hash): accumulate, compact, sort by row, emit.
template <class SR, bool MERGE, int GROUP_WARPS, int LOG2T>
__global__ void __launch_bounds__(GROUP_WARPS == 1 ? 256 : GROUP_WARPS * 32, GROUP_WARPS == 1 ? 6 : 1)
num_hash_kernel(Source<SR, MERGE> s, const int32_t *order, int64_t count, const int64_t *taskptr, int32_t *Cir,
                typename SR::out_t *Cval) {
  typedef typename SR::acc_t acc_t;
  constexpr int T = 1 << LOG2T;
  constexpr int GROUPS = GROUP_WARPS == 1 ? 8 : 1;
  constexpr int GT = GROUP_WARPS * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout per CTA: sortbuf[GROUPS*T/2] u64 | acc[GROUPS*T] acc_t | keys[GROUPS*T] u32 | counters[GROUPS]
  // (a task of this class has at most T/2 outputs, so the sort buffer needs T/2 entries)
  unsigned long long *sort_all = reinterpret_cast<unsigned long long *>(smem_raw);
  acc_t *acc_all = reinterpret_cast<acc_t *>(sort_all + GROUPS * (T / 2));
  unsigned *keys_all = reinterpret_cast<unsigned *>(acc_all + GROUPS * T);
  int *cnt_all = reinterpret_cast<int *>(keys_all + GROUPS * T);
  __shared__ typename std::conditional<GROUP_WARPS == 1, NoQueue, CtaQueueT<GROUP_WARPS * 32>>::type queue;
  const int group = GROUP_WARPS == 1 ? (threadIdx.x >> 5) : 0;
  const int gtid = GROUP_WARPS == 1 ? lane_id() : threadIdx.x;
  int64_t ti = (int64_t)blockIdx.x * GROUPS + group;
  if (GROUP_WARPS == 1 && ti >= count) return;
  unsigned long long *sortbuf = sort_all + group * (T / 2);
  acc_t *acc = acc_all + group * T;
  unsigned *keys = keys_all + group * T;
  int *cnt = cnt_all + group;
  for (int i = gtid; i < T; i += GT) {
    keys[i] = kEmptyKey;
    acc[i] = SR::identity();
  }
  if (gtid == 0) *cnt = 0;
  group_sync<GROUP_WARPS>();
  int t = order[ti];
  Task k = load_task(s, t);
  task_segments(s, k);
  typedef typename Source<SR, MERGE>::aval_t aval_t;
  auto ld = [&](int64_t pos) { return RowVal<aval_t>{k.rows[pos], ((const aval_t *)k.vals)[pos]}; };
  auto use = [&](const RowVal<aval_t> &x, typename SR::b_t mu) {
    bool fresh;
    int slot = table_insert<LOG2T>(keys, (unsigned)x.row, fresh);
    acc_t v;
    if (MERGE) v = SR::from_out((typename SR::out_t)x.val);
    else v = SR::mul((typename SR::a_t)x.val, mu);
    SR::accumulate(&acc[slot], v);
  };
  if (GROUP_WARPS == 1) {
    warp_walk<SR, MERGE, true>(s, k, ld, use);
    __syncwarp();
  } else {
    cta_walk<SR, MERGE, true>(s, k, reinterpret_cast<CtaQueueT<GROUP_WARPS * 32> *>(&queue), ld, use);
  }
  // compact the occupied slots (one warp: positions by ballot; a CTA: one shared counter bump per warp and round)
  for (int i0 = 0; i0 < T; i0 += GT) {
    const int i = i0 + gtid; // T is a multiple of GT
    const unsigned key = keys[i];
    const bool hit = key != kEmptyKey;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
    int base = 0;
    if (GROUP_WARPS == 1) {
      base = *cnt;
      __syncwarp();
      if (lane_id() == 0) *cnt = base + __popc(m);
      __syncwarp();
    } else {
      if (lane_id() == 0 && m) base = atomicAdd(cnt, __popc(m));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
    }
    if (hit) sortbuf[base + __popc(m & ((1u << lane_id()) - 1u))] = ((unsigned long long)key << 32) | (unsigned)i;
  }
  group_sync<GROUP_WARPS>();
  const int n = *cnt;
  const int64_t obase = taskptr[t];
  int P = 2;
  while (P < n) P <<= 1;
  for (int i = n + gtid; i < P; i += GT) sortbuf[i] = ~0ull;
  group_sync<GROUP_WARPS>();
  group_bitonic<GROUP_WARPS>(sortbuf, P, gtid, GT);
  for (int i = gtid; i < n; i += GT) {
    unsigned long long e = sortbuf[i];
    Cir[obase + i] = (int32_t)(e >> 32);
    Cval[obase + i] = SR::to_out(acc[(unsigned)e]);
  }
}

// ------------------------------------------------------------------------------------------------ register-sort path
// Tasks with at most 8 G products AND at most 8 G segments, G = 8 / 16 / 32 lanes per task (4 / 2 / 1 tasks per warp). These are
// the columns the reference's kernel picks its heap for (compression close to 1, mtSpGEMM.h:276-279) and all of an
// Erdos-Renyi product (d^2 = 64 products per column). A hash table + compaction + sort in shared memory cost them about
// 2500 warp instructions per task (a 64-product task!); here the products of a task go, 8 per lane, through a bitonic
// network held in REGISTERS: exchanges below distance 8 are register to register, the others one shuffle inside the
// group; equal rows are then neighbours and are folded by a sweep inside the lane plus a segmented scan over the lanes of
// the group. No table, no probing, no barrier; shared memory only hands the products from the lane that loaded a
// segment to the lane that sorts them. The symbolic variant sorts the rows alone and counts the distinct ones.
constexpr int kRsItems = 8;
constexpr unsigned kRsPad = 0xFFFFFFFFu; // sorts behind every row id

template <class T>
__device__ __forceinline__ T rs_shfl_xor(unsigned mask, T v, int d, int width) {
  if constexpr (sizeof(T) == 8) {
    unsigned long long u;
    memcpy(&u, &v, 8);
    unsigned lo = __shfl_xor_sync(mask, (unsigned)u, d, width), hi = __shfl_xor_sync(mask, (unsigned)(u >> 32), d, width);
    u = ((unsigned long long)hi << 32) | lo;
    memcpy(&v, &u, 8);
    return v;
  } else {
    unsigned u;
    memcpy(&u, &v, 4);
    u = __shfl_xor_sync(mask, u, d, width);
    memcpy(&v, &u, 4);
    return v;
  }
}
template <class T>
__device__ __forceinline__ T rs_shfl_up(unsigned mask, T v, int d, int width) {
  if constexpr (sizeof(T) == 8) {
    unsigned long long u;
    memcpy(&u, &v, 8);
    unsigned lo = __shfl_up_sync(mask, (unsigned)u, d, width), hi = __shfl_up_sync(mask, (unsigned)(u >> 32), d, width);
    u = ((unsigned long long)hi << 32) | lo;
    memcpy(&v, &u, 8);
    return v;
  } else {
    unsigned u;
    memcpy(&u, &v, 4);
    u = __shfl_up_sync(mask, u, d, width);
    memcpy(&v, &u, 4);
    return v;
  }
}

// PACKED (numeric pass, blocks of fewer than 2^(32 - log2 CAP) rows): the network sorts row << log2(CAP) | staging position
// -- one register per element instead of key + value -- and the values are fetched from the staging area afterwards.
template <class SR, bool MERGE, int G, bool NUMERIC, bool PACKED = false>
__global__ void __launch_bounds__(256, NUMERIC ? 4 : 6)
regsort_kernel(Source<SR, MERGE> s, const TaskRec *recs, int64_t count, int64_t *tasknnz, int32_t *Cir, typename SR::out_t *Cval) {
  typedef typename SR::acc_t acc_t;
  typedef typename Source<SR, MERGE>::aval_t aval_t;
  typedef typename SR::b_t mult_t;
  constexpr int E = kRsItems, CAP = G * E, GROUPS = 256 / G;
  constexpr int LOG2CAP = G == 8 ? 6 : (G == 16 ? 7 : 8);
  constexpr bool SORT_VALS = NUMERIC && !PACKED; // values travel through the network
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned *skey_all = reinterpret_cast<unsigned *>(smem_raw);             // [GROUPS * CAP] = 2048 rows
  acc_t *sval_all = reinterpret_cast<acc_t *>(skey_all + GROUPS * CAP);    // [GROUPS * CAP] (numeric only)
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), group = threadIdx.x / G;
  const unsigned gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const int64_t ti = (int64_t)blockIdx.x * GROUPS + group;
  if (ti >= count) return; // whole group; every collective below names the lanes of the group only
  unsigned *skey = skey_all + group * CAP;
  acc_t *sval = sval_all + group * CAP;
  const TaskRec r = recs[ti]; // launch-order record: segment range, window, output offset (symbolic: the task id) in one load
  const Task k = task_from_record(s, r);
  const int nseg = (int)(k.seg_end - k.seg_begin); // <= CAP by classification, and so is the number of products
  // ---- load: lane gl takes segments gl, gl + G, ...; their products land in segment order in the group's staging area
  int base = 0;
  for (int s0 = 0; s0 < nseg; s0 += G) { // uniform per group
    const int si = s0 + gl;
    int64_t beg = 0;
    int len = 0;
    mult_t mult = mult_t();
    if (si < nseg) load_segment<SR, MERGE, NUMERIC>(s, k, k.seg_begin + si, beg, len, mult);
    int incl = len;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const int v = __shfl_up_sync(gmask, incl, d, G);
      if (gl >= d) incl += v;
    }
    const int pos = base + incl - len;
    const int32_t *__restrict__ rp = k.rows + beg;
    const aval_t *__restrict__ vp = reinterpret_cast<const aval_t *>(k.vals) + beg;
    for (int i0 = 0; i0 < len; i0 += 4) { // four loads in flight before the first store (the stores may alias for all the compiler knows)
      int r[4];
      aval_t a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u < len) {
          r[u] = rp[i0 + u];
          if (NUMERIC) a[u] = vp[i0 + u];
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u < len) {
          skey[pos + i0 + u] = (unsigned)r[u];
          if (NUMERIC) sval[pos + i0 + u] = MERGE ? SR::from_out((typename SR::out_t)a[u]) : SR::mul((typename SR::a_t)a[u], mult);
        }
    }
    base += __shfl_sync(gmask, incl, G - 1, G);
  }
  const int P = base;
  __syncwarp(gmask);
  // ---- sort: element e = gl * E + slot, ascending by row; pads behind
  unsigned key[E];
  acc_t val[E];
#pragma unroll
  for (int q = 0; q < E; ++q) {
    const int e = gl * E + q;
    key[q] = e < P ? (PACKED ? ((skey[e] << LOG2CAP) | (unsigned)e) : skey[e]) : kRsPad;
    if (SORT_VALS) val[q] = e < P ? sval[e] : SR::identity();
  }
#pragma unroll
  for (int kk = 2; kk <= CAP; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= E) { // partner in lane gl ^ (j / E), same slot
        const int lj = j / E;
        const bool up = ((gl * E) & kk) == 0;
        const bool keep_min = ((gl & lj) == 0) == up;
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const unsigned ok = __shfl_xor_sync(gmask, key[q], lj, G);
          acc_t ov;
          if (SORT_VALS) ov = rs_shfl_xor<acc_t>(gmask, val[q], lj, G);
          const bool take = keep_min ? (ok < key[q]) : (ok > key[q]);
          if (take) {
            key[q] = ok;
            if (SORT_VALS) val[q] = ov;
          }
        }
      } else { // partner in the same lane
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int p = q ^ j;
          if (p > q) {
            const bool up = ((gl * E + q) & kk) == 0;
            if ((key[q] > key[p]) == up) {
              const unsigned tk = key[q];
              key[q] = key[p];
              key[p] = tk;
              if (SORT_VALS) {
                const acc_t tv = val[q];
                val[q] = val[p];
                val[p] = tv;
              }
            }
          }
        }
      }
    }
  }
  if (PACKED) { // rows back in place, values from the staging area by position
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const unsigned kq = key[q];
      if (NUMERIC) val[q] = kq != kRsPad ? sval[kq & (unsigned)(CAP - 1)] : SR::identity();
      key[q] = kq != kRsPad ? (kq >> LOG2CAP) : kRsPad;
    }
  }
  // ---- fold equal rows
  unsigned prev_last = __shfl_up_sync(gmask, key[E - 1], 1, G);  // last row of the lane before
  unsigned next_first = __shfl_down_sync(gmask, key[0], 1, G);   // first row of the lane after
  const bool from_prev = gl > 0 && prev_last == key[0];          // my first run started in an earlier lane
  const bool into_next = gl < G - 1 && next_first == key[E - 1]; // my last run goes on in the next lane
  if (!NUMERIC) {
    int heads = 0;
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const bool head = key[q] != kRsPad && (q == 0 ? !from_prev : key[q] != key[q - 1]);
      heads += head ? 1 : 0;
    }
#pragma unroll
    for (int d = G >> 1; d >= 1; d >>= 1) heads += __shfl_xor_sync(gmask, heads, d, G);
    if (gl == 0) tasknnz[r.obase] = heads;
    return;
  } else {
    // inside the lane: the last element of every run collects the run
#pragma unroll
    for (int q = 1; q < E; ++q)
      if (key[q] == key[q - 1]) val[q] = SR::acc_add(val[q - 1], val[q]);
    // across lanes: C(l) = what the run that ends lane l has collected up to the end of lane l. A lane chains to the left only
    // if it is one single run that started before it; segmented inclusive scan over the lanes of the group.
    acc_t sum = val[E - 1];
    bool open = from_prev && key[0] == key[E - 1];
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const acc_t os = rs_shfl_up<acc_t>(gmask, sum, d, G);
      const int oo = __shfl_up_sync(gmask, (int)open, d, G);
      if (gl >= d && open) {
        sum = SR::acc_add(os, sum);
        open = oo != 0;
      }
    }
    const acc_t carry = rs_shfl_up<acc_t>(gmask, sum, 1, G); // C(l - 1): goes into my first run when it started earlier
    bool pending = from_prev;
    int tails = 0;
    bool tail[E];
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const bool lane_tail = (q == E - 1) || key[q] != key[q + 1];
      if (pending && lane_tail) {
        val[q] = SR::acc_add(carry, val[q]);
        pending = false;
      }
      tail[q] = lane_tail && key[q] != kRsPad && !(q == E - 1 && into_next);
      tails += tail[q] ? 1 : 0;
    }
    int incl = tails;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const int v = __shfl_up_sync(gmask, incl, d, G);
      if (gl >= d) incl += v;
    }
    int64_t o = r.obase + (incl - tails);
#pragma unroll
    for (int q = 0; q < E; ++q)
      if (tail[q]) {
        Cir[o] = (int32_t)key[q];
        Cval[o] = SR::to_out(val[q]);
        ++o;
      }
  }
}

// ------------------------------------------------------------------------------------------------ bitmap path
// The presence bitmap of a row window is an array of 32-row WORDS in shared memory; after the scan a second array holds
// for every word the number of present rows in all earlier words, so the output slot of a product is
//     rank(row) = rank[w] + popc(bits[w] & ((1 << b) - 1)),  w = (row - rbase) >> 5, b = (row - rbase) & 31:
// no sort, no probing. The symbolic pass can hand the words of a task to the numeric pass through HBM (4 bytes per 32
// rows), which then only repeats the scan.
__host__ __device__ __forceinline__ int words_of_rows(int64_t rows) { return (int)((rows + 31) >> 5); }

// block-wide exclusive scan of one int per thread (blockDim <= 1024); returns exclusive prefix, total in *total
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_sums /*[32]*/, int *total) {
  const int lane = lane_id(), warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += x;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int x = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (lane >= d) wi += x;
    }
    warp_sums[lane] = wi - w; // exclusive warp offsets
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  return warp_sums[warp] + incl - v;
}

// the row window of a task
struct Window {
  int rbase, wrows, nword;
};
template <class Src>
__device__ __forceinline__ Window task_window(const Src &s, const Task &k, int64_t m) {
  const int64_t rbase64 = (int64_t)k.wlo << s.wlog2;
  const int64_t rend = (k.whi == s.nwin) ? m : ((int64_t)k.whi << s.wlog2);
  Window w;
  w.rbase = (int)rbase64;
  w.wrows = (int)(rend - rbase64);
  w.nword = words_of_rows(w.wrows);
  return w;
}

// clear the words, then set the bit of every row that occurs in the task
template <class SR, bool MERGE, int THREADS>
__device__ __forceinline__ void bitmap_mark(const Source<SR, MERGE> &s, const Task &k, FlatQueueT<THREADS> *q, unsigned *bits,
                                            int nword, int rbase, bool stage_vals = false) {
  uint4 *c4 = reinterpret_cast<uint4 *>(bits);
  const int nvec = (nword + 3) >> 2; // the array is padded to a multiple of 4 words
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) c4[i] = make_uint4(0, 0, 0, 0);
  typedef typename Source<SR, MERGE>::aval_t aval_t;
  auto use = [&](int row, aval_t, typename SR::b_t) {
    const unsigned r = (unsigned)(row - rbase);
    atomicOr(&bits[r >> 5], 1u << (r & 31u));
  };
  bitmap_walk<SR, MERGE, false>(s, k, q, use, stage_vals); // starts and ends with __syncthreads
}

// number of present rows; with RANKS the exclusive prefix of every word goes to rank[]. Ends with __syncthreads.
// warp_sums has 33 entries.
template <bool RANKS>
__device__ __forceinline__ int bitmap_scan(const unsigned *bits, unsigned *rank, int nword, int *warp_sums) {
  if (!RANKS) { // count only: 16-byte loads, thread t takes vectors t, t + blockDim.x, ... (the array is padded to whole vectors)
    const uint4 *b4 = reinterpret_cast<const uint4 *>(bits);
    const int nvec = (nword + 3) >> 2;
    int mine = 0;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      const uint4 w = b4[v];
      mine += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
    block_exclusive_scan(mine, warp_sums, warp_sums + 32);
    const int total = warp_sums[32];
    __syncthreads();
    return total;
  }
  // every thread owns a contiguous run of words; an ODD run length keeps the lanes of a warp on different banks
  const int wpt = ((nword + blockDim.x - 1) / blockDim.x) | 1;
  const int c0 = min(nword, (int)threadIdx.x * wpt), c1 = min(nword, c0 + wpt);
  int mine = 0;
  for (int c = c0; c < c1; ++c) mine += __popc(bits[c]);
  int run = block_exclusive_scan(mine, warp_sums, warp_sums + 32);
  const int total = warp_sums[32];
  if (RANKS) {
    for (int c = c0; c < c1; ++c) {
      rank[c] = (unsigned)run;
      run += __popc(bits[c]);
    }
  }
  __syncthreads(); // ranks final; warp_sums free for the next use
  return total;
}

// K2 (bitmap): rows of the window present in the task. Tasks whose record carries a hand-over slot also store their
// words at saved + slot * save_stride for the numeric pass (slot_of_task[t] = slot).
template <class SR, bool MERGE, int THREADS>
__global__ void __launch_bounds__(THREADS)
sym_bitmap_kernel(Source<SR, MERGE> s, const TaskRec *recs, int64_t count, int64_t m, int64_t *tasknnz, unsigned *saved,
                  int64_t save_stride, int save_count, int32_t *slot_of_task) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned *bits = reinterpret_cast<unsigned *>(smem_raw);
  __shared__ FlatQueueT<THREADS> queue;
  __shared__ int warp_sums[33];
  const TaskRec r = recs[blockIdx.x];
  const int t = (int)r.obase;
  const Task k = task_from_record(s, r);
  const Window w = task_window(s, k, m);
  bitmap_mark(s, k, &queue, bits, w.nword, w.rbase);
  const int nnz = bitmap_scan<false>(bits, nullptr, w.nword, warp_sums);
  if (threadIdx.x == 0) {
    tasknnz[t] = nnz;
    if (r.slot >= 0) { // the words leave as one bulk store of the TMA unit (the scan ended with a barrier: they are final)
      fence_proxy_async();
      bulk_store(saved + (int64_t)r.slot * save_stride, bits, (unsigned)((w.nword + 3) >> 2) * 16u);
      bulk_commit();
      slot_of_task[t] = r.slot;
      bulk_wait_read(); // shared memory must outlive the read; the write completes before the kernel does
    }
  }
}

// ranked words of a task in shared memory: taken from the symbolic pass if it stored them, marked again otherwise
template <class SR, bool MERGE, int THREADS>
__device__ __forceinline__ void bitmap_obtain(const Source<SR, MERGE> &s, const Task &k, FlatQueueT<THREADS> *q, unsigned *bits,
                                              unsigned *rank, const Window &w, int *warp_sums, const unsigned *saved,
                                              int64_t save_stride, int slot, bool stage_vals = false) {
  if (slot >= 0) {
    const uint4 *src = reinterpret_cast<const uint4 *>(saved + (int64_t)slot * save_stride);
    uint4 *dst = reinterpret_cast<uint4 *>(bits);
    const int nvec = (w.nword + 3) >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  } else {
    bitmap_mark(s, k, q, bits, w.nword, w.rbase, stage_vals);
  }
  bitmap_scan<true>(bits, rank, w.nword, warp_sums);
}

// K4 (bitmap, accumulators in C itself): the fallback for tasks whose outputs do not fit shared memory. Every product is
// one RED into the task's (L2-resident) slice of C's value array.
template <class SR, bool MERGE, int THREADS, int MINB = (THREADS == 512 ? 2 : 4)>
__global__ void __launch_bounds__(THREADS, MINB)
num_bitmap_kernel(Source<SR, MERGE> s, const TaskRec *recs, int64_t count, int64_t m, int max_words, int32_t *Cir,
                  typename SR::out_t *Cval, const unsigned *saved, int64_t save_stride) {
  typedef typename SR::acc_t acc_t;
  typedef typename SR::out_t out_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned *bits = reinterpret_cast<unsigned *>(smem_raw);
  unsigned *rank = bits + max_words;
  __shared__ FlatQueueT<THREADS> queue;
  __shared__ int warp_sums[33];
  const TaskRec r = recs[blockIdx.x];
  const Task k = task_from_record(s, r);
  const Window w = task_window(s, k, m);
  const int rbase = w.rbase;
  const int64_t obase = r.obase;
  const int nnz = r.nnz;
  bitmap_obtain(s, k, &queue, bits, rank, w, warp_sums, saved, save_stride, r.slot);
  // rows: every thread unpacks words c, c + THREADS, ... (neighbouring lanes write neighbouring pieces of Cir)
  for (int c = threadIdx.x; c < w.nword; c += THREADS) {
    unsigned b = bits[c];
    int32_t *o = Cir + obase + rank[c];
    const int rowbase = rbase + (c << 5);
    while (b) {
      *o++ = rowbase + __ffs(b) - 1;
      b &= b - 1;
    }
  }
  for (int i = threadIdx.x; i < nnz; i += THREADS) Cval[obase + i] = SR::to_out(SR::identity());
  typedef typename Source<SR, MERGE>::aval_t aval_t;
  auto use = [&](int row, aval_t aval, typename SR::b_t mu) {
    const unsigned r = (unsigned)(row - rbase);
    const unsigned wd = r >> 5;
    const unsigned slot = rank[wd] + (unsigned)__popc(bits[wd] & ((1u << (r & 31u)) - 1u));
    acc_t v;
    if (MERGE) v = SR::from_out((out_t)aval);
    else v = SR::mul((typename SR::a_t)aval, mu);
    SR::accumulate_out(&Cval[obase + slot], v);
  };
  bitmap_walk<SR, MERGE, true>(s, k, &queue, use); // starts with __syncthreads: the identities are in place
}

#ifdef CBGPU_PHASE_TIMING // tuning builds only (make EXTRA=-DCBGPU_PHASE_TIMING): cycles of thread 0 per phase and CTA shape
static __device__ unsigned long long g_phase_cycles[3][8];
#define CB_PHASE_BEGIN() long long cb_tprev = clock64()
#define CB_PHASE(idx)                                                                                                  \
  do {                                                                                                                 \
    if (threadIdx.x == 0) {                                                                                            \
      const long long cb_now = clock64();                                                                              \
      atomicAdd(&g_phase_cycles[THREADS == 1024 ? 0 : (THREADS == 512 ? 1 : 2)][idx], (unsigned long long)(cb_now - cb_tprev)); \
      cb_tprev = cb_now;                                                                                               \
    }                                                                                                                  \
  } while (0)
#else
#define CB_PHASE_BEGIN() do {} while (0)
#define CB_PHASE(idx) do {} while (0)
#endif

// K4 (bitmap, accumulators in shared memory): the numeric kernel of every task whose outputs fit the CTA's shared memory.
// Ranked words as above; the sorted rows are unpacked by rank into a staging area and leave with coalesced stores; every
// product then finds its slot with two shared loads and a popcount and is added to a shared-memory accumulator with the
// exchange protocol of semiring.cuh (no CAS loop, no L2 reduction); the values leave once, coalesced. Nothing of C is
// touched twice and nothing is pre-filled.
// FIRST: most outputs of the class receive a single product (compression close to 1).
template <class SR, bool MERGE, int THREADS, int MINB, bool FIRST>
__global__ void __launch_bounds__(THREADS, MINB)
num_sacc_kernel(Source<SR, MERGE> s, const TaskRec *recs, int64_t m, int max_words, int32_t *Cir, typename SR::out_t *Cval,
                const unsigned *saved, int64_t save_stride) {
  typedef typename SR::acc_t acc_t;
  typedef typename SR::out_t out_t;
  typedef typename Source<SR, MERGE>::aval_t aval_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: bits[max_words] u32 | rank[max_words] u32 | acc[outputs of the largest task of the class] acc_t
  unsigned *bits = reinterpret_cast<unsigned *>(smem_raw);
  unsigned *rank = bits + max_words;
  acc_t *acc = reinterpret_cast<acc_t *>(rank + max_words);
  __shared__ FlatQueueT<THREADS> queue;
  __shared__ int warp_sums[33];
  const TaskRec r = recs[blockIdx.x];
  const Task k = task_from_record(s, r);
  const Window w = task_window(s, k, m);
  const int rbase = w.rbase;
  const int64_t obase = r.obase;
  const int nnz = r.nnz;
  // a task whose segments fit one chunk stages them once for both of its walks (mark, accumulate)
  const bool restage = !MERGE && r.slot < 0 && r.nseg <= THREADS;
  CB_PHASE_BEGIN();
#ifdef CBGPU_PHASE_TIMING
  if (r.slot >= 0) {
    const uint4 *src = reinterpret_cast<const uint4 *>(saved + (int64_t)r.slot * save_stride);
    uint4 *dst = reinterpret_cast<uint4 *>(bits);
    const int nvec = (w.nword + 3) >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    CB_PHASE(0); // presence words from the symbolic pass
  } else {
    bitmap_mark(s, k, &queue, bits, w.nword, w.rbase, restage);
    CB_PHASE(1); // staging + mark walk
  }
  bitmap_scan<true>(bits, rank, w.nword, warp_sums);
  CB_PHASE(2); // scan + rank prefixes
#else
  bitmap_obtain(s, k, &queue, bits, rank, w, warp_sums, saved, save_stride, r.slot, restage);
#endif
  // rows: unpacked into the staging area (the accumulator array, not yet in use) at their ranks, then copied out
  int32_t *stage = reinterpret_cast<int32_t *>(acc);
  for (int c = threadIdx.x; c < w.nword; c += THREADS) {
    unsigned b = bits[c];
    int o = (int)rank[c];
    const int rowbase = rbase + (c << 5);
    while (b) {
      stage[o++] = rowbase + __ffs(b) - 1;
      b &= b - 1;
    }
  }
  __syncthreads();
  CB_PHASE(3); // rows unpacked
  for (int i = threadIdx.x; i < nnz; i += THREADS) Cir[obase + i] = stage[i];
  __syncthreads();
  for (int i = threadIdx.x; i < nnz; i += THREADS) acc[i] = SR::identity();
  auto use = [&](int row, aval_t aval, typename SR::b_t mu) {
    const unsigned r = (unsigned)(row - rbase);
    const unsigned wd = r >> 5;
    const unsigned slot = rank[wd] + (unsigned)__popc(bits[wd] & ((1u << (r & 31u)) - 1u));
    acc_t v;
    if (MERGE) v = SR::from_out((out_t)aval);
    else v = SR::mul((typename SR::a_t)aval, mu);
    SR::template accumulate_shared<FIRST>(&acc[slot], v);
  };
  __syncthreads(); // accumulators initialised (the walk itself only synchronises when it stages)
  CB_PHASE(4); // rows stored, accumulators initialised
  bitmap_walk<SR, MERGE, true>(s, k, &queue, use, false, restage); // ends with __syncthreads
  CB_PHASE(5); // accumulate walk
  for (int i = threadIdx.x; i < nnz; i += THREADS) Cval[obase + i] = SR::to_out(acc[i]);
  CB_PHASE(6); // values stored
#ifdef CBGPU_PHASE_TIMING
  if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[THREADS == 1024 ? 0 : (THREADS == 512 ? 1 : 2)][7], 1ull);
#endif
}

// ------------------------------------------------------------------------------------------------ shared accumulators, second version
// What the per-phase cycle counts of the kernel above showed (profiles/r2_phase_cycles.txt): a task of the small shape spends
// three quarters of its time in passes over the 4096 words of the window -- scan + rank 15 %, unpacking the sorted rows 24 %
// -- although six words in seven are empty. This version
//  * keeps the rank prefixes as 16-bit numbers (a task of these classes has < 65536 outputs): 8 KB less shared memory per task
//    at 2^17-row windows, which pays for
//  * a 16-bit row array next to the accumulators: every product stores its row offset at its slot (all products of a slot store
//    the same value), so the sorted rows fall out of the accumulate walk and the unpack pass over the words disappears
//    (ROWS_BY_WALK); in the medium shape that pass was 22 % of all instructions executed (divergent: a warp iterates as often as
//    its fullest word has bits; profiles/r2_ncu_sacc2_s22.txt);
//  * scans the words with 16-byte loads, thread t taking vectors t, t + THREADS, ...: the up to four per-thread counts travel
//    through ONE block scan as 16-bit fields of a 64-bit word, and the four ranks of a vector leave as one 8-byte store;
//  * fetches handed-over presence words with one bulk copy of the TMA unit instead of a load/store loop of the whole CTA.
template <int THREADS>
__device__ __forceinline__ void bitmap_scan16(const unsigned *bits, unsigned short *rank, int nword, unsigned long long *wtot /*[32]*/) {
  constexpr int nwarp = THREADS >> 5;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  constexpr int J = (1024 + THREADS - 1) / THREADS; // vectors per thread: windows of at most 4096 words (checked by the host)
  const int nvec = (nword + 3) >> 2;
  const uint4 *b4 = reinterpret_cast<const uint4 *>(bits);
  uint4 w[J];
  unsigned long long x = 0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int v = j * THREADS + (int)threadIdx.x;
    w[j] = make_uint4(0u, 0u, 0u, 0u);
    if (v < nvec) w[j] = b4[v];
    x |= (unsigned long long)(unsigned)(__popc(w[j].x) + __popc(w[j].y) + __popc(w[j].z) + __popc(w[j].w)) << (16 * j);
  }
  unsigned long long incl = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  unsigned long long wv = lane < nwarp ? wtot[lane] : 0ull, winc = wv;
#pragma unroll
  for (int d = 1; d < nwarp; d <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, winc, d);
    if (lane >= d) winc += v;
  }
  const unsigned long long total = __shfl_sync(0xFFFFFFFFu, winc, nwarp - 1);
  const unsigned long long excl = __shfl_sync(0xFFFFFFFFu, winc - wv, warp) + incl - x;
  unsigned base = 0; // outputs of the vector blocks before block j
  uint2 *r2 = reinterpret_cast<uint2 *>(rank);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int v = j * THREADS + (int)threadIdx.x;
    if (v < nvec) {
      const unsigned r0 = base + (unsigned)((excl >> (16 * j)) & 0xFFFFull);
      const unsigned r1 = r0 + (unsigned)__popc(w[j].x), r2v = r1 + (unsigned)__popc(w[j].y), r3 = r2v + (unsigned)__popc(w[j].z);
      r2[v] = make_uint2(r0 | (r1 << 16), r2v | (r3 << 16));
    }
    base += (unsigned)((total >> (16 * j)) & 0xFFFFull);
  }
  __syncthreads();
}

template <class SR, bool MERGE, int THREADS, int MINB, bool FIRST, bool ROWS_BY_WALK>
__global__ void __launch_bounds__(THREADS, MINB)
num_sacc2_kernel(Source<SR, MERGE> s, const TaskRec *recs, int64_t m, int max_words, int cap, int32_t *Cir, typename SR::out_t *Cval,
                 const unsigned *saved, int64_t save_stride) {
  typedef typename SR::acc_t acc_t;
  typedef typename SR::out_t out_t;
  typedef typename Source<SR, MERGE>::aval_t aval_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: bits[max_words] u32 | acc[cap] acc_t | rows[cap] u16 (ROWS_BY_WALK) | rank[max_words] u16; cap is a multiple of 4.
  // A row is kept as the low 16 bits of its offset in the window (at most 2^17 rows, checked by the host): outputs are sorted by
  // row, so the 17th bit of output i is simply i >= (number of outputs in the lower 2^16 rows) = rank[2048].
  unsigned *bits = reinterpret_cast<unsigned *>(smem_raw);
  acc_t *acc = reinterpret_cast<acc_t *>(bits + max_words);
  unsigned short *srow = reinterpret_cast<unsigned short *>(acc + cap);
  unsigned short *rank = ROWS_BY_WALK ? srow + cap : srow;
  __shared__ FlatQueueT<THREADS> queue;
  __shared__ unsigned long long wtot[32];
  __shared__ __align__(8) unsigned long long bar;
  const TaskRec r = recs[blockIdx.x];
  const Task k = task_from_record(s, r);
  const Window w = task_window(s, k, m);
  const int rbase = w.rbase;
  const int64_t obase = r.obase;
  const int nnz = r.nnz;
  const bool restage = !MERGE && r.slot < 0 && r.nseg <= THREADS;
  const int first = min(nnz, cap); // outputs of the first (usually only) pass
  if (r.slot >= 0) { // uniform per CTA: the presence words of the symbolic pass arrive as one bulk copy
    const unsigned bytes = (unsigned)((w.nword + 3) >> 2) * 16u;
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      mbar_arrive_expect_tx(&bar, bytes);
      bulk_load(bits, saved + (int64_t)r.slot * save_stride, bytes, &bar);
    }
    for (int i = threadIdx.x; i < first; i += THREADS) acc[i] = SR::identity(); // while the words are in flight
    __syncthreads(); // the barrier is initialised for everybody
    mbar_wait(&bar, 0);
  } else {
    bitmap_mark(s, k, &queue, bits, w.nword, w.rbase, restage);
    for (int i = threadIdx.x; i < first; i += THREADS) acc[i] = SR::identity();
  }
  bitmap_scan16<THREADS>(bits, rank, w.nword, wtot); // ends with __syncthreads: ranks and identities in place
  if (!ROWS_BY_WALK) {
    // rows unpacked from the words into Cir directly (the large shape: few words are empty, the walk dominates)
    for (int c = threadIdx.x; c < w.nword; c += THREADS) {
      unsigned b = bits[c];
      int32_t *o = Cir + obase + rank[c];
      const int rowbase = rbase + (c << 5);
      while (b) {
        *o++ = rowbase + __ffs(b) - 1;
        b &= b - 1;
      }
    }
  }
  const int upper = w.nword > 2048 ? (int)rank[2048] : nnz; // first output of the upper 2^16 rows
  // OVERFLOW (large shape only; the host sends it tasks of up to 65532 outputs): the first `cap` outputs accumulate in shared
  // memory as always, the rest go straight into C with one L2 reduction per product, as in num_bitmap_kernel -- which used to
  // take such a task whole. A task of 30 k outputs with room for 17.8 k keeps 59 % of its products out of the L2 reductions
  // (1.5 cycles per lane and SM against 0.4 for an exchange). Rows and identities of the overflowing outputs are written first.
  constexpr bool OVERFLOW = true; // only the large shape is ever sent such a task; compiling the branch out of the other shapes was measured 30 ms per step SLOWER (code generation at the 64-register limit: profiles/r2_sacc_v2_sweep.txt)
  if (OVERFLOW && nnz > cap) { // uniform per CTA
    for (int i = cap + (int)threadIdx.x; i < nnz; i += THREADS) Cval[obase + i] = SR::to_out(SR::identity());
    if (ROWS_BY_WALK) { // (without the row array the unpack above has already written all rows)
      for (int c = threadIdx.x; c < w.nword; c += THREADS) {
        unsigned b = bits[c];
        int o = (int)rank[c];
        if (o + __popc(b) <= cap) continue;
        const int rowbase = rbase + (c << 5);
        while (b) {
          if (o >= cap) Cir[obase + o] = rowbase + __ffs(b) - 1;
          ++o;
          b &= b - 1;
        }
      }
    }
    __syncthreads(); // the identities are in place before any thread of the CTA reduces into them
  }
  auto use = [&](int row, aval_t aval, typename SR::b_t mu) {
    const unsigned rr = (unsigned)(row - rbase);
    const unsigned wd = rr >> 5;
    const unsigned slot = (unsigned)rank[wd] + (unsigned)__popc(bits[wd] & ((1u << (rr & 31u)) - 1u));
    acc_t v;
    if (MERGE) v = SR::from_out((out_t)aval);
    else v = SR::mul((typename SR::a_t)aval, mu);
    if (!OVERFLOW || slot < (unsigned)cap) {
      if (ROWS_BY_WALK) srow[slot] = (unsigned short)rr;
      SR::template accumulate_shared<FIRST>(&acc[slot], v);
    } else {
      SR::accumulate_out(&Cval[obase + slot], v);
    }
  };
  bitmap_walk<SR, MERGE, true>(s, k, &queue, use, false, restage); // ends with __syncthreads
  if (ROWS_BY_WALK)
    for (int i = threadIdx.x; i < first; i += THREADS) Cir[obase + i] = rbase + (int)srow[i] + (i >= upper ? 65536 : 0);
  for (int i = threadIdx.x; i < first; i += THREADS) Cval[obase + i] = SR::to_out(acc[i]);
}

} // namespace cbgpu
