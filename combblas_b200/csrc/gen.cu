// Synthetic inputs: seeded R-MAT / Erdos-Renyi edge generator with identical integer arithmetic on host and
// device, plus device-side assembly of the edge list into a DCSC block. Input construction only -- outside
// the timed hot path (the reference's counterpart is DistEdgeList::GenGraph500Data, DistEdgeList.cpp:223-279,
// and the SpParMat(DistEdgeList) constructor, SpParMat.cpp:3153; initiator as in 3DSpGEMM/mpipspgemm.cpp:126-133).
// The radix sort used for assembly is CUB's (a library call, never on the multiply path).
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"
#include "util.cuh"

namespace cbgpu {

struct RmatParams {
  int scale;
  uint64_t seed;
  uint64_t ta, tab, tabc; // thresholds of a, a+b, a+b+c scaled to 2^53
  int scramble;
};

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t &s) {
  s += 0x9E3779B97F4A7C15ULL;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// bijection on [0, 2^scale): odd multiply, add, xorshift, odd multiply -- all modulo 2^scale
__host__ __device__ __forceinline__ uint64_t scramble_vertex(uint64_t v, int scale, uint64_t seed) {
  const uint64_t mask = (scale >= 64) ? ~0ULL : ((1ULL << scale) - 1);
  uint64_t s = seed ^ 0xD1B54A32D192ED03ULL;
  uint64_t m1 = splitmix64(s) | 1ULL, c1 = splitmix64(s), m2 = splitmix64(s) | 1ULL;
  v = (v * m1 + c1) & mask;
  int sh = scale / 2 > 0 ? scale / 2 : 1;
  v ^= v >> sh;
  v = (v * m2) & mask;
  return v;
}

__host__ __device__ __forceinline__ void rmat_edge(const RmatParams &p, uint64_t e, int64_t *row, int64_t *col) {
  uint64_t s = p.seed * 0x9E3779B97F4A7C15ULL + e * 0xD6E8FEB86659FD93ULL + 0x2545F4914F6CDD1DULL;
  uint64_t r = 0, c = 0;
  for (int l = 0; l < p.scale; ++l) {
    uint64_t u = splitmix64(s) >> 11; // 53 uniform bits
    int rb, cb;
    if (u < p.ta) { rb = 0; cb = 0; }
    else if (u < p.tab) { rb = 0; cb = 1; }
    else if (u < p.tabc) { rb = 1; cb = 0; }
    else { rb = 1; cb = 1; }
    r = (r << 1) | (uint64_t)rb;
    c = (c << 1) | (uint64_t)cb;
  }
  if (p.scramble) {
    r = scramble_vertex(r, p.scale, p.seed);
    c = scramble_vertex(c, p.scale, p.seed);
  }
  *row = (int64_t)r;
  *col = (int64_t)c;
}

static RmatParams make_params(int scale, uint64_t seed, double a, double b, double c, int scramble) {
  RmatParams p;
  p.scale = scale;
  p.seed = seed;
  const double two53 = 9007199254740992.0;
  p.ta = (uint64_t)(a * two53);
  p.tab = (uint64_t)((a + b) * two53);
  p.tabc = (uint64_t)((a + b + c) * two53);
  p.scramble = scramble;
  return p;
}

// keys of the edges that fall into the block [r0,r1) x [c0,c1), with block-local indices; every other edge gets the
// all-ones key, which sorts behind everything and is dropped
__global__ void rmat_keys_kernel(RmatParams p, int64_t nedges, int64_t r0, int64_t r1, int64_t c0, int64_t c1, uint64_t *keys,
                                 unsigned long long *inside) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool in = false;
  if (e < nedges) {
    int64_t r, c;
    rmat_edge(p, (uint64_t)e, &r, &c);
    in = r >= r0 && r < r1 && c >= c0 && c < c1;
    keys[e] = in ? (((uint64_t)(c - c0) << 32) | (uint64_t)(r - r0)) : ~0ull; // column-major order after sorting
  }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(inside, (unsigned long long)__popc(m));
}

__global__ void head_flags_kernel(const uint64_t *keys, int64_t n, int64_t *flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

template <class VT>
__global__ void emit_unique_kernel(const uint64_t *keys, const int64_t *pos, int64_t n, int value_mode, int64_t row_begin,
                                   int32_t *rows, VT *vals, int64_t *colcount) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = keys[i];
  if (i != 0 && keys[i - 1] == k) return;
  int64_t run = 1;
  while (i + run < n && keys[i + run] == k) ++run;
  int64_t o = pos[i];
  int32_t r = (int32_t)(k & 0xFFFFFFFFu);
  rows[o] = r;
  VT v;
  if (value_mode == 0) v = (VT)run;
  else if (value_mode == 1) v = (VT)1;
  else v = (VT)(1 + row_begin + (long long)r); // the GLOBAL row id
  vals[o] = v;
  atomicAdd((unsigned long long *)&colcount[k >> 32], 1ull);
}

} // namespace cbgpu

using namespace cbgpu;

extern "C" int cbgpu_rmat_edges_host(int scale, int64_t nedges, uint64_t seed, double a, double b, double c, int scramble,
                                     int64_t *rows, int64_t *cols) {
  if (scale < 1 || scale > 31 || nedges < 0 || !rows || !cols) return CBGPU_ERR_INVALID;
  RmatParams p = make_params(scale, seed, a, b, c, scramble);
  for (int64_t e = 0; e < nedges; ++e) rmat_edge(p, (uint64_t)e, &rows[e], &cols[e]);
  return CBGPU_OK;
}

template <class VT>
static int emit_unique(cbgpu_ctx *ctx, const uint64_t *keys, const int64_t *pos, int64_t n, int value_mode, int64_t row_begin,
                       cbgpu_mat *M, int64_t *colcount) {
  emit_unique_kernel<VT><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, pos, n, value_mode, row_begin, M->ir, (VT *)M->numx, colcount);
  CB_LAUNCH_CHECK(ctx);
  return CBGPU_OK;
}

extern "C" int cbgpu_gen_rmat(cbgpu_ctx *ctx, int scale, int64_t nedges, uint64_t seed, double a, double b, double c,
                              int scramble, int dtype, int value_mode, cbgpu_mat **out) {
  if (scale < 1 || scale > 30) return CBGPU_ERR_INVALID;
  const int64_t n = (int64_t)1 << scale;
  return cbgpu_gen_rmat_block(ctx, scale, nedges, seed, a, b, c, scramble, dtype, value_mode, 0, n, 0, n, out);
}

extern "C" int cbgpu_gen_rmat_block(cbgpu_ctx *ctx, int scale, int64_t nedges, uint64_t seed, double a, double b, double c,
                                    int scramble, int dtype, int value_mode, int64_t row_begin, int64_t row_end,
                                    int64_t col_begin, int64_t col_end, cbgpu_mat **out) {
  if (!ctx || !out || scale < 1 || scale > 30 || nedges < 1 || dtype_size(dtype) == 0) return CBGPU_ERR_INVALID;
  const int64_t n = (int64_t)1 << scale;
  if (row_begin < 0 || row_end > n || row_begin > row_end || col_begin < 0 || col_end > n || col_begin > col_end) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  RmatParams p = make_params(scale, seed, a, b, c, scramble);
  const int64_t bm = row_end - row_begin, bn = col_end - col_begin;
  uint64_t *keys = nullptr, *sorted = nullptr;
  unsigned long long *inside_d = nullptr;
  CB_TRY(dev_alloc_t(ctx, &keys, (size_t)nedges));
  CB_TRY(dev_alloc_t(ctx, &sorted, (size_t)nedges));
  CB_TRY(dev_alloc_t(ctx, &inside_d, 2));
  CB_CUDA(ctx, cudaMemsetAsync(inside_d, 0, 16, ctx->stream));
  rmat_keys_kernel<<<(unsigned)((nedges + 255) / 256), 256, 0, ctx->stream>>>(p, nedges, row_begin, row_end, col_begin, col_end, keys,
                                                                              inside_d);
  CB_LAUNCH_CHECK(ctx);
  unsigned long long inside = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&inside, inside_d, 8, cudaMemcpyDeviceToHost, ctx->stream));
  size_t tmp_bytes = 0;
  CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, sorted, nedges, 0, 64, ctx->stream));
  void *tmp = nullptr;
  CB_TRY(dev_alloc(ctx, &tmp, tmp_bytes));
  CB_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, sorted, nedges, 0, 64, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_free(ctx, tmp));
  CB_TRY(dev_free(ctx, keys));
  CB_TRY(dev_free(ctx, inside_d));
  const int64_t kept = (int64_t)inside; // the sorted keys of the block are the first `kept` entries
  int64_t *flags = nullptr, *pos = nullptr, *colcount = nullptr;
  CB_TRY(dev_alloc_t(ctx, &flags, (size_t)kept + 1));
  CB_TRY(dev_alloc_t(ctx, &pos, (size_t)kept + 1));
  if (kept > 0) {
    head_flags_kernel<<<(unsigned)((kept + 255) / 256), 256, 0, ctx->stream>>>(sorted, kept, flags);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(exclusive_scan_i64(ctx, flags, pos, kept));
  int64_t nnz = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&nnz, pos + kept, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_free(ctx, flags));
  cbgpu_mat *M = nullptr;
  CB_TRY(mat_alloc(ctx, bm, bn, nnz, -1, dtype, &M));
  CB_TRY(dev_alloc_t(ctx, &colcount, (size_t)bn + 1));
  CB_CUDA(ctx, cudaMemsetAsync(colcount, 0, ((size_t)bn + 1) * 8, ctx->stream));
  int rc = CBGPU_OK;
  if (kept > 0) {
    switch (dtype) {
      case CBGPU_F64: rc = emit_unique<double>(ctx, sorted, pos, kept, value_mode, row_begin, M, colcount); break;
      case CBGPU_F32: rc = emit_unique<float>(ctx, sorted, pos, kept, value_mode, row_begin, M, colcount); break;
      case CBGPU_I64: rc = emit_unique<long long>(ctx, sorted, pos, kept, value_mode, row_begin, M, colcount); break;
      case CBGPU_I32: rc = emit_unique<int>(ctx, sorted, pos, kept, value_mode, row_begin, M, colcount); break;
      case CBGPU_BOOL: rc = emit_unique<uint8_t>(ctx, sorted, pos, kept, 1, row_begin, M, colcount); break;
    }
  }
  if (rc == CBGPU_OK) rc = dev_alloc_t(ctx, &M->colptr, (size_t)bn + 1);
  if (rc == CBGPU_OK) rc = exclusive_scan_i64(ctx, colcount, M->colptr, bn);
  if (rc == CBGPU_OK) rc = compact_columns(ctx, nullptr, M->colptr, bn, &M->jc, &M->cp, &M->nzc);
  dev_free(ctx, colcount);
  dev_free(ctx, pos);
  dev_free(ctx, sorted);
  if (rc != CBGPU_OK) { mat_release(ctx, M); return rc; }
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = M;
  return CBGPU_OK;
}
