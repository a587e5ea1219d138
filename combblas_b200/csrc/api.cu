// extern "C" surface of libcbgpu.so (include/cbgpu.h): lifecycle, DCSC staging in HBM, local multiply,
// merge, column slabs. Distributed entry points live in dist.cu, the synthetic generators in gen.cu.
#include <string.h>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "util.cuh"

using namespace cbgpu;

namespace cbgpu {

#define DECL_SR(i)                                                                                                     \
  int spgemm_sr##i(const SpgemmArgs &);                                                                                \
  int merge_sr##i(const MergeArgs &);
DECL_SR(0) DECL_SR(1) DECL_SR(2) DECL_SR(3) DECL_SR(4) DECL_SR(5) DECL_SR(6) DECL_SR(7) DECL_SR(8)
DECL_SR(9) DECL_SR(10) DECL_SR(11) DECL_SR(12) DECL_SR(13) DECL_SR(14)

// user-defined semirings (include/combblas_b200/device_semiring.cuh): entry points instantiated in a translation unit of the
// application, registered at run time under ids from CBGPU_SR_USER_BASE on
struct UserSemiring {
  spgemm_fn spgemm;
  merge_fn merge;
  int ta, tb, tc;
};
static std::mutex &user_sr_mutex() {
  static std::mutex m;
  return m;
}
static std::vector<UserSemiring> &user_srs() {
  static std::vector<UserSemiring> v;
  return v;
}
static bool user_sr(int sr, UserSemiring *out) {
  std::lock_guard<std::mutex> lock(user_sr_mutex());
  const int i = sr - CBGPU_SR_USER_BASE;
  if (i < 0 || i >= (int)user_srs().size()) return false;
  *out = user_srs()[(size_t)i];
  return true;
}

spgemm_fn spgemm_entry(int sr) {
  static const spgemm_fn t[CBGPU_SR_COUNT] = {spgemm_sr0, spgemm_sr1, spgemm_sr2, spgemm_sr3, spgemm_sr4,
                                              spgemm_sr5, spgemm_sr6, spgemm_sr7, spgemm_sr8, spgemm_sr9,
                                              spgemm_sr10, spgemm_sr11, spgemm_sr12, spgemm_sr13, spgemm_sr14};
  UserSemiring u;
  if (user_sr(sr, &u)) return u.spgemm;
  return (sr >= 0 && sr < CBGPU_SR_COUNT) ? t[sr] : nullptr;
}
merge_fn merge_entry(int sr) {
  static const merge_fn t[CBGPU_SR_COUNT] = {merge_sr0, merge_sr1, merge_sr2, merge_sr3, merge_sr4,
                                             merge_sr5, merge_sr6, merge_sr7, merge_sr8, merge_sr9,
                                             merge_sr10, merge_sr11, merge_sr12, merge_sr13, merge_sr14};
  UserSemiring u;
  if (user_sr(sr, &u)) return u.merge;
  return (sr >= 0 && sr < CBGPU_SR_COUNT) ? t[sr] : nullptr;
}
int semiring_types(int sr, int *a, int *b, int *c) {
  static const int t[CBGPU_SR_COUNT][3] = {
      {CBGPU_F64, CBGPU_F64, CBGPU_F64}, {CBGPU_F32, CBGPU_F32, CBGPU_F32},  {CBGPU_I64, CBGPU_I64, CBGPU_I64},
      {CBGPU_BOOL, CBGPU_I64, CBGPU_I64}, {CBGPU_F64, CBGPU_F64, CBGPU_F64}, {CBGPU_BOOL, CBGPU_BOOL, CBGPU_BOOL},
      {CBGPU_BOOL, CBGPU_F64, CBGPU_F64}, {CBGPU_I32, CBGPU_I32, CBGPU_I32}, {CBGPU_I64, CBGPU_I64, CBGPU_I64},
      {CBGPU_BOOL, CBGPU_F64, CBGPU_F64}, {CBGPU_F64, CBGPU_BOOL, CBGPU_F64}, {CBGPU_BOOL, CBGPU_I64, CBGPU_I64},
      {CBGPU_I64, CBGPU_BOOL, CBGPU_I64}, {CBGPU_BOOL, CBGPU_BOOL, CBGPU_BOOL}, {CBGPU_BOOL, CBGPU_BOOL, CBGPU_BOOL}};
  UserSemiring u;
  if (user_sr(sr, &u)) {
    *a = u.ta; *b = u.tb; *c = u.tc;
    return CBGPU_OK;
  }
  if (sr < 0 || sr >= CBGPU_SR_COUNT) return CBGPU_ERR_INVALID;
  *a = t[sr][0]; *b = t[sr][1]; *c = t[sr][2];
  return CBGPU_OK;
}
int register_user_semiring(spgemm_fn spgemm, merge_fn merge, int ta, int tb, int tc) {
  if (!spgemm || !merge || dtype_size(ta) == 0 || dtype_size(tb) == 0 || dtype_size(tc) == 0) return CBGPU_ERR_INVALID;
  std::lock_guard<std::mutex> lock(user_sr_mutex());
  user_srs().push_back(UserSemiring{spgemm, merge, ta, tb, tc});
  return CBGPU_SR_USER_BASE + (int)user_srs().size() - 1;
}

// ---- small conversion kernels for the staging path
__global__ void narrow_i64_i32(const int64_t *in, int32_t *out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}
__global__ void widen_i32_i64(const int32_t *in, int64_t *out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void narrow_ptr_i64_i32(const int64_t *in, int32_t *out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}
// one warp per stored column: expand column ids (and optionally widen rows) for COO export
template <class IT>
__global__ void expand_coo_kernel(const int64_t *jc, const int64_t *cp, const int32_t *ir, int64_t nzc, IT *rows, IT *cols) {
  int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nzc) return;
  IT col = (IT)jc[c];
  for (int64_t p = cp[c] + (threadIdx.x & 31); p < cp[c + 1]; p += 32) {
    rows[p] = (IT)ir[p];
    cols[p] = col;
  }
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__global__ void checksum_kernel(const int64_t *jc, const int64_t *cp, const int32_t *ir, const unsigned char *vals,
                                int vbytes, int64_t nzc, int64_t row_offset, int64_t col_offset, unsigned long long *sums) {
  int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint64_t ps = 0, vs = 0;
  if (c < nzc) {
    uint64_t col = (uint64_t)(jc[c] + col_offset);
    for (int64_t p = cp[c] + (threadIdx.x & 31); p < cp[c + 1]; p += 32) {
      uint64_t key = (col << 32) ^ (uint64_t)(uint32_t)((int64_t)ir[p] + row_offset);
      uint64_t h = mix64(key + 0x9E3779B97F4A7C15ULL);
      uint64_t vb = 0;
      for (int b = 0; b < vbytes; ++b) vb |= (uint64_t)vals[p * vbytes + b] << (8 * b);
      if (vbytes == 8 && vb == 0x8000000000000000ULL) vb = 0; // -0.0 == +0.0
      if (vbytes == 4 && vb == 0x80000000ULL) vb = 0;
      ps += h;
      vs += mix64(h ^ mix64(vb + 0x632BE59BD9B4E019ULL));
    }
  }
  for (int d = 16; d >= 1; d >>= 1) {
    ps += __shfl_xor_sync(0xFFFFFFFFu, ps, d);
    vs += __shfl_xor_sync(0xFFFFFFFFu, vs, d);
  }
  if ((threadIdx.x & 31) == 0 && c < nzc) {
    atomicAdd(&sums[0], (unsigned long long)ps);
    atomicAdd(&sums[1], (unsigned long long)vs);
  }
}
__global__ void lower_bound_kernel(const int64_t *jc, int64_t nzc, int64_t v0, int64_t v1, int64_t *out) {
  if (threadIdx.x < 2) {
    int64_t v = threadIdx.x == 0 ? v0 : v1;
    int64_t a = 0, b = nzc;
    while (a < b) {
      int64_t mid = (a + b) >> 1;
      if (jc[mid] < v) a = mid + 1;
      else b = mid;
    }
    out[threadIdx.x] = a;
  }
}
__global__ void rebase_kernel(const int64_t *in, int64_t n, int64_t sub, int64_t add, int64_t *out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] - sub + add;
}

static inline unsigned nblocks(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

static int upload_index_array(cbgpu_ctx_impl *ctx, const void *host, int idx_bytes, int64_t count, int64_t *dst64,
                              int32_t *dst32) {
  if (count <= 0) return CBGPU_OK;
  if (dst64) {
    if (idx_bytes == 8) {
      CB_CUDA(ctx, cudaMemcpyAsync(dst64, host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      int32_t *tmp = nullptr;
      CB_TRY(dev_alloc_t(ctx, &tmp, (size_t)count));
      CB_CUDA(ctx, cudaMemcpyAsync(tmp, host, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream));
      widen_i32_i64<<<nblocks(count), 256, 0, ctx->stream>>>(tmp, dst64, count);
      CB_LAUNCH_CHECK(ctx);
      CB_TRY(dev_free(ctx, tmp));
    }
  } else {
    if (idx_bytes == 4) {
      CB_CUDA(ctx, cudaMemcpyAsync(dst32, host, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      int64_t *tmp = nullptr;
      CB_TRY(dev_alloc_t(ctx, &tmp, (size_t)count));
      CB_CUDA(ctx, cudaMemcpyAsync(tmp, host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
      narrow_i64_i32<<<nblocks(count), 256, 0, ctx->stream>>>(tmp, dst32, count);
      CB_LAUNCH_CHECK(ctx);
      CB_TRY(dev_free(ctx, tmp));
    }
  }
  return CBGPU_OK;
}

static int download_index_array(cbgpu_ctx_impl *ctx, void *host, int idx_bytes, int64_t count, const int64_t *src64,
                                const int32_t *src32) {
  if (count <= 0 || !host) return CBGPU_OK;
  if (src64) {
    if (idx_bytes == 8) {
      CB_CUDA(ctx, cudaMemcpyAsync(host, src64, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      int32_t *tmp = nullptr;
      CB_TRY(dev_alloc_t(ctx, &tmp, (size_t)count));
      narrow_ptr_i64_i32<<<nblocks(count), 256, 0, ctx->stream>>>(src64, tmp, count);
      CB_LAUNCH_CHECK(ctx);
      CB_CUDA(ctx, cudaMemcpyAsync(host, tmp, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CB_TRY(dev_free(ctx, tmp));
    }
  } else {
    if (idx_bytes == 4) {
      CB_CUDA(ctx, cudaMemcpyAsync(host, src32, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      int64_t *tmp = nullptr;
      CB_TRY(dev_alloc_t(ctx, &tmp, (size_t)count));
      widen_i32_i64<<<nblocks(count), 256, 0, ctx->stream>>>(src32, tmp, count);
      CB_LAUNCH_CHECK(ctx);
      CB_CUDA(ctx, cudaMemcpyAsync(host, tmp, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CB_TRY(dev_free(ctx, tmp));
    }
  }
  return CBGPU_OK;
}

int mat_colslice(cbgpu_ctx_impl *ctx, const cbgpu_mat_impl *M, int64_t c0, int64_t c1, cbgpu_mat_impl **out) {
  if (c0 < 0 || c1 < c0 || c1 > M->n) return set_error(ctx, CBGPU_ERR_INVALID, "bad column range");
  int64_t lb[2] = {0, 0}, pb[2] = {0, 0};
  if (M->nzc > 0) {
    int64_t *d = nullptr;
    CB_TRY(dev_alloc_t(ctx, &d, 2));
    lower_bound_kernel<<<1, 32, 0, ctx->stream>>>(M->jc, M->nzc, c0, c1, d);
    CB_LAUNCH_CHECK(ctx);
    CB_CUDA(ctx, cudaMemcpyAsync(lb, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(&pb[0], M->cp + lb[0], 8, cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(&pb[1], M->cp + lb[1], 8, cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CB_TRY(dev_free(ctx, d));
  }
  int64_t nzc = lb[1] - lb[0], nnz = pb[1] - pb[0];
  cbgpu_mat_impl *S = nullptr;
  CB_TRY(mat_alloc(ctx, M->m, c1 - c0, nnz, nzc, M->dtype, &S));
  if (nzc > 0) {
    rebase_kernel<<<nblocks(nzc), 256, 0, ctx->stream>>>(M->jc + lb[0], nzc, c0, 0, S->jc);
    CB_LAUNCH_CHECK(ctx);
  }
  rebase_kernel<<<nblocks(nzc + 1), 256, 0, ctx->stream>>>(M->cp ? M->cp + lb[0] : nullptr, M->nzc > 0 ? nzc + 1 : 0, pb[0], 0, S->cp);
  CB_LAUNCH_CHECK(ctx);
  if (M->nzc == 0) CB_CUDA(ctx, cudaMemsetAsync(S->cp, 0, 8, ctx->stream));
  if (nnz > 0) {
    CB_CUDA(ctx, cudaMemcpyAsync(S->ir, M->ir + pb[0], (size_t)nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    size_t vb = dtype_size(M->dtype);
    CB_CUDA(ctx, cudaMemcpyAsync(S->numx, (const char *)M->numx + (size_t)pb[0] * vb, (size_t)nnz * vb,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
  }
  *out = S;
  return CBGPU_OK;
}

// rows [r0, r1) of every stored column: counts, then copy with rebased row ids (rows are ascending per column)
__global__ void rowrange_count_kernel(const int64_t *cp, const int32_t *ir, int64_t nzc, int64_t r0, int64_t r1, int64_t *lo,
                                      int64_t *cnt) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nzc) return;
  int64_t a = cp[c], b = cp[c + 1], e = b;
  while (a < b) {
    int64_t mid = (a + b) >> 1;
    if (ir[mid] < r0) a = mid + 1;
    else b = mid;
  }
  int64_t first = a;
  b = e;
  while (a < b) {
    int64_t mid = (a + b) >> 1;
    if (ir[mid] < r1) a = mid + 1;
    else b = mid;
  }
  lo[c] = first;
  cnt[c] = a - first;
}
__global__ void rowrange_copy_kernel(const int64_t *lo, const int64_t *newptr, const int32_t *ir, const unsigned char *vals,
                                     int vbytes, int64_t nzc, int32_t r0, int32_t *oir, unsigned char *ovals) {
  int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nzc) return;
  int64_t src = lo[c], dst = newptr[c], n = newptr[c + 1] - dst;
  for (int64_t i = threadIdx.x & 31; i < n; i += 32) {
    oir[dst + i] = ir[src + i] - r0;
    for (int b = 0; b < vbytes; ++b) ovals[(dst + i) * vbytes + b] = vals[(src + i) * vbytes + b];
  }
}

int mat_submatrix(cbgpu_ctx_impl *ctx, const cbgpu_mat_impl *M, int64_t r0, int64_t r1, int64_t c0, int64_t c1,
                  cbgpu_mat_impl **out) {
  if (r0 < 0 || r1 < r0 || r1 > M->m) return set_error(ctx, CBGPU_ERR_INVALID, "bad row range");
  cbgpu_mat_impl *S = nullptr;
  CB_TRY(mat_colslice(ctx, M, c0, c1, &S));
  if (r0 == 0 && r1 == M->m) {
    *out = S;
    return CBGPU_OK;
  }
  int64_t *lo = nullptr, *cnt = nullptr, *newptr = nullptr;
  CB_TRY(dev_alloc_t(ctx, &lo, (size_t)S->nzc + 1));
  CB_TRY(dev_alloc_t(ctx, &cnt, (size_t)S->nzc + 1));
  CB_TRY(dev_alloc_t(ctx, &newptr, (size_t)S->nzc + 1));
  if (S->nzc > 0) {
    rowrange_count_kernel<<<nblocks(S->nzc), 256, 0, ctx->stream>>>(S->cp, S->ir, S->nzc, r0, r1, lo, cnt);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(exclusive_scan_i64(ctx, cnt, newptr, S->nzc));
  int64_t nnz = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&nnz, newptr + S->nzc, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cbgpu_mat_impl *R = nullptr;
  CB_TRY(mat_alloc(ctx, r1 - r0, S->n, nnz, -1, S->dtype, &R));
  if (S->nzc > 0) {
    rowrange_copy_kernel<<<nblocks(S->nzc * 32), 256, 0, ctx->stream>>>(lo, newptr, S->ir, (const unsigned char *)S->numx,
                                                                       (int)dtype_size(S->dtype), S->nzc, (int32_t)r0, R->ir,
                                                                       (unsigned char *)R->numx);
    CB_LAUNCH_CHECK(ctx);
  }
  int rc = compact_columns(ctx, S->jc, newptr, S->nzc, &R->jc, &R->cp, &R->nzc);
  dev_free(ctx, lo);
  dev_free(ctx, cnt);
  dev_free(ctx, newptr);
  mat_release(ctx, S);
  if (rc != CBGPU_OK) {
    mat_release(ctx, R);
    return rc;
  }
  *out = R;
  return CBGPU_OK;
}

// vertical stack [B_0; B_1; ...] of blocks with equal column count: column j of the result is the concatenation of the
// parts' columns j with their rows shifted by the rows of the parts above. Used by the fused SUMMA, where
// sum_i A_i * B_i is computed as ONE multiply [A_0 A_1 ...] * [B_0; B_1; ...] instead of per-stage products + a merge.
constexpr int kMaxStack = 16;
struct StackParts {
  int parts;
  const int64_t *colptr[kMaxStack];
  const int32_t *ir[kMaxStack];
  const unsigned char *val[kMaxStack];
  int64_t rowoff[kMaxStack];
};
__global__ void rowstack_count_kernel(StackParts sp, int64_t n, int64_t *cnt) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int64_t c = 0;
  for (int i = 0; i < sp.parts; ++i) c += sp.colptr[i][j + 1] - sp.colptr[i][j];
  cnt[j] = c;
}
template <int VB>
__global__ void rowstack_copy_kernel(StackParts sp, int64_t n, const int64_t *outptr, int32_t *oir, unsigned char *oval) {
  int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n) return;
  const int lane = threadIdx.x & 31;
  int64_t dst = outptr[j];
  for (int i = 0; i < sp.parts; ++i) {
    int64_t src = sp.colptr[i][j], len = sp.colptr[i][j + 1] - src;
    for (int64_t q = lane; q < len; q += 32) {
      oir[dst + q] = sp.ir[i][src + q] + (int32_t)sp.rowoff[i];
      if (VB == 8) reinterpret_cast<uint64_t *>(oval)[dst + q] = reinterpret_cast<const uint64_t *>(sp.val[i])[src + q];
      else if (VB == 4) reinterpret_cast<uint32_t *>(oval)[dst + q] = reinterpret_cast<const uint32_t *>(sp.val[i])[src + q];
      else oval[dst + q] = sp.val[i][src + q];
    }
    dst += len;
  }
}

int mat_rowstack(cbgpu_ctx_impl *ctx, int parts, cbgpu_mat_impl *const *in, cbgpu_mat_impl **out) {
  if (parts < 1) return set_error(ctx, CBGPU_ERR_INVALID, "rowstack needs at least one part");
  if (parts > kMaxStack) {
    // more parts than one kernel takes (grids wider than 16 process rows / more than 16 layers): stack in rounds, the
    // running result is the first part of the next round
    cbgpu_mat_impl *acc = nullptr;
    int done = 0;
    while (done < parts) {
      std::vector<cbgpu_mat_impl *> round;
      if (acc) round.push_back(acc);
      while (done < parts && (int)round.size() < kMaxStack) round.push_back(in[done++]);
      cbgpu_mat_impl *next = nullptr;
      int rc = mat_rowstack(ctx, (int)round.size(), round.data(), &next);
      mat_release(ctx, acc);
      if (rc != CBGPU_OK) return rc;
      acc = next;
    }
    *out = acc;
    return CBGPU_OK;
  }
  StackParts sp;
  memset(&sp, 0, sizeof(sp));
  sp.parts = parts;
  int64_t m = 0, nnz = 0;
  const int64_t n = in[0]->n;
  for (int i = 0; i < parts; ++i) {
    if (in[i]->n != n || in[i]->dtype != in[0]->dtype) return set_error(ctx, CBGPU_ERR_DIMMISMATCH, "rowstack: column count or type differs");
    CB_TRY(ensure_dense_colptr(ctx, in[i]));
    sp.colptr[i] = in[i]->colptr;
    sp.ir[i] = in[i]->ir;
    sp.val[i] = (const unsigned char *)in[i]->numx;
    sp.rowoff[i] = m;
    m += in[i]->m;
    nnz += in[i]->nnz;
  }
  if (m >= ((int64_t)1 << 31) - 1) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "stacked block has too many rows");
  cbgpu_mat_impl *S = nullptr;
  CB_TRY(mat_alloc(ctx, m, n, nnz, -1, in[0]->dtype, &S));
  int64_t *cnt = nullptr;
  CB_TRY(dev_alloc_t(ctx, &cnt, (size_t)n + 1));
  CB_TRY(dev_alloc_t(ctx, &S->colptr, (size_t)n + 1));
  if (n > 0) {
    rowstack_count_kernel<<<nblocks(n), 256, 0, ctx->stream>>>(sp, n, cnt);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(exclusive_scan_i64(ctx, cnt, S->colptr, n));
  if (n > 0 && nnz > 0) {
    const int vb = (int)dtype_size(S->dtype);
    unsigned nb = nblocks(n * 32);
    if (vb == 8) rowstack_copy_kernel<8><<<nb, 256, 0, ctx->stream>>>(sp, n, S->colptr, S->ir, (unsigned char *)S->numx);
    else if (vb == 4) rowstack_copy_kernel<4><<<nb, 256, 0, ctx->stream>>>(sp, n, S->colptr, S->ir, (unsigned char *)S->numx);
    else rowstack_copy_kernel<1><<<nb, 256, 0, ctx->stream>>>(sp, n, S->colptr, S->ir, (unsigned char *)S->numx);
    CB_LAUNCH_CHECK(ctx);
  }
  int rc = compact_columns(ctx, nullptr, S->colptr, n, &S->jc, &S->cp, &S->nzc);
  dev_free(ctx, cnt);
  if (rc != CBGPU_OK) {
    mat_release(ctx, S);
    return rc;
  }
  *out = S;
  return CBGPU_OK;
}

int mat_colconcat(cbgpu_ctx_impl *ctx, int parts, cbgpu_mat_impl *const *in, cbgpu_mat_impl **out) {
  if (parts < 1) return set_error(ctx, CBGPU_ERR_INVALID, "colconcat needs at least one part");
  int64_t n = 0, nnz = 0, nzc = 0;
  for (int i = 0; i < parts; ++i) {
    if (in[i]->m != in[0]->m || in[i]->dtype != in[0]->dtype)
      return set_error(ctx, CBGPU_ERR_DIMMISMATCH, "colconcat: row dimension or type differs");
    n += in[i]->n; nnz += in[i]->nnz; nzc += in[i]->nzc;
  }
  cbgpu_mat_impl *S = nullptr;
  CB_TRY(mat_alloc(ctx, in[0]->m, n, nnz, nzc, in[0]->dtype, &S));
  int64_t coff = 0, poff = 0, zoff = 0;
  size_t vb = dtype_size(S->dtype);
  for (int i = 0; i < parts; ++i) {
    const cbgpu_mat_impl *P = in[i];
    if (P->nzc > 0) {
      rebase_kernel<<<nblocks(P->nzc), 256, 0, ctx->stream>>>(P->jc, P->nzc, 0, coff, S->jc + zoff);
      CB_LAUNCH_CHECK(ctx);
      rebase_kernel<<<nblocks(P->nzc), 256, 0, ctx->stream>>>(P->cp, P->nzc, 0, poff, S->cp + zoff);
      CB_LAUNCH_CHECK(ctx);
    }
    if (P->nnz > 0) {
      CB_CUDA(ctx, cudaMemcpyAsync(S->ir + poff, P->ir, (size_t)P->nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      CB_CUDA(ctx, cudaMemcpyAsync((char *)S->numx + (size_t)poff * vb, P->numx, (size_t)P->nnz * vb,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    }
    coff += P->n; poff += P->nnz; zoff += P->nzc;
  }
  CB_CUDA(ctx, cudaMemcpyAsync(S->cp + nzc, &nnz, 8, cudaMemcpyHostToDevice, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // &nnz is a stack variable
  *out = S;
  return CBGPU_OK;
}

} // namespace cbgpu

extern "C" {

int cbgpu_version(void) { return CBGPU_VERSION; }

int cbgpu_device_count(int *count) {
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    return CBGPU_ERR_CUDA;
  }
  return CBGPU_OK;
}

int cbgpu_create(int device, void *stream, cbgpu_ctx **out) {
  if (!out) return CBGPU_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    fprintf(stderr, "[cbgpu] no usable CUDA device %d (found %d); the device path has no CPU fallback\n", device, n);
    return CBGPU_ERR_CUDA;
  }
  cbgpu_ctx *ctx = new cbgpu_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CBGPU_ERR_CUDA; }
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CBGPU_ERR_CUDA; }
    ctx->own_stream = true;
  }
  for (int i = 0; i < 6; ++i) cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 2 * CBGPU_K_COUNT; ++i) cudaEventCreate(&ctx->kev[i]);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  // keep freed blocks in the stream-ordered pool: repeated multiplies must not hit cudaMalloc/cudaFree
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = ctx;
  return CBGPU_OK;
}

int cbgpu_destroy(cbgpu_ctx *ctx) {
  if (!ctx) return CBGPU_OK;
  cudaSetDevice(ctx->device);
  release_cached_blocks(ctx);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2 * CBGPU_K_COUNT; ++i) if (ctx->kev[i]) cudaEventDestroy(ctx->kev[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CBGPU_OK;
}

const char *cbgpu_last_error(const cbgpu_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }

int cbgpu_memory_in_use(cbgpu_ctx *ctx, int64_t *live_bytes) {
  if (!ctx || !live_bytes) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  return pool_live_bytes(ctx, live_bytes);
}

int cbgpu_sync(cbgpu_ctx *ctx) {
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CBGPU_OK;
}

static int64_t *option_slot(cbgpu_ctx *ctx, const char *name) {
  Options &o = ctx->opt;
  if (!strcmp(name, "bitmap_window_log2")) return &o.bitmap_window_log2;
  if (!strcmp(name, "bitmap_min_nnz")) return &o.bitmap_min_nnz;
  if (!strcmp(name, "shared_acc")) return &o.shared_acc;
  if (!strcmp(name, "shared_acc_max")) return &o.shared_acc_max;
  if (!strcmp(name, "shared_acc_small_max")) return &o.shared_acc_small_max;
  if (!strcmp(name, "bitmap_cta_threads")) return &o.bitmap_cta_threads;
  if (!strcmp(name, "bitmap_small_threads")) return &o.bitmap_small_threads;
  if (!strcmp(name, "bitmap_save_mb")) return &o.bitmap_save_mb;
  if (!strcmp(name, "bitmap_save_min_flop")) return &o.bitmap_save_min_flop;
  if (!strcmp(name, "light_max")) return &o.light_max;
  if (!strcmp(name, "bitmap_small_minblocks")) return &o.bitmap_small_minblocks;
  if (!strcmp(name, "force_path")) return &o.force_path;
  if (!strcmp(name, "merge_engine")) return &o.merge_engine;
  if (!strcmp(name, "regsort")) return &o.regsort;
  if (!strcmp(name, "regsort_packed")) return &o.regsort_packed;
  if (!strcmp(name, "merge_tma")) return &o.merge_tma;
  if (!strcmp(name, "validate_uploads")) return &o.validate_uploads;
  if (!strcmp(name, "sacc_v2")) return &o.sacc_v2;
  if (!strcmp(name, "sacc_overflow")) return &o.sacc_overflow;
  if (!strcmp(name, "summa_fused")) return &o.summa_fused;
  if (!strcmp(name, "fiber_fused")) return &o.fiber_fused;
  if (!strcmp(name, "fiber_pipeline")) return &o.fiber_pipeline;
  return nullptr;
}
int cbgpu_calculate_phases(int64_t max_local_nnz_a, int64_t nnz_product_per_process, int idx_bytes, int in_val_bytes,
                           int out_val_bytes, int64_t per_process_memory_gb) {
  if (max_local_nnz_a < 0 || nnz_product_per_process < 0 || idx_bytes <= 0 || in_val_bytes <= 0 || out_val_bytes <= 0)
    return CBGPU_ERR_INVALID;
  const int64_t per_in = 2 * (int64_t)idx_bytes + in_val_bytes, per_out = 2 * (int64_t)idx_bytes + out_val_bytes;
  const int64_t input_mem = max_local_nnz_a * per_in * 4;           // four copies, two of them SUMMA's (ParFriends.h:802)
  const int64_t asquare_mem = nnz_product_per_process * per_out * 2; // an extra copy in the merge / selection (:806)
  const int64_t remaining = per_process_memory_gb * 1000000000ll - input_mem; // every phase result is discarded (:825)
  if (remaining <= 0) return CBGPU_ERR_INVALID;
  const int64_t phases = 1 + asquare_mem / remaining;
  return phases > 0x7FFFFFFF ? 0x7FFFFFFF : (int)phases;
}

int cbgpu_set_option(cbgpu_ctx *ctx, const char *name, int64_t value) {
  int64_t *s = option_slot(ctx, name);
  if (!s) return set_error(ctx, CBGPU_ERR_INVALID, "unknown option %s", name);
  if (!strcmp(name, "bitmap_cta_threads") && value != 256 && value != 512)
    return set_error(ctx, CBGPU_ERR_INVALID, "bitmap_cta_threads must be 256 or 512");
  if (!strcmp(name, "bitmap_small_threads") && value != 128 && value != 256)
    return set_error(ctx, CBGPU_ERR_INVALID, "bitmap_small_threads must be 128 or 256");
  *s = value;
  return CBGPU_OK;
}
int cbgpu_get_option(cbgpu_ctx *ctx, const char *name, int64_t *value) {
  int64_t *s = option_slot(ctx, name);
  if (!s) return set_error(ctx, CBGPU_ERR_INVALID, "unknown option %s", name);
  *value = *s;
  return CBGPU_OK;
}
int64_t cbgpu_launch_count(const cbgpu_ctx *ctx) { return ctx->launches; }

// ---- structural validation of a resident block (the engine's window searches and merges assume it)
static __global__ void validate_cols_kernel(const int64_t *jc, const int64_t *cp, int64_t nzc, int64_t n, int64_t nnz, int *bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nzc) return;
  if (i == 0 && cp[0] != 0) atomicOr(bad, 1);
  if (i == nzc) {
    if (cp[nzc] != nnz) atomicOr(bad, 1);
    return;
  }
  if (cp[i + 1] <= cp[i]) atomicOr(bad, 1);                       // a listed column is non-empty (dcsc.h:125-132), pointers ascend
  if (jc[i] < 0 || jc[i] >= n || (i > 0 && jc[i] <= jc[i - 1])) atomicOr(bad, 2); // column ids ascending, in range
}
static __global__ void validate_rows_kernel(const int64_t *cp, int64_t nzc, const int32_t *ir, int64_t m, int *bad) {
  // one warp per listed column: rows in range and strictly ascending
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nzc) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = cp[c], e = cp[c + 1];
  for (int64_t p = b + lane; p < e; p += 32) {
    const int32_t r = ir[p];
    if (r < 0 || (int64_t)r >= m) atomicOr(bad, 4);
    if (p > b && ir[p - 1] >= r) atomicOr(bad, 8);
  }
}

int cbgpu_mat_validate(cbgpu_ctx *ctx, const cbgpu_mat *M) {
  if (!ctx || !M) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (M->nnz == 0 && M->nzc == 0) return CBGPU_OK;
  if (M->nzc <= 0 || M->nnz < M->nzc) return set_error(ctx, CBGPU_ERR_INVALID, "block with nnz=%lld, nzc=%lld", (long long)M->nnz, (long long)M->nzc);
  Scratch scratch(ctx);
  int *bad = nullptr;
  CB_TRY(scratch.alloc(&bad, 1));
  CB_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
  validate_cols_kernel<<<(unsigned)((M->nzc + 1 + 255) / 256), 256, 0, ctx->stream>>>(M->jc, M->cp, M->nzc, M->n, M->nnz, bad);
  CB_LAUNCH_CHECK(ctx);
  int h = 0;
  CB_CUDA(ctx, cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h == 0) { // the row check follows the column pointers, so they must be sound first
    validate_rows_kernel<<<(unsigned)((M->nzc * 32 + 255) / 256), 256, 0, ctx->stream>>>(M->cp, M->nzc, M->ir, M->m, bad);
    CB_LAUNCH_CHECK(ctx);
    CB_CUDA(ctx, cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (h == 0) return CBGPU_OK;
  return set_error(ctx, CBGPU_ERR_INVALID, "block is not a valid DCSC:%s%s%s%s", (h & 1) ? " column pointers not ascending from 0 to nnz;" : "",
                   (h & 2) ? " column ids not ascending or out of range;" : "", (h & 4) ? " row ids out of range;" : "",
                   (h & 8) ? " row ids not strictly ascending inside a column (sort the block: the reference's sort=false paths leave them unordered);" : "");
}

int cbgpu_mat_upload(cbgpu_ctx *ctx, const cbgpu_dcsc_view *h, cbgpu_mat **out) {
  if (!ctx || !h || !out) return CBGPU_ERR_INVALID;
  if (h->idx_bytes != 4 && h->idx_bytes != 8) return set_error(ctx, CBGPU_ERR_INVALID, "idx_bytes must be 4 or 8");
  if (dtype_size(h->dtype) == 0) return set_error(ctx, CBGPU_ERR_INVALID, "unknown dtype %d", h->dtype);
  if (h->m >= ((int64_t)1 << 31) - 1 || h->n >= ((int64_t)1 << 31) - 1)
    return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "local block dimensions must stay below 2^31-1");
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_mat_impl *M = nullptr;
  CB_TRY(mat_alloc(ctx, h->m, h->n, h->nnz, h->nzc, h->dtype, &M));
  int rc = CBGPU_OK;
  if (h->nzc > 0) {
    if ((rc = upload_index_array(ctx, h->jc, h->idx_bytes, h->nzc, M->jc, nullptr)) != CBGPU_OK) goto fail;
    if ((rc = upload_index_array(ctx, h->cp, h->idx_bytes, h->nzc + 1, M->cp, nullptr)) != CBGPU_OK) goto fail;
  } else {
    cudaMemsetAsync(M->cp, 0, 8, ctx->stream);
  }
  if (h->nnz > 0) {
    if ((rc = upload_index_array(ctx, h->ir, h->idx_bytes, h->nnz, nullptr, M->ir)) != CBGPU_OK) goto fail;
    cudaError_t e = cudaMemcpyAsync(M->numx, h->numx, (size_t)h->nnz * dtype_size(h->dtype), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { rc = set_error(ctx, CBGPU_ERR_CUDA, "H2D of values failed: %s", cudaGetErrorString(e)); goto fail; }
  }
  // the caller's buffers may be reused as soon as we return
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = set_error(ctx, CBGPU_ERR_CUDA, "upload failed"); goto fail; }
  if (ctx->opt.validate_uploads && (rc = cbgpu_mat_validate(ctx, M)) != CBGPU_OK) goto fail;
  *out = (M);
  return CBGPU_OK;
fail:
  mat_release(ctx, M);
  return rc;
}

int cbgpu_mat_from_device_csc(cbgpu_ctx *ctx, int64_t m, int64_t n, int64_t nnz, const int64_t *colptr,
                              const int32_t *rows, const void *vals, int dtype, cbgpu_mat **out) {
  if (!ctx || !out || dtype_size(dtype) == 0) return CBGPU_ERR_INVALID;
  cbgpu_mat_impl *M = nullptr;
  CB_TRY(mat_alloc(ctx, m, n, nnz, -1, dtype, &M));
  if (nnz > 0) {
    CB_CUDA(ctx, cudaMemcpyAsync(M->ir, rows, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CB_CUDA(ctx, cudaMemcpyAsync(M->numx, vals, (size_t)nnz * dtype_size(dtype), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  CB_TRY(dev_alloc_t(ctx, &M->colptr, (size_t)n + 1));
  CB_CUDA(ctx, cudaMemcpyAsync(M->colptr, colptr, (size_t)(n + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  int rc = compact_columns(ctx, nullptr, M->colptr, n, &M->jc, &M->cp, &M->nzc);
  if (rc != CBGPU_OK) { mat_release(ctx, M); return rc; }
  *out = (M);
  return CBGPU_OK;
}

int cbgpu_mat_info(const cbgpu_mat *M, cbgpu_mat_info_t *info) {
  if (!M || !info) return CBGPU_ERR_INVALID;
  info->m = M->m; info->n = M->n; info->nnz = M->nnz; info->nzc = M->nzc; info->dtype = M->dtype;
  info->device_bytes = M->nnz * (4 + (int64_t)dtype_size(M->dtype)) + (2 * M->nzc + 1) * 8 + (M->colptr ? (M->n + 1) * 8 : 0);
  return CBGPU_OK;
}

int cbgpu_mat_download(cbgpu_ctx *ctx, const cbgpu_mat *M, const cbgpu_dcsc_out *h) {
  if (!ctx || !M || !h) return CBGPU_ERR_INVALID;
  if (h->idx_bytes != 4 && h->idx_bytes != 8) return set_error(ctx, CBGPU_ERR_INVALID, "idx_bytes must be 4 or 8");
  if (h->idx_bytes == 4 && (M->nnz > 0x7FFFFFFFLL || M->n > 0x7FFFFFFFLL))
    return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "block with %lld entries does not fit 32-bit column pointers", (long long)M->nnz);
  CB_TRY(download_index_array(ctx, h->jc, h->idx_bytes, M->nzc, M->jc, nullptr));
  CB_TRY(download_index_array(ctx, h->cp, h->idx_bytes, M->nzc + 1, M->cp, nullptr));
  CB_TRY(download_index_array(ctx, h->ir, h->idx_bytes, M->nnz, nullptr, M->ir));
  if (M->nnz > 0 && h->numx)
    CB_CUDA(ctx, cudaMemcpyAsync(h->numx, M->numx, (size_t)M->nnz * dtype_size(M->dtype), cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CBGPU_OK;
}

int cbgpu_mat_download_coo(cbgpu_ctx *ctx, const cbgpu_mat *M, void *rows, void *cols, void *vals, int idx_bytes) {
  if (!ctx || !M) return CBGPU_ERR_INVALID;
  if (idx_bytes != 4 && idx_bytes != 8) return set_error(ctx, CBGPU_ERR_INVALID, "idx_bytes must be 4 or 8");
  if (M->nnz == 0) return CBGPU_OK;
  void *dr = nullptr, *dc = nullptr;
  CB_TRY(dev_alloc(ctx, &dr, (size_t)M->nnz * idx_bytes));
  CB_TRY(dev_alloc(ctx, &dc, (size_t)M->nnz * idx_bytes));
  unsigned nb = nblocks(M->nzc * 32);
  if (idx_bytes == 8) expand_coo_kernel<int64_t><<<nb, 256, 0, ctx->stream>>>(M->jc, M->cp, M->ir, M->nzc, (int64_t *)dr, (int64_t *)dc);
  else expand_coo_kernel<int32_t><<<nb, 256, 0, ctx->stream>>>(M->jc, M->cp, M->ir, M->nzc, (int32_t *)dr, (int32_t *)dc);
  CB_LAUNCH_CHECK(ctx);
  CB_CUDA(ctx, cudaMemcpyAsync(rows, dr, (size_t)M->nnz * idx_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaMemcpyAsync(cols, dc, (size_t)M->nnz * idx_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaMemcpyAsync(vals, M->numx, (size_t)M->nnz * dtype_size(M->dtype), cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_free(ctx, dr));
  CB_TRY(dev_free(ctx, dc));
  return CBGPU_OK;
}

int cbgpu_mat_device_arrays(const cbgpu_mat *M, const int64_t **jc, const int64_t **cp, const int32_t **ir, const void **numx) {
  if (!M) return CBGPU_ERR_INVALID;
  if (jc) *jc = M->jc;
  if (cp) *cp = M->cp;
  if (ir) *ir = M->ir;
  if (numx) *numx = M->numx;
  return CBGPU_OK;
}

int cbgpu_mat_free(cbgpu_ctx *ctx, cbgpu_mat *M) {
  if (!ctx) return CBGPU_ERR_INVALID;
  return mat_release(ctx, M);
}

int cbgpu_mat_checksum(cbgpu_ctx *ctx, const cbgpu_mat *M, uint64_t *pattern_sum, uint64_t *value_sum) {
  return cbgpu_mat_checksum_at(ctx, M, 0, 0, pattern_sum, value_sum);
}

int cbgpu_mat_checksum_at(cbgpu_ctx *ctx, const cbgpu_mat *M, int64_t row_offset, int64_t col_offset, uint64_t *pattern_sum,
                          uint64_t *value_sum) {
  if (!ctx || !M || row_offset < 0 || col_offset < 0 || row_offset + M->m > ((int64_t)1 << 32)) return CBGPU_ERR_INVALID;
  unsigned long long *d = nullptr, h[2] = {0, 0};
  CB_TRY(dev_alloc_t(ctx, &d, 2));
  CB_CUDA(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
  if (M->nzc > 0) {
    checksum_kernel<<<nblocks(M->nzc * 32), 256, 0, ctx->stream>>>(M->jc, M->cp, M->ir, (const unsigned char *)M->numx,
                                                                   (int)dtype_size(M->dtype), M->nzc, row_offset, col_offset, d);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_CUDA(ctx, cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_free(ctx, d));
  if (pattern_sum) *pattern_sum = h[0];
  if (value_sum) *value_sum = h[1];
  return CBGPU_OK;
}

int cbgpu_mat_colslice(cbgpu_ctx *ctx, const cbgpu_mat *M, int64_t c0, int64_t c1, cbgpu_mat **out) {
  if (!ctx || !M || !out) return CBGPU_ERR_INVALID;
  cbgpu_mat_impl *S = nullptr;
  CB_TRY(mat_colslice(ctx, M, c0, c1, &S));
  *out = (S);
  return CBGPU_OK;
}

int cbgpu_mat_submatrix(cbgpu_ctx *ctx, const cbgpu_mat *M, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cbgpu_mat **out) {
  if (!ctx || !M || !out) return CBGPU_ERR_INVALID;
  cbgpu_mat_impl *S = nullptr;
  CB_TRY(mat_submatrix(ctx, M, r0, r1, c0, c1, &S));
  *out = S;
  return CBGPU_OK;
}

/* ColSplit (dcsc.cpp:1202-1277): `parts` contiguous column ranges of floor(n/parts) columns, the last one takes the rest */
int cbgpu_mat_colsplit(cbgpu_ctx *ctx, const cbgpu_mat *M, int parts, cbgpu_mat **out) {
  if (!ctx || !M || !out || parts < 1) return CBGPU_ERR_INVALID;
  int64_t per = M->n / parts;
  for (int i = 0; i < parts; ++i) {
    int64_t c0 = per * i, c1 = (i == parts - 1) ? M->n : per * (i + 1);
    cbgpu_mat_impl *S = nullptr;
    int rc = mat_colslice(ctx, M, c0, c1, &S);
    if (rc != CBGPU_OK) {
      for (int j = 0; j < i; ++j) mat_release(ctx, out[j]);
      return rc;
    }
    out[i] = (S);
  }
  return CBGPU_OK;
}

int cbgpu_mat_colconcat(cbgpu_ctx *ctx, int parts, cbgpu_mat *const *in, cbgpu_mat **out) {
  if (!ctx || !in || !out) return CBGPU_ERR_INVALID;
  std::vector<cbgpu_mat_impl *> v(in, in + parts);
  cbgpu_mat_impl *S = nullptr;
  CB_TRY(mat_colconcat(ctx, parts, v.data(), &S));
  *out = (S);
  return CBGPU_OK;
}

static int check_operands(cbgpu_ctx *ctx, int semiring, const cbgpu_mat *A, const cbgpu_mat *B) {
  int ta, tb, tc;
  if (semiring_types(semiring, &ta, &tb, &tc) != CBGPU_OK) return set_error(ctx, CBGPU_ERR_INVALID, "unknown semiring %d", semiring);
  if (A->n != B->m)
    return set_error(ctx, CBGPU_ERR_DIMMISMATCH, "Can not multiply, dimensions does not match: %lld != %lld", (long long)A->n, (long long)B->m);
  if (A->dtype != ta || B->dtype != tb)
    return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "operand value types (%d,%d) do not match semiring %d (%d,%d)", A->dtype, B->dtype, semiring, ta, tb);
  return CBGPU_OK;
}

int cbgpu_spgemm_local(cbgpu_ctx *ctx, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, cbgpu_mat **C, cbgpu_stats *stats) {
  if (!ctx || !A || !B || !C) return CBGPU_ERR_INVALID;
  CB_TRY(check_operands(ctx, semiring, A, B));
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_mat_impl *Cm = nullptr;
  SpgemmArgs a{ctx, const_cast<cbgpu_mat *>(A), const_cast<cbgpu_mat *>(B), &Cm, stats, nullptr, nullptr};
  CB_TRY(spgemm_entry(semiring)(a));
  *C = (Cm);
  return CBGPU_OK;
}

int cbgpu_semiring_types(int semiring, int *a_dtype, int *b_dtype, int *c_dtype) {
  int ta, tb, tc;
  if (semiring_types(semiring, &ta, &tb, &tc) != CBGPU_OK) return CBGPU_ERR_INVALID;
  if (a_dtype) *a_dtype = ta;
  if (b_dtype) *b_dtype = tb;
  if (c_dtype) *c_dtype = tc;
  return CBGPU_OK;
}

int cbgpu_spgemm_symbolic(cbgpu_ctx *ctx, const cbgpu_mat *A, const cbgpu_mat *B, int64_t *flops, int64_t *nnz_out) {
  return cbgpu_spgemm_symbolic_columns(ctx, A, B, flops, nnz_out, nullptr, nullptr);
}

int cbgpu_spgemm_symbolic_columns(cbgpu_ctx *ctx, const cbgpu_mat *A, const cbgpu_mat *B, int64_t *flops, int64_t *nnz_out,
                                  int64_t *col_flops, int64_t *col_nnz) {
  if (!ctx || !A || !B) return CBGPU_ERR_INVALID;
  if (A->n != B->m) return set_error(ctx, CBGPU_ERR_DIMMISMATCH, "dimensions do not match");
  // the pattern does not depend on the semiring: pick the instance whose operand types match in size
  int sr = -1;
  for (int s = 0; s < CBGPU_SR_COUNT && sr < 0; ++s) {
    int ta, tb, tc;
    semiring_types(s, &ta, &tb, &tc);
    if (ta == A->dtype && tb == B->dtype) sr = s;
  }
  if (sr < 0) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "no semiring instance for operand types (%d,%d)", A->dtype, B->dtype);
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  SpgemmArgs a{ctx, const_cast<cbgpu_mat *>(A), const_cast<cbgpu_mat *>(B), nullptr, nullptr, flops, nnz_out};
  a.col_flops_host = col_flops;
  a.col_nnz_host = col_nnz;
  if ((col_flops || col_nnz) && (A->nnz == 0 || B->nnz == 0)) { // isZero() operands: every column is empty
    for (int64_t j = 0; j < B->nzc; ++j) {
      if (col_flops) col_flops[j] = 0;
      if (col_nnz) col_nnz[j] = 0;
    }
  }
  return spgemm_entry(sr)(a);
}

int cbgpu_spgemm_local_host(cbgpu_ctx *ctx, int semiring, const cbgpu_dcsc_view *A, const cbgpu_dcsc_view *B, cbgpu_mat **C,
                            cbgpu_stats *stats) {
  if (!ctx || !A || !B || !C) return CBGPU_ERR_INVALID;
  cbgpu_mat *dA = nullptr, *dB = nullptr;
  CB_TRY(cbgpu_mat_upload(ctx, A, &dA));
  int rc = cbgpu_mat_upload(ctx, B, &dB);
  if (rc == CBGPU_OK) rc = cbgpu_spgemm_local(ctx, semiring, dA, dB, C, stats);
  cbgpu_mat_free(ctx, dA);
  cbgpu_mat_free(ctx, dB);
  return rc;
}

int cbgpu_merge(cbgpu_ctx *ctx, int semiring, int k, const cbgpu_mat *const *lists, cbgpu_mat **out, cbgpu_stats *stats) {
  if (!ctx || !lists || !out || k < 1) return CBGPU_ERR_INVALID;
  int ta, tb, tc;
  if (semiring_types(semiring, &ta, &tb, &tc) != CBGPU_OK) return set_error(ctx, CBGPU_ERR_INVALID, "unknown semiring %d", semiring);
  std::vector<cbgpu_mat_impl *> v;
  for (int i = 0; i < k; ++i) {
    if (!lists[i]) return CBGPU_ERR_INVALID;
    if (lists[i]->m != lists[0]->m || lists[i]->n != lists[0]->n)
      return set_error(ctx, CBGPU_ERR_DIMMISMATCH, "merge: list %d has a different shape", i);
    if (lists[i]->dtype != tc) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "merge: list %d has value type %d, semiring output is %d", i, lists[i]->dtype, tc);
    v.push_back(const_cast<cbgpu_mat *>(lists[i]));
  }
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_mat_impl *Cm = nullptr;
  MergeArgs a{ctx, k, v.data(), &Cm, stats};
  CB_TRY(merge_entry(semiring)(a));
  *out = (Cm);
  return CBGPU_OK;
}

} // extern "C"
