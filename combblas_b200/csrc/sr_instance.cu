// One translation unit per semiring (-DCB_SR=<id>): instantiates the accumulation engine for the local
// multiply (LocalHybridSpGEMM / LocalSpGEMMHash, mtSpGEMM.h:213,:463) and the k-way merge
// (MultiwayMerge / MultiwayMergeHash, MultiwayMerge.h:428,:553) of that semiring.
#include "engine_host.cuh"

#ifndef CB_SR
#error "compile with -DCB_SR=<semiring id>"
#endif

namespace cbgpu {

typedef Semiring<CB_SR> SR;

static int spgemm_impl(const SpgemmArgs &a) {
  cbgpu_ctx_impl *ctx = a.ctx;
  cbgpu_mat_impl *A = a.A, *B = a.B;
  CB_TRY(ensure_dense_colptr(ctx, A));
  int nwin, wlog2;
  engine_windows(ctx, A->m, &nwin, &wlog2);
  if (nwin > 1 && A->nnz > 0 && B->nnz > 0) CB_TRY(ensure_window_major(ctx, A, nwin, wlog2)); // cached on A
  Source<SR, false> src;
  memset(&src, 0, sizeof(src));
  src.T2 = A->win_T2;
  src.Wir = A->win_ir;
  src.Wval = reinterpret_cast<const SR::a_t *>(A->win_val);
  src.Air = A->ir;
  src.Aval = reinterpret_cast<const SR::a_t *>(A->numx);
  src.Bcp = B->cp;
  src.Bir = B->ir;
  src.Bval = reinterpret_cast<const SR::b_t *>(B->numx);
  src.k = 0;
  src.n = B->n;
  EngineIO io;
  memset(&io, 0, sizeof(io));
  io.Acolptr = A->colptr;
  io.ncolA = A->n;
  io.m = A->m;
  io.ncol = (A->nnz == 0 || B->nnz == 0) ? 0 : B->nzc; // isZero() operands give an empty product (mtSpGEMM.h:224-227)
  io.out_col_ids = B->jc;
  io.n_out = B->n;
  int ta, tb, tc;
  semiring_types(CB_SR, &ta, &tb, &tc);
  io.out_dtype = tc;
  io.C = a.C;
  io.stats = a.stats;
  io.flops_out = a.flops_out;
  io.nnz_out = a.nnz_out;
  return run_engine<SR, false>(ctx, src, io);
}

static __global__ void concat_colptr_kernel(const int64_t *colptr, int64_t n, int64_t offset, int64_t *out, bool last) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = colptr[j] + offset;
  if (last && j == n) out[n] = colptr[n] + offset;
}

static int merge_impl(const MergeArgs &a) {
  cbgpu_ctx_impl *ctx = a.ctx;
  const int k = a.k;
  typedef SR::out_t out_t;
  int64_t m = a.lists[0]->m, n = a.lists[0]->n, total = 0;
  for (int i = 0; i < k; ++i) total += a.lists[i]->nnz;
  // the k lists become one segment source: rows/values concatenated, column index of list s at [s*n, (s+1)*n)
  int32_t *rows = nullptr;
  out_t *vals = nullptr;
  int64_t *colptr = nullptr;
  CB_TRY(dev_alloc_t(ctx, &rows, (size_t)total));
  CB_TRY(dev_alloc_t(ctx, &vals, (size_t)total));
  CB_TRY(dev_alloc_t(ctx, &colptr, (size_t)k * n + 1));
  int64_t off = 0;
  for (int i = 0; i < k; ++i) {
    cbgpu_mat_impl *L = a.lists[i];
    CB_TRY(ensure_dense_colptr(ctx, L));
    if (L->nnz > 0) {
      CB_CUDA(ctx, cudaMemcpyAsync(rows + off, L->ir, sizeof(int32_t) * (size_t)L->nnz, cudaMemcpyDeviceToDevice, ctx->stream));
      CB_CUDA(ctx, cudaMemcpyAsync(vals + off, L->numx, sizeof(out_t) * (size_t)L->nnz, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    concat_colptr_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(L->colptr, n, off, colptr + (int64_t)i * n,
                                                                                 i == k - 1);
    CB_LAUNCH_CHECK(ctx);
    off += L->nnz;
  }
  Source<SR, true> src;
  memset(&src, 0, sizeof(src));
  int nwin, wlog2;
  engine_windows(ctx, m, &nwin, &wlog2);
  int64_t *T2 = nullptr;
  int32_t *Wir = nullptr;
  void *Wval = nullptr;
  (void)nwin; // merge segments are few and long: they are cut per window by binary search, no window-major copy
  src.T2 = T2;
  src.Wir = Wir;
  src.Wval = reinterpret_cast<const out_t *>(Wval);
  src.Air = rows;
  src.Aval = vals;
  src.k = k;
  src.n = n;
  EngineIO io;
  memset(&io, 0, sizeof(io));
  io.Acolptr = colptr;
  io.ncolA = (int64_t)k * n;
  io.m = m;
  io.ncol = total == 0 ? 0 : n;
  io.out_col_ids = nullptr;
  io.n_out = n;
  io.out_dtype = a.lists[0]->dtype;
  io.C = a.out;
  io.stats = a.stats;
  int rc = run_engine<SR, true>(ctx, src, io);
  dev_free(ctx, T2);
  dev_free(ctx, Wir);
  dev_free(ctx, Wval);
  dev_free(ctx, rows);
  dev_free(ctx, vals);
  dev_free(ctx, colptr);
  return rc;
}

#define CB_CAT2(a, b) a##b
#define CB_CAT(a, b) CB_CAT2(a, b)
int CB_CAT(spgemm_sr, CB_SR)(const SpgemmArgs &a) { return spgemm_impl(a); }
int CB_CAT(merge_sr, CB_SR)(const MergeArgs &a) { return merge_impl(a); }

} // namespace cbgpu
