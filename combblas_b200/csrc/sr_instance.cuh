// The per-semiring entry points of the accumulation engine as templates: the local multiply (LocalHybridSpGEMM /
// LocalSpGEMMHash, mtSpGEMM.h:213,:463) and the k-way merge (MultiwayMerge / MultiwayMergeHash, MultiwayMerge.h:428,:553).
// Instantiated once per library semiring by sr_instance.cu and once per user-defined semiring by
// include/combblas_b200/device_semiring.cuh (a translation unit of the application, compiled with nvcc).
#pragma once
#include "engine_host.cuh"
#include "merge2.cuh"

namespace cbgpu {

template <class SR>
int spgemm_impl(const SpgemmArgs &a, int out_dtype) {
  cbgpu_ctx_impl *ctx = a.ctx;
  cbgpu_mat_impl *A = a.A, *B = a.B;
  CB_TRY(ensure_dense_colptr(ctx, A));
  int nwin, wlog2;
  engine_windows(ctx, A->m, &nwin, &wlog2);
  // the window-major copy (16-byte aligned pieces) is what the bitmap kernels read, whatever the number of windows; cached on A
  if (A->nnz > 0 && B->nnz > 0) CB_TRY(ensure_window_major(ctx, A, nwin, wlog2));
  Source<SR, false> src;
  memset(&src, 0, sizeof(src));
  src.T2 = A->win_T2;
  src.Wir = A->win_ir;
  src.Wval = reinterpret_cast<const typename SR::a_t *>(A->win_val);
  src.Air = A->ir;
  src.Aval = reinterpret_cast<const typename SR::a_t *>(A->numx);
  src.Bcp = B->cp;
  src.Bir = B->ir;
  src.Bval = reinterpret_cast<const typename SR::b_t *>(B->numx);
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
  io.out_dtype = out_dtype;
  io.C = a.C;
  io.stats = a.stats;
  io.flops_out = a.flops_out;
  io.nnz_out = a.nnz_out;
  io.col_flops_host = a.col_flops_host;
  io.col_nnz_host = a.col_nnz_host;
  return run_engine<SR, false>(ctx, src, io);
}

static __global__ void concat_colptr_kernel(const int64_t *colptr, int64_t n, int64_t offset, int64_t *out, bool last) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = colptr[j] + offset;
  if (last && j == n) out[n] = colptr[n] + offset;
}

// k-way merge by rounds of streaming 2-way merges (k = 2 is the case 2x2 SUMMA and 2-layer fibers produce);
// the accumulation engine remains available as the general k-way path (option "merge_engine" = 1).
template <class SR>
int merge_engine_impl(const MergeArgs &a);
template <class SR>
int merge_impl(const MergeArgs &a) {
  cbgpu_ctx_impl *ctx = a.ctx;
  if (ctx->opt.merge_engine || a.k < 2) return merge_engine_impl<SR>(a);
  std::vector<cbgpu_mat_impl *> cur(a.lists, a.lists + a.k);
  std::vector<bool> owned(a.k, false);
  cbgpu_stats total;
  memset(&total, 0, sizeof(total));
  while (cur.size() > 1) {
    std::vector<cbgpu_mat_impl *> next;
    std::vector<bool> next_owned;
    for (size_t i = 0; i + 1 < cur.size(); i += 2) {
      cbgpu_mat_impl *C = nullptr;
      cbgpu_stats st;
      int rc = merge2_run<SR>(ctx, cur[i], cur[i + 1], &C, &st);
      if (owned[i]) mat_release(ctx, cur[i]);
      if (owned[i + 1]) mat_release(ctx, cur[i + 1]);
      if (rc != CBGPU_OK) {
        for (size_t q = i + 2; q < cur.size(); ++q) if (owned[q]) mat_release(ctx, cur[q]);
        for (size_t q = 0; q < next.size(); ++q) if (next_owned[q]) mat_release(ctx, next[q]);
        return rc;
      }
      total.flops += st.flops; total.tasks += st.tasks; total.kernel_launches += st.kernel_launches;
      total.ms_setup += st.ms_setup; total.ms_symbolic += st.ms_symbolic; total.ms_numeric += st.ms_numeric; total.ms_total += st.ms_total;
      next.push_back(C);
      next_owned.push_back(true);
    }
    if (cur.size() & 1) {
      next.push_back(cur.back());
      next_owned.push_back(owned.back());
    }
    cur.swap(next);
    owned.swap(next_owned);
  }
  total.nnz_out = cur[0]->nnz;
  total.nzc_out = cur[0]->nzc;
  if (a.stats) *a.stats = total;
  *a.out = cur[0];
  return CBGPU_OK;
}

template <class SR>
int merge_engine_impl(const MergeArgs &a) {
  cbgpu_ctx_impl *ctx = a.ctx;
  const int k = a.k;
  typedef typename SR::out_t out_t;
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

} // namespace cbgpu
