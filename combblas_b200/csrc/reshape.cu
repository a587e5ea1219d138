// Operand reshaping on the device, next to the hot path:
//   Transpose                    SpDCCols::Transpose / TransposeConst (SpDCCols.cpp:871-905)  -> cbgpu_mat_transpose
//   2D <-> 3D redistribution     SpParMat3D(const SpParMat&, nlayers, colsplit, special) (SpParMat3D.cpp:187-283),
//                                SpParMat3D::Convert2D (:441-570), both built on ExchangeData (:51, :97)
//                                                                                           -> cbgpu_redistribute
// Neither is on the timed path of the multiply; the sort used by the transpose is CUB's (a library call, like the
// generator's). The redistribution moves DCSC blocks device to device with grouped ncclSend/ncclRecv.
#include <algorithm>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"
#include "util.cuh"

namespace cbgpu {
int mat_submatrix(cbgpu_ctx_impl *ctx, const cbgpu_mat_impl *M, int64_t r0, int64_t r1, int64_t c0, int64_t c1, cbgpu_mat_impl **out);
int mat_colconcat(cbgpu_ctx_impl *ctx, int parts, cbgpu_mat_impl *const *in, cbgpu_mat_impl **out);
int mat_rowstack(cbgpu_ctx_impl *ctx, int parts, cbgpu_mat_impl *const *in, cbgpu_mat_impl **out);
int exchange_blocks_world(cbgpu_ctx *ctx, cbgpu_comm *comm, std::vector<cbgpu_mat *> &send, std::vector<cbgpu_mat *> &recv,
                          int64_t *bytes); // dist.cu

// key = row << 32 | column for every entry (one warp per non-empty column), payload = position
static __global__ void transpose_keys_kernel(const int64_t *jc, const int64_t *cp, const int32_t *ir, int64_t nzc, uint64_t *keys,
                                             uint32_t *pos) {
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nzc) return;
  const uint64_t col = (uint64_t)jc[c];
  for (int64_t p = cp[c] + (threadIdx.x & 31); p < cp[c + 1]; p += 32) {
    keys[p] = ((uint64_t)(uint32_t)ir[p] << 32) | col;
    pos[p] = (uint32_t)p;
  }
}
template <int VB>
static __global__ void transpose_emit_kernel(const uint64_t *keys, const uint32_t *pos, const unsigned char *vals, int64_t nnz,
                                             int32_t *rows, unsigned char *ovals, int64_t *colcount) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nnz) return;
  const uint64_t k = keys[q];
  rows[q] = (int32_t)(k & 0xFFFFFFFFu); // the old column id is the new row id
  const uint32_t p = pos[q];
  if (VB == 8) reinterpret_cast<uint64_t *>(ovals)[q] = reinterpret_cast<const uint64_t *>(vals)[p];
  else if (VB == 4) reinterpret_cast<uint32_t *>(ovals)[q] = reinterpret_cast<const uint32_t *>(vals)[p];
  else ovals[q] = vals[p];
  atomicAdd((unsigned long long *)&colcount[k >> 32], 1ull);
}

} // namespace cbgpu

using namespace cbgpu;

extern "C" int cbgpu_mat_transpose(cbgpu_ctx *ctx, const cbgpu_mat *M, cbgpu_mat **out) {
  if (!ctx || !M || !out) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (M->nnz >= ((int64_t)1 << 32) || M->n >= ((int64_t)1 << 31) - 1)
    return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "transpose: at most 2^32 entries and 2^31 columns per block");
  Scratch scratch(ctx);
  cbgpu_mat *T = nullptr;
  MatGuard guard(ctx, &T);
  CB_TRY(mat_alloc(ctx, M->n, M->m, M->nnz, -1, M->dtype, &T));
  int64_t *colcount = nullptr;
  CB_TRY(scratch.alloc(&colcount, (size_t)M->m + 1));
  CB_CUDA(ctx, cudaMemsetAsync(colcount, 0, ((size_t)M->m + 1) * 8, ctx->stream));
  if (M->nnz > 0) {
    uint64_t *keys = nullptr, *keys2 = nullptr;
    uint32_t *pos = nullptr, *pos2 = nullptr;
    CB_TRY(scratch.alloc(&keys, (size_t)M->nnz));
    CB_TRY(scratch.alloc(&keys2, (size_t)M->nnz));
    CB_TRY(scratch.alloc(&pos, (size_t)M->nnz));
    CB_TRY(scratch.alloc(&pos2, (size_t)M->nnz));
    transpose_keys_kernel<<<(unsigned)((M->nzc * 32 + 255) / 256), 256, 0, ctx->stream>>>(M->jc, M->cp, M->ir, M->nzc, keys, pos);
    CB_LAUNCH_CHECK(ctx);
    size_t tmp_bytes = 0;
    CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, pos, pos2, M->nnz, 0, 64, ctx->stream));
    unsigned char *tmp = nullptr;
    CB_TRY(scratch.alloc(&tmp, tmp_bytes));
    CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, pos, pos2, M->nnz, 0, 64, ctx->stream));
    ctx->launches += 4;
    const unsigned nb = (unsigned)((M->nnz + 255) / 256);
    const unsigned char *v = (const unsigned char *)M->numx;
    unsigned char *ov = (unsigned char *)T->numx;
    const int vb = (int)dtype_size(M->dtype);
    if (vb == 8) transpose_emit_kernel<8><<<nb, 256, 0, ctx->stream>>>(keys2, pos2, v, M->nnz, T->ir, ov, colcount);
    else if (vb == 4) transpose_emit_kernel<4><<<nb, 256, 0, ctx->stream>>>(keys2, pos2, v, M->nnz, T->ir, ov, colcount);
    else transpose_emit_kernel<1><<<nb, 256, 0, ctx->stream>>>(keys2, pos2, v, M->nnz, T->ir, ov, colcount);
    CB_LAUNCH_CHECK(ctx);
  }
  CB_TRY(dev_alloc_t(ctx, &T->colptr, (size_t)M->m + 1));
  CB_TRY(exclusive_scan_i64(ctx, colcount, T->colptr, M->m));
  CB_TRY(compact_columns(ctx, nullptr, T->colptr, M->m, &T->jc, &T->cp, &T->nzc));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = T;
  guard.armed = false;
  return CBGPU_OK;
}

// Every rank holds the block `local` at global position source[rank] = {r0, r1, c0, c1} and wants target[rank]; source and
// target rectangles of ALL ranks are known to everyone (they follow from the grids' owner arithmetic). The sources tile the
// matrix as a product of a row partition and a column partition (true for the 2D layout and both 3D layouts), so what a rank
// receives tiles its target the same way: pieces with equal column range are stacked by rows, the stacks are concatenated.
extern "C" int cbgpu_redistribute(cbgpu_ctx *ctx, cbgpu_comm *comm, const cbgpu_mat *local, const int64_t *source,
                                  const int64_t *target, int world, int rank, cbgpu_mat **out, int64_t *bytes_moved) {
  if (!ctx || !comm || !local || !source || !target || !out || world < 1 || rank < 0 || rank >= world) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t *S = source + 4 * rank, *T = target + 4 * rank;
  if (local->m != S[1] - S[0] || local->n != S[3] - S[2])
    return set_error(ctx, CBGPU_ERR_DIMMISMATCH, "redistribute: the local block is %lld x %lld, its source rectangle %lld x %lld",
                     (long long)local->m, (long long)local->n, (long long)(S[1] - S[0]), (long long)(S[3] - S[2]));
  std::vector<cbgpu_mat *> send(world, nullptr), recv(world, nullptr);
  int rc = CBGPU_OK;
  auto cleanup = [&]() {
    for (cbgpu_mat *&m : send) { mat_release(ctx, m); m = nullptr; }
    for (cbgpu_mat *&m : recv) { mat_release(ctx, m); m = nullptr; }
  };
  // what I own of every rank's target (an empty block when the rectangles do not meet)
  for (int d = 0; d < world && rc == CBGPU_OK; ++d) {
    const int64_t *D = target + 4 * d;
    int64_t r0 = std::max(S[0], D[0]), r1 = std::min(S[1], D[1]), c0 = std::max(S[2], D[2]), c1 = std::min(S[3], D[3]);
    if (r1 <= r0 || c1 <= c0) r0 = r1 = S[0], c0 = c1 = S[2];
    rc = mat_submatrix(ctx, local, r0 - S[0], r1 - S[0], c0 - S[2], c1 - S[2], &send[d]);
  }
  int64_t bytes = 0;
  if (rc == CBGPU_OK && world > 1) rc = exchange_blocks_world(ctx, comm, send, recv, &bytes);
  if (rc != CBGPU_OK) {
    cleanup();
    return rc;
  }
  // assemble: pieces by (column range, row start)
  struct Piece { int64_t r0, r1, c0, c1; cbgpu_mat *m; };
  std::vector<Piece> pieces;
  for (int s = 0; s < world; ++s) {
    const int64_t *Q = source + 4 * s;
    const int64_t r0 = std::max(Q[0], T[0]), r1 = std::min(Q[1], T[1]), c0 = std::max(Q[2], T[2]), c1 = std::min(Q[3], T[3]);
    if (r1 <= r0 || c1 <= c0) continue;
    pieces.push_back(Piece{r0, r1, c0, c1, s == rank ? send[s] : recv[s]});
  }
  std::sort(pieces.begin(), pieces.end(), [](const Piece &a, const Piece &b) { return a.c0 != b.c0 ? a.c0 < b.c0 : a.r0 < b.r0; });
  std::vector<cbgpu_mat *> stacks;
  auto release_stacks = [&]() {
    for (cbgpu_mat *m : stacks) mat_release(ctx, m);
  };
  int64_t col_at = T[2];
  for (size_t i = 0; i < pieces.size() && rc == CBGPU_OK;) {
    size_t j = i;
    std::vector<cbgpu_mat *> col;
    int64_t row_at = T[0];
    while (j < pieces.size() && pieces[j].c0 == pieces[i].c0) {
      if (pieces[j].c1 != pieces[i].c1 || pieces[j].r0 != row_at)
        rc = set_error(ctx, CBGPU_ERR_GRID, "redistribute: the source rectangles do not tile the target as rows x columns");
      row_at = pieces[j].r1;
      col.push_back(pieces[j].m);
      ++j;
    }
    if (rc == CBGPU_OK && (row_at != T[1] || pieces[i].c0 != col_at))
      rc = set_error(ctx, CBGPU_ERR_GRID, "redistribute: the target rectangle is not covered by the sources");
    if (rc == CBGPU_OK) {
      cbgpu_mat *st = nullptr;
      rc = mat_rowstack(ctx, (int)col.size(), col.data(), &st);
      if (rc == CBGPU_OK) stacks.push_back(st);
      col_at = pieces[i].c1;
    }
    i = j;
  }
  if (rc == CBGPU_OK && col_at != T[3] && !(T[1] == T[0] || T[3] == T[2]))
    rc = set_error(ctx, CBGPU_ERR_GRID, "redistribute: the target rectangle is not covered by the sources");
  cbgpu_mat *res = nullptr;
  if (rc == CBGPU_OK) {
    if (stacks.empty()) { // an empty target rectangle
      rc = mat_alloc(ctx, T[1] - T[0], T[3] - T[2], 0, 0, local->dtype, &res);
      if (rc == CBGPU_OK) CB_CUDA(ctx, cudaMemsetAsync(res->cp, 0, 8, ctx->stream));
    } else {
      rc = mat_colconcat(ctx, (int)stacks.size(), stacks.data(), &res);
    }
  }
  release_stacks();
  cleanup();
  if (rc != CBGPU_OK) return rc;
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bytes_moved) *bytes_moved = bytes;
  *out = res;
  return CBGPU_OK;
}
