// Distributed semiring SpGEMM: process-grid arithmetic (host only) and the 2D / 3D Sparse SUMMA drivers with
// NCCL collectives on device-resident DCSC blocks.
//
//   Mult_AnXBn_Synch      ParFriends.h:1447-1556  -> cbgpu_summa2d
//   Mult_AnXBn_SUMMA3D    ParFriends.h:3374-3667  -> cbgpu_summa3d
//   GetSetSizes/BCastMatrix SpParHelper.cpp:798,:583 -> one ncclAllGather of the essentials + grouped ncclBroadcast
//   fiber Alltoallv       ParFriends.h:3578-3612  -> grouped ncclSend/ncclRecv of column slabs
//   CommGrid / CommGrid3D rank maps: src/CommGrid.cpp:57-58, CommGrid3D.h:75-93
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- inside a torch process this is torch's own copy), so
// the library loads on machines without NCCL and single-GPU use never touches it.
#include <condition_variable>
#include <dlfcn.h>
#include <mutex>
#include <thread>
#include <math.h>
#include <nccl.h>
#include "common.cuh"
#include "util.cuh"

namespace cbgpu {
int mat_colslice(cbgpu_ctx_impl *ctx, const cbgpu_mat_impl *M, int64_t c0, int64_t c1, cbgpu_mat_impl **out);
int mat_colconcat(cbgpu_ctx_impl *ctx, int parts, cbgpu_mat_impl *const *in, cbgpu_mat_impl **out);
int mat_rowstack(cbgpu_ctx_impl *ctx, int parts, cbgpu_mat_impl *const *in, cbgpu_mat_impl **out);

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
#define LOAD(field, sym)                                                                                               \
  *(void **)(&api.field) = dlsym(api.handle, sym);                                                                     \
  if (!api.field) return api;
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommSplit, "ncclCommSplit")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(Broadcast, "ncclBroadcast")
  LOAD(AllGather, "ncclAllGather")
  LOAD(Send, "ncclSend")
  LOAD(Recv, "ncclRecv")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  api.ok = true;
  return api;
}

#define CB_NCCL(ctx, expr)                                                                                             \
  do {                                                                                                                 \
    ncclResult_t _r = (expr);                                                                                          \
    if (_r != ncclSuccess)                                                                                             \
      return set_error((ctx), CBGPU_ERR_NCCL, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                     \
                       nccl().GetErrorString ? nccl().GetErrorString(_r) : "?");                                       \
  } while (0)

} // namespace cbgpu

using namespace cbgpu;

struct cbgpu_comm {
  cbgpu_grid grid;
  ncclComm_t world = nullptr, row = nullptr, col = nullptr, fiber = nullptr;
  cbgpu_ctx *ctx2 = nullptr; // second context (own stream) for the fiber stage of the pipelined phased driver
};

extern "C" {

// ---------------------------------------------------------------------------------------------- grid arithmetic
int cbgpu_grid_make(int world, int rank, int layers, cbgpu_grid *g) {
  if (!g || world < 1 || rank < 0 || rank >= world || layers < 1 || world % layers != 0) return CBGPU_ERR_GRID;
  int per_layer = world / layers;
  int pr = (int)floor(sqrt((double)per_layer) + 0.5);
  if (pr * pr != per_layer) return CBGPU_ERR_GRID; // NOTSQUARE (src/CommGrid.cpp:44-54, CommGrid3D.h:51-58)
  g->world = world;
  g->rank = rank;
  g->layers = layers;
  g->grid_rows = pr;
  g->grid_cols = pr;
  g->my_layer = rank / per_layer;       // CommGrid3D.h:75
  int in_layer = rank % per_layer;      // CommGrid3D.h:76
  g->my_row = in_layer / pr;            // src/CommGrid.cpp:57
  g->my_col = in_layer % pr;            // src/CommGrid.cpp:58
  return CBGPU_OK;
}

// the rank map of the older 3D code (3DSpGEMM/CCGrid.h:14-17): the layer index runs fastest
int cbgpu_grid_make_ccgrid(int world, int rank, int layers, cbgpu_grid *g) {
  int rc = cbgpu_grid_make(world, rank, layers, g);
  if (rc != CBGPU_OK) return rc;
  g->my_layer = rank % layers;          // layer_grid  = myrank % c_factor
  int in_layer = rank / layers;         // RankInLayer = myrank / c_factor
  g->my_row = in_layer / g->grid_cols;  // RankInCol   = RankInLayer / GridCols
  g->my_col = in_layer % g->grid_cols;  // RankInRow   = RankInLayer % GridCols
  return CBGPU_OK;
}

int cbgpu_block_range(int64_t dim, int parts, int index, int64_t *begin, int64_t *end) {
  if (parts < 1 || index < 0 || index >= parts || dim < 0) return CBGPU_ERR_INVALID;
  int64_t per = dim / parts;
  *begin = per * index;
  *end = (index == parts - 1) ? dim : per * (index + 1);
  return CBGPU_OK;
}

int cbgpu_block_owner(int64_t dim, int parts, int64_t gi) {
  if (parts < 1 || gi < 0 || gi >= dim) return -1;
  int64_t per = dim / parts;
  if (per == 0) return parts - 1;
  int64_t o = gi / per;
  return (int)(o < parts - 1 ? o : parts - 1);
}

int cbgpu_grid_local_range(const cbgpu_grid *g, int64_t m, int64_t n, int split_cols, int64_t *r0, int64_t *r1,
                           int64_t *c0, int64_t *c1) {
  if (!g) return CBGPU_ERR_INVALID;
  int64_t rb, re, cb, ce;
  cbgpu_block_range(m, g->grid_rows, g->my_row, &rb, &re);
  cbgpu_block_range(n, g->grid_cols, g->my_col, &cb, &ce);
  if (g->layers > 1) {
    int64_t s0, s1;
    if (split_cols) { // A and C: columns of the 2D block cut into `layers` contiguous chunks (SpParMat3D.cpp:337-402)
      cbgpu_block_range(ce - cb, g->layers, g->my_layer, &s0, &s1);
      ce = cb + s1;
      cb = cb + s0;
    } else { // B: rows of the 2D block
      cbgpu_block_range(re - rb, g->layers, g->my_layer, &s0, &s1);
      re = rb + s1;
      rb = rb + s0;
    }
  }
  *r0 = rb; *r1 = re; *c0 = cb; *c1 = ce;
  return CBGPU_OK;
}

// ---------------------------------------------------------------------------------------------- communicators
int cbgpu_nccl_unique_id(void *id128) {
  if (!id128) return CBGPU_ERR_INVALID;
  if (!nccl().ok) return CBGPU_ERR_NCCL;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return CBGPU_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return CBGPU_OK;
}

int cbgpu_comm_create(cbgpu_ctx *ctx, const cbgpu_grid *grid, const void *id128, cbgpu_comm **out) {
  if (!ctx || !grid || !id128 || !out) return CBGPU_ERR_INVALID;
  if (!nccl().ok) return set_error(ctx, CBGPU_ERR_NCCL, "libnccl.so.2 could not be loaded");
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_comm *c = new cbgpu_comm();
  c->grid = *grid;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  const int pr = grid->grid_rows, pc = grid->grid_cols;
  ncclResult_t r = nccl().CommInitRank(&c->world, grid->world, id, grid->rank);
  // row world: same layer, same grid row, ordered by column  (src/CommGrid.cpp:66)
  if (r == ncclSuccess) r = nccl().CommSplit(c->world, grid->my_layer * pr + grid->my_row, grid->my_col, &c->row, nullptr);
  // column world: same layer, same grid column, ordered by row (src/CommGrid.cpp:67)
  if (r == ncclSuccess) r = nccl().CommSplit(c->world, grid->my_layer * pc + grid->my_col, grid->my_row, &c->col, nullptr);
  // fiber world: same position in every layer, ordered by layer (CommGrid3D.h:77-78)
  if (r == ncclSuccess) r = nccl().CommSplit(c->world, grid->my_row * pc + grid->my_col, grid->my_layer, &c->fiber, nullptr);
  if (r != ncclSuccess) {
    cbgpu_comm_destroy(c); // whatever was created so far
    return set_error(ctx, CBGPU_ERR_NCCL, "communicator setup failed: %s", nccl().GetErrorString(r));
  }
  *out = c;
  return CBGPU_OK;
}

int cbgpu_comm_destroy(cbgpu_comm *c) {
  if (!c) return CBGPU_OK;
  if (c->ctx2) cbgpu_destroy(c->ctx2);
  if (nccl().ok) {
    if (c->row) nccl().CommDestroy(c->row);
    if (c->col) nccl().CommDestroy(c->col);
    if (c->fiber) nccl().CommDestroy(c->fiber);
    if (c->world) nccl().CommDestroy(c->world);
  }
  delete c;
  return CBGPU_OK;
}

} // extern "C"

namespace cbgpu {

static ncclDataType_t nccl_bytes() { return ncclUint8; }

// essentials {nnz, nzc, m, n} of every block in a communicator (GetSetSizes, SpParHelper.cpp:798-809)
static int gather_essentials(cbgpu_ctx *ctx, ncclComm_t comm, int nranks, const cbgpu_mat *M, std::vector<int64_t> &ess) {
  int64_t mine[4] = {M->nnz, M->nzc, M->m, M->n};
  int64_t *d = nullptr;
  CB_TRY(dev_alloc_t(ctx, &d, (size_t)4 * (nranks + 1)));
  CB_CUDA(ctx, cudaMemcpyAsync(d, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  CB_NCCL(ctx, nccl().AllGather(d, d + 4, 4, ncclInt64, comm, ctx->stream));
  ess.resize((size_t)4 * nranks);
  CB_CUDA(ctx, cudaMemcpyAsync(ess.data(), d + 4, sizeof(int64_t) * 4 * nranks, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_free(ctx, d));
  return CBGPU_OK;
}

// BCastMatrix (SpParHelper.cpp:583-603): the four DCSC arrays of the root's block, device to device
static int bcast_block(cbgpu_ctx *ctx, ncclComm_t comm, int root, int me, const cbgpu_mat *own, const int64_t *ess,
                       int dtype, cbgpu_mat **recv, int64_t *bytes) {
  const int64_t nnz = ess[0], nzc = ess[1], m = ess[2], n = ess[3];
  cbgpu_mat *R = nullptr;
  const cbgpu_mat *src = own;
  if (me != root) {
    CB_TRY(mat_alloc(ctx, m, n, nnz, nzc, dtype, &R));
    src = R;
  }
  const size_t vb = dtype_size(dtype);
  if (nnz > 0) {
    ncclResult_t r = nccl().GroupStart();
    if (r == ncclSuccess) r = nccl().Broadcast(src->jc, src->jc, (size_t)nzc * 8, nccl_bytes(), root, comm, ctx->stream);
    if (r == ncclSuccess) r = nccl().Broadcast(src->cp, src->cp, (size_t)(nzc + 1) * 8, nccl_bytes(), root, comm, ctx->stream);
    if (r == ncclSuccess) r = nccl().Broadcast(src->ir, src->ir, (size_t)nnz * 4, nccl_bytes(), root, comm, ctx->stream);
    if (r == ncclSuccess) r = nccl().Broadcast(src->numx, src->numx, (size_t)nnz * vb, nccl_bytes(), root, comm, ctx->stream);
    const ncclResult_t r2 = nccl().GroupEnd(); // always closes the group, also after a failed call inside it
    if (r != ncclSuccess || r2 != ncclSuccess) {
      mat_release(ctx, R);
      return set_error(ctx, CBGPU_ERR_NCCL, "block broadcast failed: %s", nccl().GetErrorString(r != ncclSuccess ? r : r2));
    }
    *bytes += nzc * 16 + 8 + nnz * (4 + (int64_t)vb);
  } else if (R) {
    CB_CUDA(ctx, cudaMemsetAsync(R->cp, 0, 8, ctx->stream));
  }
  *recv = R;
  return CBGPU_OK;
}

static void add_stats(cbgpu_stats &acc, const cbgpu_stats &s) {
  acc.flops += s.flops; acc.tasks += s.tasks; acc.kernel_launches += s.kernel_launches;
  acc.ms_setup += s.ms_setup; acc.ms_symbolic += s.ms_symbolic; acc.ms_numeric += s.ms_numeric; acc.ms_total += s.ms_total;
  acc.tasks_hash_warp += s.tasks_hash_warp; acc.tasks_hash_cta += s.tasks_hash_cta;
  acc.tasks_bitmap_smem += s.tasks_bitmap_smem; acc.tasks_bitmap_gmem += s.tasks_bitmap_gmem;
  acc.flops_hash_warp += s.flops_hash_warp; acc.flops_hash_cta += s.flops_hash_cta;
  acc.flops_bitmap_smem += s.flops_bitmap_smem; acc.flops_bitmap_gmem += s.flops_bitmap_gmem;
  acc.nnz_out += s.nnz_out;
  for (int i = 0; i < CBGPU_K_COUNT; ++i) {
    acc.ms_kernel[i] += s.ms_kernel[i];
    acc.class_tasks[i] += s.class_tasks[i];
    acc.class_flops[i] += s.class_flops[i];
    acc.class_nnz[i] += s.class_nnz[i];
  }
}

struct Timer {
  cudaEvent_t a, b;
  cudaStream_t s;
  Timer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
  void start() { cudaEventRecord(a, s); }
  float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

// the SUMMA loop of one layer (ParFriends.h:1482-1532): stage i multiplies A(:, i-th block) by B(i-th block, :)
static int summa_layer(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, cbgpu_mat **C,
                       cbgpu_dist_stats *ds) {
  const cbgpu_grid &g = comm->grid;
  const int stages = g.grid_cols;
  int ta, tb, tc;
  CB_TRY(semiring_types(semiring, &ta, &tb, &tc));
  if (A->dtype != ta || B->dtype != tb) return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "operand types do not match the semiring");
  std::vector<int64_t> essA, essB;
  Timer tm(ctx->stream);
  tm.start();
  if (stages > 1) {
    CB_TRY(gather_essentials(ctx, comm->row, stages, A, essA));
    CB_TRY(gather_essentials(ctx, comm->col, stages, B, essB));
  } else {
    essA = {A->nnz, A->nzc, A->m, A->n};
    essB = {B->nnz, B->nzc, B->m, B->n};
  }
  ds->ms_bcast += tm.stop();
  int rc = CBGPU_OK;
  if (stages > 1 && ctx->opt.summa_fused) {
    // Fused SUMMA: receive the blocks of every stage, then compute sum_i A_i (x) B_i as ONE local multiply
    // [A_0 A_1 ...] (x) [B_0; B_1; ...]. Same products and the same result as the stage loop + MultiwayMerge
    // (ParFriends.h:1482-1548), but the partial results are never materialised and never merged.
    std::vector<cbgpu_mat *> Ablk(stages, nullptr), Bblk(stages, nullptr), Arecv(stages, nullptr), Brecv(stages, nullptr);
    tm.start();
    for (int i = 0; i < stages && rc == CBGPU_OK; ++i) {
      rc = bcast_block(ctx, comm->row, i, g.my_col, A, &essA[4 * i], ta, &Arecv[i], &ds->bytes_bcast);
      if (rc == CBGPU_OK) rc = bcast_block(ctx, comm->col, i, g.my_row, B, &essB[4 * i], tb, &Brecv[i], &ds->bytes_bcast);
      Ablk[i] = Arecv[i] ? Arecv[i] : const_cast<cbgpu_mat *>(A);
      Bblk[i] = Brecv[i] ? Brecv[i] : const_cast<cbgpu_mat *>(B);
      if (rc == CBGPU_OK && Ablk[i]->n != Bblk[i]->m)
        rc = set_error(ctx, CBGPU_ERR_DIMMISMATCH, "stage %d: inner block dimensions differ (%lld vs %lld)", i, (long long)Ablk[i]->n, (long long)Bblk[i]->m);
      ds->stages++;
    }
    ds->ms_bcast += tm.stop();
    cbgpu_mat *Acat = nullptr, *Bcat = nullptr;
    if (rc == CBGPU_OK) {
      tm.start();
      rc = mat_colconcat(ctx, stages, Ablk.data(), &Acat);
      if (rc == CBGPU_OK) rc = mat_rowstack(ctx, stages, Bblk.data(), &Bcat);
      for (int i = 0; i < stages; ++i) {
        mat_release(ctx, Arecv[i]);
        mat_release(ctx, Brecv[i]);
        Arecv[i] = Brecv[i] = nullptr;
      }
      ds->ms_merge += tm.stop(); // time of assembling the stacked operands
    }
    if (rc == CBGPU_OK) {
      cbgpu_stats st;
      memset(&st, 0, sizeof(st));
      tm.start();
      rc = cbgpu_spgemm_local(ctx, semiring, Acat, Bcat, C, &st);
      ds->ms_multiply += tm.stop();
      if (rc == CBGPU_OK) add_stats(ds->local, st);
    }
    mat_release(ctx, Acat);
    mat_release(ctx, Bcat);
    for (int i = 0; i < stages; ++i) {
      mat_release(ctx, Arecv[i]);
      mat_release(ctx, Brecv[i]);
    }
    return rc;
  }
  std::vector<cbgpu_mat *> partial;
  for (int i = 0; i < stages && rc == CBGPU_OK; ++i) {
    cbgpu_mat *Ar = nullptr, *Br = nullptr;
    const cbgpu_mat *Ause = A, *Buse = B;
    tm.start();
    if (stages > 1) {
      rc = bcast_block(ctx, comm->row, i, g.my_col, A, &essA[4 * i], ta, &Ar, &ds->bytes_bcast);
      if (rc == CBGPU_OK) rc = bcast_block(ctx, comm->col, i, g.my_row, B, &essB[4 * i], tb, &Br, &ds->bytes_bcast);
      if (Ar) Ause = Ar;
      if (Br) Buse = Br;
    }
    ds->ms_bcast += tm.stop();
    if (rc == CBGPU_OK) {
      if (Ause->n != Buse->m) rc = set_error(ctx, CBGPU_ERR_DIMMISMATCH, "stage %d: inner block dimensions differ (%lld vs %lld)", i, (long long)Ause->n, (long long)Buse->m);
    }
    if (rc == CBGPU_OK && Ause->nnz > 0 && Buse->nnz > 0) {
      cbgpu_mat *Ci = nullptr;
      cbgpu_stats st;
      memset(&st, 0, sizeof(st));
      tm.start();
      rc = cbgpu_spgemm_local(ctx, semiring, Ause, Buse, &Ci, &st);
      ds->ms_multiply += tm.stop();
      if (rc == CBGPU_OK) {
        add_stats(ds->local, st);
        if (Ci->nnz > 0) partial.push_back(Ci); // empty stage results are not merged (ParFriends.h:1524)
        else mat_release(ctx, Ci);
      }
    }
    mat_release(ctx, Ar);
    mat_release(ctx, Br);
    ds->stages++;
  }
  if (rc == CBGPU_OK) {
    tm.start();
    if (partial.empty()) {
      cbgpu_mat *E = nullptr;
      rc = mat_alloc(ctx, A->m, B->n, 0, 0, tc, &E);
      if (rc == CBGPU_OK) {
        cudaMemsetAsync(E->cp, 0, 8, ctx->stream);
        *C = E;
      }
    } else if (partial.size() == 1) {
      *C = partial[0]; // MultiwayMerge steals a single list (MultiwayMerge.h:437-442)
      partial.clear();
    } else {
      cbgpu_stats st;
      memset(&st, 0, sizeof(st));
      rc = cbgpu_merge(ctx, semiring, (int)partial.size(), partial.data(), C, &st);
      ds->local.kernel_launches += st.kernel_launches;
    }
    ds->ms_merge += tm.stop();
  }
  for (cbgpu_mat *p : partial) mat_release(ctx, p);
  return rc;
}

// The fiber stage of the 3D algorithm (ParFriends.h:3578-3642): cut this layer's partial product Cl (consumed) into L
// column slabs, ship slab l to fiber rank l, merge what arrives with the own slab. Runs entirely on ctx's stream.
static int fiber_reduce(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, cbgpu_mat *Cl, cbgpu_mat **C, cbgpu_dist_stats *ds,
                        const int64_t *widths = nullptr) {
  const cbgpu_grid &g = comm->grid;
  const int L = g.layers;
  Timer tm(ctx->stream);
  int rc = CBGPU_OK;
  // ---- cut the layer result into L column slabs (CalculateColSplitDistributionOfLayer, SpParMat3D.cpp:576-609;
  //      send ranges ParFriends->h:3578-3600); slab l belongs to fiber rank l
  tm.start();
  std::vector<cbgpu_mat *> slab(L, nullptr), recv(L, nullptr);
  int64_t cut = 0;
  for (int l = 0; l < L && rc == CBGPU_OK; ++l) {
    int64_t c0, c1;
    if (widths) {
      c0 = cut;
      c1 = cut + widths[l];
      cut = c1;
    } else {
      cbgpu_block_range(Cl->n, L, l, &c0, &c1);
    }
    rc = mat_colslice(ctx, Cl, c0, c1, &slab[l]);
  }
  mat_release(ctx, Cl);
  // ---- sizes: every rank tells every fiber peer {nnz, nzc} of the slab it will send (ParFriends->h:3602)
  std::vector<int64_t> sizes((size_t)2 * L * L, 0);
  if (rc == CBGPU_OK) {
    std::vector<int64_t> mine((size_t)2 * L);
    for (int l = 0; l < L; ++l) { mine[2 * l] = slab[l]->nnz; mine[2 * l + 1] = slab[l]->nzc; }
    int64_t *d = nullptr;
    rc = dev_alloc_t(ctx, &d, (size_t)2 * L * (L + 1));
    if (rc == CBGPU_OK) {
      cudaMemcpyAsync(d, mine.data(), sizeof(int64_t) * 2 * L, cudaMemcpyHostToDevice, ctx->stream);
      ncclResult_t r = nccl().AllGather(d, d + 2 * L, (size_t)2 * L, ncclInt64, comm->fiber, ctx->stream);
      if (r != ncclSuccess) rc = set_error(ctx, CBGPU_ERR_NCCL, "fiber allgather failed: %s", nccl().GetErrorString(r));
      cudaMemcpyAsync(sizes.data(), d + 2 * L, sizeof(int64_t) * 2 * L * L, cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      dev_free(ctx, d);
    }
  }
  // ---- exchange (ParFriends->h:3612): grouped send/recv of the four arrays per peer
  const int me = g.my_layer;
  int tc = slab[0] ? slab[0]->dtype : CBGPU_F64;
  const size_t vb = dtype_size(tc);
  if (rc == CBGPU_OK) {
    for (int p = 0; p < L && rc == CBGPU_OK; ++p) {
      if (p == me) continue;
      int64_t nnz = sizes[(size_t)2 * L * p + 2 * me], nzc = sizes[(size_t)2 * L * p + 2 * me + 1];
      rc = mat_alloc(ctx, slab[me]->m, slab[me]->n, nnz, nzc, tc, &recv[p]);
      if (rc == CBGPU_OK && nzc == 0) cudaMemsetAsync(recv[p]->cp, 0, 8, ctx->stream);
    }
  }
  if (rc == CBGPU_OK) {
    ncclResult_t r = nccl().GroupStart();
    for (int p = 0; p < L && r == ncclSuccess; ++p) {
      if (p == me) continue;
      cbgpu_mat *S = slab[p], *R = recv[p];
      if (S->nnz > 0) {
        r = nccl().Send(S->jc, (size_t)S->nzc * 8, nccl_bytes(), p, comm->fiber, ctx->stream);
        if (r == ncclSuccess) r = nccl().Send(S->cp, (size_t)(S->nzc + 1) * 8, nccl_bytes(), p, comm->fiber, ctx->stream);
        if (r == ncclSuccess) r = nccl().Send(S->ir, (size_t)S->nnz * 4, nccl_bytes(), p, comm->fiber, ctx->stream);
        if (r == ncclSuccess) r = nccl().Send(S->numx, (size_t)S->nnz * vb, nccl_bytes(), p, comm->fiber, ctx->stream);
        ds->bytes_fiber += S->nzc * 16 + 8 + S->nnz * (4 + (int64_t)vb);
      }
      if (r == ncclSuccess && R->nnz > 0) {
        r = nccl().Recv(R->jc, (size_t)R->nzc * 8, nccl_bytes(), p, comm->fiber, ctx->stream);
        if (r == ncclSuccess) r = nccl().Recv(R->cp, (size_t)(R->nzc + 1) * 8, nccl_bytes(), p, comm->fiber, ctx->stream);
        if (r == ncclSuccess) r = nccl().Recv(R->ir, (size_t)R->nnz * 4, nccl_bytes(), p, comm->fiber, ctx->stream);
        if (r == ncclSuccess) r = nccl().Recv(R->numx, (size_t)R->nnz * vb, nccl_bytes(), p, comm->fiber, ctx->stream);
      }
    }
    ncclResult_t r2 = nccl().GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess)
      rc = set_error(ctx, CBGPU_ERR_NCCL, "fiber exchange failed: %s", nccl().GetErrorString(r != ncclSuccess ? r : r2));
  }
  ds->ms_fiber_exchange += tm.stop();
  // ---- merge what arrived with my own slab (ParFriends->h:3642)
  if (rc == CBGPU_OK) {
    tm.start();
    std::vector<cbgpu_mat *> lists;
    for (int p = 0; p < L; ++p) {
      cbgpu_mat *M = (p == me) ? slab[me] : recv[p];
      if (M && M->nnz > 0) lists.push_back(M);
    }
    if (lists.empty()) {
      *C = slab[me];
      slab[me] = nullptr;
    } else if (lists.size() == 1) {
      *C = lists[0];
      for (int p = 0; p < L; ++p) {
        if (slab[p] == lists[0]) slab[p] = nullptr;
        if (recv[p] == lists[0]) recv[p] = nullptr;
      }
    } else {
      cbgpu_stats st;
      memset(&st, 0, sizeof(st));
      rc = cbgpu_merge(ctx, semiring, (int)lists.size(), lists.data(), C, &st);
      ds->local.kernel_launches += st.kernel_launches;
    }
    ds->ms_fiber_merge += tm.stop();
  }
  for (int p = 0; p < L; ++p) {
    mat_release(ctx, slab[p]);
    mat_release(ctx, recv[p]);
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------- fiber-fused 3D multiply
// Option `fiber_fused` (off until it has been validated on several GPUs): the 3D product without partial results.
// The reference's 3D algorithm lets every layer l compute the full-size partial product A(:,K_l) B(K_l,:) and then
// reduce-scatters C along the fiber (ParFriends.h:3578-3642): the traffic and the merge are O(nnz(C)) per layer. For
// SpGEMM nnz(C) >> nnz(A) + nnz(B), so it is far cheaper to replicate the INPUTS along the fiber instead: rank (i,j,l')
// collects the blocks A_l(i,k) of its process row from every stage k of every layer l (row broadcasts + fiber
// all-gather) and the column sub-slab l' of B_l(k,j) from every (l,k) (column broadcasts + fiber all-to-all), and
// computes its final piece C(i,j)[:, sub-slab l'] = [A_0(i,:) A_1(i,:) ...] (x) [B_0(:,j); B_1(:,j); ...][:, sub-slab l'] as ONE
// local multiply. Same operand distribution (A column-split, B row-split across layers), same products, same output
// distribution (C column-split across layers) as Mult_AnXBn_SUMMA3D; nothing is merged.

// All-to-all of DCSC blocks along communicator c (nranks ranks, this rank is `me`): slab[p] goes to rank p, recv[p] arrives
// from rank p (p != me). Essentials {nnz, nzc, m, n} by one all-gather, then grouped send/recv of the four arrays
// (the device-to-device counterpart of the tuple Alltoallv of ParFriends.h:3612 and SpParMat3D.cpp:51-97).
static int exchange_blocks(cbgpu_ctx *ctx, ncclComm_t c, int nranks, int me, std::vector<cbgpu_mat *> &slab,
                           std::vector<cbgpu_mat *> &recv, int64_t *bytes) {
  const int L = nranks;
  const int dt = slab[me]->dtype;
  const size_t vb = dtype_size(dt);
  std::vector<int64_t> mine((size_t)4 * L), sizes((size_t)4 * L * L, 0);
  for (int l = 0; l < L; ++l) {
    mine[4 * l] = slab[l]->nnz;
    mine[4 * l + 1] = slab[l]->nzc;
    mine[4 * l + 2] = slab[l]->m;
    mine[4 * l + 3] = slab[l]->n;
  }
  int64_t *d = nullptr;
  CB_TRY(dev_alloc_t(ctx, &d, (size_t)4 * L * (L + 1)));
  CB_CUDA(ctx, cudaMemcpyAsync(d, mine.data(), sizeof(int64_t) * 4 * L, cudaMemcpyHostToDevice, ctx->stream));
  CB_NCCL(ctx, nccl().AllGather(d, d + 4 * L, (size_t)4 * L, ncclInt64, c, ctx->stream));
  CB_CUDA(ctx, cudaMemcpyAsync(sizes.data(), d + 4 * L, sizeof(int64_t) * 4 * L * L, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CB_TRY(dev_free(ctx, d));
  for (int p = 0; p < L; ++p) {
    if (p == me) continue;
    const int64_t *e = &sizes[(size_t)4 * L * p + 4 * me];
    CB_TRY(mat_alloc(ctx, e[2], e[3], e[0], e[1], dt, &recv[p]));
    if (e[0] == 0) CB_CUDA(ctx, cudaMemsetAsync(recv[p]->cp, 0, 8, ctx->stream));
  }
  ncclResult_t r = nccl().GroupStart();
  for (int p = 0; p < L && r == ncclSuccess; ++p) {
    if (p == me) continue;
    cbgpu_mat *S = slab[p], *R = recv[p];
    if (S->nnz > 0) {
      r = nccl().Send(S->jc, (size_t)S->nzc * 8, nccl_bytes(), p, c, ctx->stream);
      if (r == ncclSuccess) r = nccl().Send(S->cp, (size_t)(S->nzc + 1) * 8, nccl_bytes(), p, c, ctx->stream);
      if (r == ncclSuccess) r = nccl().Send(S->ir, (size_t)S->nnz * 4, nccl_bytes(), p, c, ctx->stream);
      if (r == ncclSuccess) r = nccl().Send(S->numx, (size_t)S->nnz * vb, nccl_bytes(), p, c, ctx->stream);
      *bytes += S->nzc * 16 + 8 + S->nnz * (4 + (int64_t)vb);
    }
    if (r == ncclSuccess && R->nnz > 0) {
      r = nccl().Recv(R->jc, (size_t)R->nzc * 8, nccl_bytes(), p, c, ctx->stream);
      if (r == ncclSuccess) r = nccl().Recv(R->cp, (size_t)(R->nzc + 1) * 8, nccl_bytes(), p, c, ctx->stream);
      if (r == ncclSuccess) r = nccl().Recv(R->ir, (size_t)R->nnz * 4, nccl_bytes(), p, c, ctx->stream);
      if (r == ncclSuccess) r = nccl().Recv(R->numx, (size_t)R->nnz * vb, nccl_bytes(), p, c, ctx->stream);
    }
  }
  ncclResult_t r2 = nccl().GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess)
    return set_error(ctx, CBGPU_ERR_NCCL, "block exchange failed: %s", nccl().GetErrorString(r != ncclSuccess ? r : r2));
  return CBGPU_OK;
}
static int fiber_exchange_slabs(cbgpu_ctx *ctx, cbgpu_comm *comm, std::vector<cbgpu_mat *> &slab, std::vector<cbgpu_mat *> &recv,
                                int64_t *bytes) {
  return exchange_blocks(ctx, comm->fiber, comm->grid.layers, comm->grid.my_layer, slab, recv, bytes);
}
// all-to-all over the whole grid (csrc/reshape.cu: 2D <-> 3D redistribution)
int exchange_blocks_world(cbgpu_ctx *ctx, cbgpu_comm *comm, std::vector<cbgpu_mat *> &send, std::vector<cbgpu_mat *> &recv,
                          int64_t *bytes) {
  if (!nccl().ok) return set_error(ctx, CBGPU_ERR_NCCL, "libnccl.so.2 could not be loaded");
  return exchange_blocks(ctx, comm->world, comm->grid.world, comm->grid.rank, send, recv, bytes);
}

// all blocks of `own` along communicator `c` (nranks ranks, this rank is `me`), in rank order; blk[me] == own
static int allgather_blocks(cbgpu_ctx *ctx, ncclComm_t c, int nranks, int me, const cbgpu_mat *own, std::vector<cbgpu_mat *> &blk,
                            std::vector<cbgpu_mat *> &owned, int64_t *bytes) {
  std::vector<int64_t> ess;
  CB_TRY(gather_essentials(ctx, c, nranks, own, ess));
  blk.assign(nranks, nullptr);
  owned.assign(nranks, nullptr);
  for (int r = 0; r < nranks; ++r) {
    CB_TRY(bcast_block(ctx, c, r, me, own, &ess[4 * r], own->dtype, &owned[r], bytes));
    blk[r] = owned[r] ? owned[r] : const_cast<cbgpu_mat *>(own);
  }
  return CBGPU_OK;
}

static void release_all(cbgpu_ctx *ctx, std::vector<cbgpu_mat *> &v) {
  for (cbgpu_mat *&m : v) {
    mat_release(ctx, m);
    m = nullptr;
  }
}

// [A_0(i,:) A_1(i,:) ...]: the A blocks of this process row from every stage of every layer, (layer, stage) order
static int gather_A_all_layers(cbgpu_ctx *ctx, cbgpu_comm *comm, const cbgpu_mat *A, cbgpu_mat **Aall, cbgpu_dist_stats *ds) {
  const cbgpu_grid &g = comm->grid;
  Timer tm(ctx->stream);
  tm.start();
  int rc = CBGPU_OK;
  cbgpu_mat *Arow = nullptr; // [A_l(i,0) ... A_l(i,stages-1)] of my layer
  const cbgpu_mat *mine = A;
  if (g.grid_cols > 1) {
    std::vector<cbgpu_mat *> blk, owned;
    rc = allgather_blocks(ctx, comm->row, g.grid_cols, g.my_col, A, blk, owned, &ds->bytes_bcast);
    if (rc == CBGPU_OK) rc = mat_colconcat(ctx, g.grid_cols, blk.data(), &Arow);
    release_all(ctx, owned);
    mine = Arow;
  }
  if (rc == CBGPU_OK) {
    std::vector<cbgpu_mat *> blk, owned;
    rc = allgather_blocks(ctx, comm->fiber, g.layers, g.my_layer, mine, blk, owned, &ds->bytes_fiber);
    if (rc == CBGPU_OK) rc = mat_colconcat(ctx, g.layers, blk.data(), Aall);
    release_all(ctx, owned);
  }
  mat_release(ctx, Arow);
  ds->ms_bcast += tm.stop();
  return rc;
}

// one column slab Bslab of this rank's B block -> this rank's final piece of C for that slab
static int fiber_fused_slab(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *Aall, const cbgpu_mat *Bslab,
                            cbgpu_mat **C, cbgpu_dist_stats *ds, int64_t *sym_flops = nullptr, int64_t *sym_nnz = nullptr,
                            const int64_t *widths = nullptr /* columns of the slab that belong to fiber rank l; null: even cut */) {
  const cbgpu_grid &g = comm->grid;
  const int L = g.layers;
  Timer tm(ctx->stream);
  int rc = CBGPU_OK;
  // [B_l(0,j); B_l(1,j); ...] of my layer (column broadcasts of the SUMMA stages)
  tm.start();
  cbgpu_mat *Bcol = nullptr;
  const cbgpu_mat *mine = Bslab;
  if (g.grid_rows > 1) {
    std::vector<cbgpu_mat *> blk, owned;
    rc = allgather_blocks(ctx, comm->col, g.grid_rows, g.my_row, Bslab, blk, owned, &ds->bytes_bcast);
    if (rc == CBGPU_OK) rc = mat_rowstack(ctx, g.grid_rows, blk.data(), &Bcol);
    release_all(ctx, owned);
    mine = Bcol;
    ds->stages += g.grid_rows;
  }
  ds->ms_bcast += tm.stop();
  // column sub-slab l' of it goes to fiber rank l' (the split fiber_reduce applies to C: SpParMat3D.cpp:576-609)
  std::vector<cbgpu_mat *> slab(L, nullptr), recv(L, nullptr), parts(L, nullptr);
  tm.start();
  int64_t cut = 0;
  for (int l = 0; l < L && rc == CBGPU_OK; ++l) {
    int64_t c0, c1;
    if (widths) {
      c0 = cut;
      c1 = cut + widths[l];
      cut = c1;
    } else {
      cbgpu_block_range(mine->n, L, l, &c0, &c1);
    }
    rc = mat_colslice(ctx, mine, c0, c1, &slab[l]);
  }
  if (rc == CBGPU_OK) rc = fiber_exchange_slabs(ctx, comm, slab, recv, &ds->bytes_fiber);
  cbgpu_mat *Ball = nullptr;
  if (rc == CBGPU_OK) {
    for (int l = 0; l < L; ++l) parts[l] = (l == g.my_layer) ? slab[l] : recv[l];
    rc = mat_rowstack(ctx, L, parts.data(), &Ball);
  }
  ds->ms_fiber_exchange += tm.stop();
  release_all(ctx, slab);
  release_all(ctx, recv);
  mat_release(ctx, Bcol);
  if (rc == CBGPU_OK && Aall->n != Ball->m)
    rc = set_error(ctx, CBGPU_ERR_DIMMISMATCH, "fiber-fused multiply: inner dimensions differ (%lld vs %lld)", (long long)Aall->n,
                   (long long)Ball->m);
  if (rc == CBGPU_OK && sym_flops) { // symbolic only: the products and the outputs this rank will produce
    tm.start();
    rc = cbgpu_spgemm_symbolic(ctx, Aall, Ball, sym_flops, sym_nnz);
    ds->ms_multiply += tm.stop();
  } else if (rc == CBGPU_OK) {
    cbgpu_stats st;
    memset(&st, 0, sizeof(st));
    tm.start();
    rc = cbgpu_spgemm_local(ctx, semiring, Aall, Ball, C, &st);
    ds->ms_multiply += tm.stop();
    if (rc == CBGPU_OK) add_stats(ds->local, st);
  }
  mat_release(ctx, Ball);
  return rc;
}

} // namespace cbgpu

extern "C" {

int cbgpu_summa2d(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, cbgpu_mat **C,
                  cbgpu_dist_stats *stats) {
  if (!ctx || !comm || !A || !B || !C) return CBGPU_ERR_INVALID;
  if (comm->grid.layers != 1) return set_error(ctx, CBGPU_ERR_GRID, "cbgpu_summa2d needs a single-layer grid");
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_dist_stats ds;
  memset(&ds, 0, sizeof(ds));
  Timer all(ctx->stream);
  all.start();
  int rc = summa_layer(ctx, comm, semiring, A, B, C, &ds);
  ds.ms_total = all.stop();
  if (stats) *stats = ds;
  return rc;
}

/* Distributed symbolic pass: the exact number of products and of outputs THIS rank produces in the distributed product, in
 * the final distribution of C (column-split across layers). replaces: EstPerProcessNnzSUMMA (ParFriends.h:1698, an estimate
 * that returns 0 in-tree) and the per-stage estimateFLOP / estimateNNZ loop of CalculateNumberOfPhases (:780-843). The inputs
 * are gathered exactly as the fiber-fused multiply gathers them; nothing of C is allocated. */
int cbgpu_summa_symbolic(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int64_t *flops,
                         int64_t *nnz_out) {
  if (!ctx || !comm || !A || !B || !flops || !nnz_out) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_dist_stats ds;
  memset(&ds, 0, sizeof(ds));
  cbgpu_mat *Aall = nullptr;
  int rc = gather_A_all_layers(ctx, comm, A, &Aall, &ds);
  if (rc == CBGPU_OK) rc = fiber_fused_slab(ctx, comm, semiring, Aall, B, nullptr, &ds, flops, nnz_out);
  mat_release(ctx, Aall);
  return rc;
}

int cbgpu_summa3d(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, cbgpu_mat **C,
                  cbgpu_dist_stats *stats) {
  if (!ctx || !comm || !A || !B || !C) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  cbgpu_dist_stats ds;
  memset(&ds, 0, sizeof(ds));
  Timer all(ctx->stream);
  all.start();
  int rc = CBGPU_OK;
  if (comm->grid.layers > 1 && ctx->opt.fiber_fused) {
    cbgpu_mat *Aall = nullptr;
    rc = gather_A_all_layers(ctx, comm, A, &Aall, &ds);
    if (rc == CBGPU_OK) rc = fiber_fused_slab(ctx, comm, semiring, Aall, B, C, &ds);
    mat_release(ctx, Aall);
  } else {
    cbgpu_mat *Cl = nullptr; // this layer's partial product: full block shape
    CB_TRY(summa_layer(ctx, comm, semiring, A, B, &Cl, &ds));
    if (comm->grid.layers == 1) *C = Cl;
    else rc = fiber_reduce(ctx, comm, semiring, Cl, C, &ds);
  }
  ds.ms_total = all.stop();
  if (stats) *stats = ds;
  return rc;
}

// Phased distributed multiply (MemEfficientSpGEMM ParFriends.h:453-777, MemEfficientSpGEMM3D :3674-4170): one SUMMA per
// column slab of B, so that C never has to exist as a whole.
//   one layer : slab p = ColSplit part p of B's local columns (dcsc.cpp:1202).
//   L layers  : the reference's plan (ParFriends.h:3774-3811): B's local columns are first cut into the L chunks the fiber
//               owns in C (CalculateColSplitDistributionOfLayer, SpParMat3D.cpp:576-609), every chunk into `phases` pieces,
//               and phase p takes piece p of EVERY chunk -- each fiber rank gets work in every phase, and the pieces a rank
//               produces over the phases concatenate to exactly its chunk: the layout of the unphased Mult_AnXBn_SUMMA3D.
// With several layers the default is the fiber-fused formulation (inputs replicated along the fiber, nothing merged). With
// fiber_fused = 0 the reference's fiber reduction runs after each slab; option fiber_pipeline = 1 overlaps the reduction of
// slab p with the multiply of slab p+1 on a second host thread, context and stream (two NCCL communicators are then in
// use concurrently from two threads: opt-in).
struct PhasePlan {
  int phases = 1, L = 1;
  std::vector<int64_t> c0, c1; // [p * L + l]: columns of B's block in piece p of the chunk of fiber rank l
  int64_t begin(int p, int l) const { return c0[(size_t)p * L + l]; }
  int64_t width(int p, int l) const { return c1[(size_t)p * L + l] - c0[(size_t)p * L + l]; }
};
static PhasePlan make_phase_plan(int64_t n, int phases, int L) {
  PhasePlan pl;
  pl.phases = phases;
  pl.L = L;
  pl.c0.assign((size_t)phases * L, 0);
  pl.c1.assign((size_t)phases * L, 0);
  for (int l = 0; l < L; ++l) {
    int64_t k0, k1;
    cbgpu_block_range(n, L, l, &k0, &k1);
    for (int p = 0; p < phases; ++p) {
      int64_t a, b;
      cbgpu_block_range(k1 - k0, phases, p, &a, &b);
      pl.c0[(size_t)p * L + l] = k0 + a;
      pl.c1[(size_t)p * L + l] = k0 + b;
    }
  }
  return pl;
}
// the columns of B that phase p multiplies: piece p of every chunk, side by side. owned: the caller releases it
static int cut_phase_slab(cbgpu_ctx *ctx, const cbgpu_mat *B, const PhasePlan &pl, int p, cbgpu_mat **out, bool *owned) {
  if (pl.phases == 1 && pl.L == 1) {
    *out = const_cast<cbgpu_mat *>(B);
    *owned = false;
    return CBGPU_OK;
  }
  *owned = true;
  if (pl.L == 1) return mat_colslice(ctx, B, pl.begin(p, 0), pl.begin(p, 0) + pl.width(p, 0), out);
  std::vector<cbgpu_mat *> parts(pl.L, nullptr);
  int rc = CBGPU_OK;
  for (int l = 0; l < pl.L && rc == CBGPU_OK; ++l) rc = mat_colslice(ctx, B, pl.begin(p, l), pl.begin(p, l) + pl.width(p, l), &parts[l]);
  if (rc == CBGPU_OK) rc = mat_colconcat(ctx, pl.L, parts.data(), out);
  for (cbgpu_mat *m : parts) mat_release(ctx, m);
  return rc;
}

} // extern "C" (plan helpers above have C++ linkage on purpose)
extern "C" int cbgpu_phase_columns(int64_t n, int phases, int layers, int phase, int layer, int64_t *begin, int64_t *end) {
  if (n < 0 || phases < 1 || layers < 1 || phase < 0 || phase >= phases || layer < 0 || layer >= layers || !begin || !end)
    return CBGPU_ERR_INVALID;
  const PhasePlan pl = make_phase_plan(n, phases, layers);
  *begin = pl.begin(phase, layer);
  *end = *begin + pl.width(phase, layer);
  return CBGPU_OK;
}
extern "C" {

// what happens to a finished piece of C before it is reported / kept (the pruning of MemEfficientSpGEMM, ParFriends.h:744)
struct SlabEpilogue {
  virtual int apply(cbgpu_ctx *c, int p, cbgpu_mat **Cp) = 0;
  virtual ~SlabEpilogue() {}
};

static int summa_phased_impl(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                             int want_checksum, bool global, int64_t row_offset, int64_t col_offset, cbgpu_mat **slabs,
                             cbgpu_slab_result *results, cbgpu_dist_stats *stats, SlabEpilogue *epilogue = nullptr);

int cbgpu_summa_phased(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                       int want_checksum, cbgpu_mat **slabs, cbgpu_slab_result *results, cbgpu_dist_stats *stats) {
  return summa_phased_impl(ctx, comm, semiring, A, B, phases, want_checksum, false, 0, 0, slabs, results, stats);
}

int cbgpu_summa_phased_global(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                              int64_t row_offset, int64_t col_offset, cbgpu_mat **slabs, cbgpu_slab_result *results,
                              cbgpu_dist_stats *stats) {
  return summa_phased_impl(ctx, comm, semiring, A, B, phases, 1, true, row_offset, col_offset, slabs, results, stats);
}

static int summa_phased_impl(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                             int want_checksum, bool global, int64_t row_offset, int64_t col_offset, cbgpu_mat **slabs,
                             cbgpu_slab_result *results, cbgpu_dist_stats *stats, SlabEpilogue *epilogue) {
  if (!ctx || !comm || !A || !B || phases < 1 || !results) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int L = comm->grid.layers, me = comm->grid.my_layer;
  const PhasePlan plan = make_phase_plan(B->n, phases, L);
  // the piece of C this rank ends up with in phase p covers the columns [plan.begin(p, me), + plan.width(p, me)) of B's block
  auto slab_checksum = [&](cbgpu_ctx *c, int p, const cbgpu_mat *Cp, cbgpu_slab_result *r) -> int {
    if (!global) return cbgpu_mat_checksum(c, Cp, &r->pattern_sum, &r->value_sum);
    return cbgpu_mat_checksum_at(c, Cp, row_offset, col_offset + plan.begin(p, me), &r->pattern_sum, &r->value_sum);
  };
  auto finish_slab = [&](cbgpu_ctx *c, int p, cbgpu_mat *Cp) -> int {
    int rc = CBGPU_OK;
    if (epilogue) rc = epilogue->apply(c, p, &Cp);
    if (rc != CBGPU_OK) {
      mat_release(c, Cp);
      return rc;
    }
    results[p].nnz = Cp->nnz;
    results[p].nzc = Cp->nzc;
    results[p].pattern_sum = results[p].value_sum = 0;
    if (want_checksum) rc = slab_checksum(c, p, Cp, &results[p]);
    if (slabs) slabs[p] = Cp;
    else mat_release(c, Cp);
    return rc;
  };
  cbgpu_dist_stats ds;
  memset(&ds, 0, sizeof(ds));
  Timer all(ctx->stream);
  all.start();
  std::vector<int64_t> widths(L, 0);
  const bool fiber_fused = L > 1 && ctx->opt.fiber_fused;
  if (fiber_fused || L == 1) {
    // L > 1: inputs replicated along the fiber instead of partial results reduced along it: A's side is gathered once for
    // all phases, every phase exchanges its (small) slab of B and multiplies; one host thread, nothing to merge
    cbgpu_mat *Aall = nullptr;
    int rc = CBGPU_OK;
    if (L > 1) rc = gather_A_all_layers(ctx, comm, A, &Aall, &ds);
    for (int p = 0; p < phases && rc == CBGPU_OK; ++p) {
      cbgpu_mat *Bs = nullptr, *Cp = nullptr;
      bool owned = false;
      rc = cut_phase_slab(ctx, B, plan, p, &Bs, &owned);
      if (rc == CBGPU_OK) {
        if (L > 1) {
          for (int l = 0; l < L; ++l) widths[l] = plan.width(p, l);
          rc = fiber_fused_slab(ctx, comm, semiring, Aall, Bs, &Cp, &ds, nullptr, nullptr, widths.data());
        } else {
          rc = summa_layer(ctx, comm, semiring, A, Bs, &Cp, &ds);
        }
      }
      if (owned) mat_release(ctx, Bs);
      if (rc == CBGPU_OK) rc = finish_slab(ctx, p, Cp);
    }
    mat_release(ctx, Aall);
    ds.ms_total = all.stop();
    if (stats) *stats = ds;
    return rc;
  }
  // ---- the reference's formulation: per-layer SUMMA of the slab, then the fiber reduction of the partial result
  const bool pipelined = ctx->opt.fiber_pipeline != 0;
  if (pipelined && !comm->ctx2) {
    int rc2 = cbgpu_create(ctx->device, nullptr, &comm->ctx2);
    if (rc2 != CBGPU_OK) return set_error(ctx, rc2, "could not create the second context of the pipeline");
    comm->ctx2->opt = ctx->opt;
  }
  struct Item { cbgpu_mat *Cl; cudaEvent_t ready; };
  std::mutex mu;
  std::condition_variable cv;
  std::vector<Item> queue((size_t)phases, Item{nullptr, nullptr});
  int produced = 0, consumed = 0;
  bool producer_failed = false;
  int worker_rc = CBGPU_OK;
  cbgpu_dist_stats wds;
  memset(&wds, 0, sizeof(wds));
  std::thread worker;
  if (pipelined) {
    worker = std::thread([&]() {
      cbgpu_ctx *c2 = comm->ctx2;
      cudaSetDevice(c2->device);
      std::vector<int64_t> w2(L, 0);
      for (int p = 0; p < phases; ++p) {
        Item it;
        {
          std::unique_lock<std::mutex> lock(mu);
          cv.wait(lock, [&] { return produced > p || producer_failed; });
          if (produced <= p) return; // the producer gave up
          it = queue[p];
        }
        cudaStreamWaitEvent(c2->stream, it.ready, 0);
        cbgpu_mat *Cp = nullptr;
        for (int l = 0; l < L; ++l) w2[l] = plan.width(p, l);
        int rc = fiber_reduce(c2, comm, semiring, it.Cl, &Cp, &wds, w2.data());
        if (rc == CBGPU_OK) rc = finish_slab(c2, p, Cp);
        cudaStreamSynchronize(c2->stream);
        cudaEventDestroy(it.ready);
        {
          std::lock_guard<std::mutex> lock(mu);
          consumed = p + 1;
          if (rc != CBGPU_OK && worker_rc == CBGPU_OK) worker_rc = rc;
        }
        cv.notify_all();
        if (rc != CBGPU_OK) return;
      }
    });
  }
  int rc = CBGPU_OK;
  for (int p = 0; p < phases && rc == CBGPU_OK; ++p) {
    cbgpu_mat *Bs = nullptr, *Cl = nullptr;
    bool owned = false;
    rc = cut_phase_slab(ctx, B, plan, p, &Bs, &owned);
    if (rc == CBGPU_OK) rc = summa_layer(ctx, comm, semiring, A, Bs, &Cl, &ds);
    if (owned) mat_release(ctx, Bs);
    if (rc != CBGPU_OK) break;
    if (!pipelined) { // one host thread, one stream: the fiber reduction of slab p runs before the multiply of slab p + 1
      cbgpu_mat *Cp = nullptr;
      for (int l = 0; l < L; ++l) widths[l] = plan.width(p, l);
      rc = fiber_reduce(ctx, comm, semiring, Cl, &Cp, &ds, widths.data());
      if (rc == CBGPU_OK) rc = finish_slab(ctx, p, Cp);
      continue;
    }
    Item it{Cl, nullptr};
    cudaEventCreateWithFlags(&it.ready, cudaEventDisableTiming);
    cudaEventRecord(it.ready, ctx->stream);
    {
      std::unique_lock<std::mutex> lock(mu);
      queue[p] = it;
      produced = p + 1;
      cv.notify_all();
      // keep at most one finished layer result waiting (memory) and stop early when the consumer failed
      cv.wait(lock, [&] { return consumed >= p || worker_rc != CBGPU_OK; });
      if (worker_rc != CBGPU_OK) rc = worker_rc;
    }
  }
  if (pipelined) {
    {
      std::lock_guard<std::mutex> lock(mu);
      if (rc != CBGPU_OK) producer_failed = true;
    }
    cv.notify_all();
    worker.join();
    if (rc == CBGPU_OK) rc = worker_rc;
    if (worker_rc != CBGPU_OK) ctx->last_error = comm->ctx2->last_error;
    ds.ms_fiber_exchange = wds.ms_fiber_exchange;
    ds.ms_fiber_merge = wds.ms_fiber_merge;
    ds.bytes_fiber = wds.bytes_fiber;
    ds.local.kernel_launches += wds.local.kernel_launches + comm->ctx2->launches;
    comm->ctx2->launches = 0;
  }
  ds.ms_total = all.stop();
  if (stats) *stats = ds;
  return rc;
}

// ---------------------------------------------------------------------------------------------- distributed MCL pruning
// MCLPruneRecoverySelect (ParFriends.h:186-354) decides column by column; on a grid a column of C is spread over the ranks
// of a process column (same grid column and layer), which the reference handles with column reductions and a distributed
// Kselect1 (SpParMat.cpp:1413-1700, reductions :1575-1660). Here the piece is re-cut along the process column instead: the
// local columns are dealt out evenly, every rank receives the row blocks of ITS columns from its peers (device-to-device
// all-to-all over NVLink), stacks them into whole columns, runs the single-GPU pruning kernels on them (csrc/prune.cu), and
// the survivors travel back to the row blocks they came from. Same per-column decisions as the reference (they depend on
// the whole column only); only the pruned entries make the return trip.
static int prune_distributed(cbgpu_ctx *ctx, cbgpu_comm *comm, cbgpu_mat *piece, double hardThreshold, int64_t selectNum,
                             int64_t recoverNum, double recoverPct, cbgpu_mat **out, cbgpu_prune_stats *pst, int64_t *bytes) {
  const cbgpu_grid &g = comm->grid;
  const int P = g.grid_rows, me = g.my_row;
  if (P == 1) return cbgpu_mcl_prune(ctx, piece, hardThreshold, selectNum, recoverNum, recoverPct, out, pst);
  int rc = CBGPU_OK;
  std::vector<cbgpu_mat *> send(P, nullptr), recv(P, nullptr), parts(P, nullptr);
  // forward: column range q of my piece goes to the rank in grid row q
  for (int q = 0; q < P && rc == CBGPU_OK; ++q) {
    int64_t c0, c1;
    cbgpu_block_range(piece->n, P, q, &c0, &c1);
    rc = mat_colslice(ctx, piece, c0, c1, &send[q]);
  }
  if (rc == CBGPU_OK) rc = exchange_blocks(ctx, comm->col, P, me, send, recv, bytes);
  cbgpu_mat *whole = nullptr, *pruned = nullptr;
  std::vector<int64_t> rows_of(P, 0);
  if (rc == CBGPU_OK) {
    for (int q = 0; q < P; ++q) {
      parts[q] = (q == me) ? send[q] : recv[q];
      rows_of[q] = parts[q]->m;
    }
    rc = mat_rowstack(ctx, P, parts.data(), &whole);
  }
  release_all(ctx, send);
  release_all(ctx, recv);
  if (rc == CBGPU_OK) rc = cbgpu_mcl_prune(ctx, whole, hardThreshold, selectNum, recoverNum, recoverPct, &pruned, pst);
  mat_release(ctx, whole);
  // backward: row block q of the pruned columns returns to the rank in grid row q
  int64_t r0 = 0;
  for (int q = 0; q < P && rc == CBGPU_OK; ++q) {
    rc = cbgpu_mat_submatrix(ctx, pruned, r0, r0 + rows_of[q], 0, pruned->n, &send[q]);
    r0 += rows_of[q];
  }
  mat_release(ctx, pruned);
  if (rc == CBGPU_OK) rc = exchange_blocks(ctx, comm->col, P, me, send, recv, bytes);
  if (rc == CBGPU_OK) {
    for (int q = 0; q < P; ++q) parts[q] = (q == me) ? send[q] : recv[q];
    rc = mat_colconcat(ctx, P, parts.data(), out);
  }
  release_all(ctx, send);
  release_all(ctx, recv);
  return rc;
}

struct PruneEpilogue : SlabEpilogue {
  cbgpu_comm *comm;
  double hard, pct;
  int64_t sel, rec;
  cbgpu_memeff_stats *ms;
  int64_t bytes = 0;
  int apply(cbgpu_ctx *c, int, cbgpu_mat **Cp) override {
    ms->nnz_unpruned += (*Cp)->nnz;
    if ((*Cp)->dtype != CBGPU_F64 && (*Cp)->dtype != CBGPU_F32) return CBGPU_OK; // the reference prunes floating point only
    cbgpu_mat *out = nullptr;
    cbgpu_prune_stats ps;
    memset(&ps, 0, sizeof(ps));
    Timer tm(c->stream);
    tm.start();
    int rc = prune_distributed(c, comm, *Cp, hard, sel, rec, pct, &out, &ps, &bytes);
    ms->ms_prune += tm.stop();
    if (rc != CBGPU_OK) return rc;
    ms->cols_recovered += ps.cols_recovered;
    ms->cols_selected += ps.cols_selected;
    ms->cols_recovered_after_select += ps.cols_recovered_after_select;
    mat_release(c, *Cp);
    *Cp = out;
    return CBGPU_OK;
  }
};

/* The phased distributed multiply WITH its pruning epilogue: replaces MemEfficientSpGEMM (ParFriends.h:452-777) on a 2D grid
 * and MemEfficientSpGEMM3D (:3673-4170, pruning :4148) on a layered one. Every finished piece of C is pruned
 * (MCLPruneRecoverySelect, column decisions over the whole distributed column) before the next slab is multiplied; the pruned
 * pieces are concatenated into this rank's block of C in the layout of Mult_AnXBn_Synch / Mult_AnXBn_SUMMA3D. phases <= 0: from
 * the distributed symbolic pass, so that an unpruned piece takes at most a quarter of the free HBM of the fullest rank. */
int cbgpu_memefficient_spgemm_dist(cbgpu_ctx *ctx, cbgpu_comm *comm, int semiring, const cbgpu_mat *A, const cbgpu_mat *B, int phases,
                                   double hardThreshold, int64_t selectNum, int64_t recoverNum, double recoverPct, cbgpu_mat **C,
                                   cbgpu_memeff_stats *stats, cbgpu_dist_stats *dstats) {
  if (!ctx || !comm || !A || !B || !C) return CBGPU_ERR_INVALID;
  CB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (comm->grid.layers > 1 && !ctx->opt.fiber_fused)
    return set_error(ctx, CBGPU_ERR_UNSUPPORTED, "the pruning driver uses the fiber-fused 3D formulation (option fiber_fused = 1)");
  cbgpu_memeff_stats ms;
  memset(&ms, 0, sizeof(ms));
  Timer all(ctx->stream);
  all.start();
  if (phases <= 0) {
    // CalculateNumberOfPhases (ParFriends.h:780-843) with exact counts: the largest piece any rank will hold decides
    int64_t flops = 0, nnz = 0;
    CB_TRY(cbgpu_summa_symbolic(ctx, comm, semiring, A, B, &flops, &nnz));
    size_t freeb = 0, totalb = 0;
    CB_CUDA(ctx, cudaMemGetInfo(&freeb, &totalb));
    int64_t want = (int64_t)ceil((double)nnz * 12.0 / (double)std::max<size_t>(freeb / 4, (size_t)1 << 30));
    if (want < 1) want = 1;
    int64_t *d = nullptr;
    CB_TRY(dev_alloc_t(ctx, &d, 2 * (size_t)comm->grid.world + 2));
    CB_CUDA(ctx, cudaMemcpyAsync(d, &want, 8, cudaMemcpyHostToDevice, ctx->stream));
    CB_NCCL(ctx, nccl().AllGather(d, d + 1, 1, ncclInt64, comm->world, ctx->stream));
    std::vector<int64_t> all_want((size_t)comm->grid.world, 1);
    CB_CUDA(ctx, cudaMemcpyAsync(all_want.data(), d + 1, 8 * (size_t)comm->grid.world, cudaMemcpyDeviceToHost, ctx->stream));
    CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CB_TRY(dev_free(ctx, d));
    for (int64_t w : all_want) want = std::max(want, w);
    phases = (int)std::min<int64_t>(want, std::max<int64_t>(1, B->n / std::max(1, comm->grid.layers)));
  }
  phases = std::max(1, phases);
  ms.phases = phases;
  PruneEpilogue ep;
  ep.comm = comm;
  ep.hard = hardThreshold;
  ep.pct = recoverPct;
  ep.sel = selectNum;
  ep.rec = recoverNum;
  ep.ms = &ms;
  std::vector<cbgpu_mat *> kept(phases, nullptr);
  std::vector<cbgpu_slab_result> res(phases);
  cbgpu_dist_stats ds;
  memset(&ds, 0, sizeof(ds));
  int rc = summa_phased_impl(ctx, comm, semiring, A, B, phases, 0, false, 0, 0, kept.data(), res.data(), &ds, &ep);
  cbgpu_mat *out = nullptr;
  if (rc == CBGPU_OK) {
    if (phases > 1) rc = mat_colconcat(ctx, phases, kept.data(), &out);
    else {
      out = kept[0];
      kept[0] = nullptr;
    }
  }
  for (cbgpu_mat *m : kept) mat_release(ctx, m);
  ms.flops = ds.local.flops;
  ms.ms_multiply = ds.ms_multiply;
  ms.ms_total = all.stop();
  ds.bytes_fiber += 0;
  if (rc == CBGPU_OK) {
    ms.nnz_out = out->nnz;
    *C = out;
  }
  if (stats) *stats = ms;
  if (dstats) {
    ds.bytes_bcast += ep.bytes; // the pruning all-to-alls travel along the process column, like the column broadcasts
    *dstats = ds;
  }
  return rc;
}

} // extern "C"
