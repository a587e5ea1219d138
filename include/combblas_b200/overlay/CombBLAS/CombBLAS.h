// Overlay of "CombBLAS/CombBLAS.h": the reference-side binding of libcbgpu.so.
//
// Put this directory FIRST on the include path and compile the (unchanged) driver with -std=c++20:
//     g++ -std=c++20 -I<repo>/include/combblas_b200/overlay -I<repo>/include -I<CombBLAS>/include ... driver.cpp \
//         -L<repo>/combblas_b200 -lcbgpu
// The real header is pulled in with #include_next; afterwards this file declares, in namespace combblas, function
// templates with the IDENTICAL signatures of the reference's local SpGEMM and merge entry points plus a `requires`
// clause. Every call site in the reference (Mult_AnXBn_Synch ParFriends.h:1516,:1548, MemEfficientSpGEMM :659,:714,
// Mult_AnXBn_SUMMA3D :3505,:3528,:3642, SUMMALayer.h:78, ...) is a dependent unqualified call on combblas types, so the
// more-constrained overload is found by ADL at instantiation and wins partial ordering whenever the
// (semiring, index, value types) combination is one the device library instantiates; otherwise the reference's own
// template is chosen, unchanged. Nothing from the reference is copied.
//
// Replaces (reference lines):  LocalHybridSpGEMM mtSpGEMM.h:213-217, LocalSpGEMMHash :463-467, LocalSpGEMM :74-78,
//                              MultiwayMerge MultiwayMerge.h:428-429, MultiwayMergeHash :553-554,
//                              and the distributed drivers themselves, so that PSpGEMM (SpParMat.h:458-471) and the HipMCL /
//                              3D drivers keep their blocks on the GPUs for the whole multiply instead of staging every SUMMA
//                              stage through the host:
//                              Mult_AnXBn_Synch ParFriends.h:1447-1449, MemEfficientSpGEMM :452-454,
//                              Mult_AnXBn_SUMMA3D :3374-3375, MemEfficientSpGEMM3D :3673-3675.
// Run-time switches: CBGPU_DISABLE=1 sends every call to the reference's own templates; CBGPU_DEVICE picks the GPU
// (default: the node-local rank the MPI launcher exports, else rank modulo the number of devices).
#ifndef CBGPU_OVERLAY_COMBBLAS_H
#define CBGPU_OVERLAY_COMBBLAS_H
#include_next "CombBLAS/CombBLAS.h"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <tuple>
#include <type_traits>
#include <vector>
#include "cbgpu.h"
#include "combblas_b200/semiring_decl.h" // dtype_of, user_semiring, CBGPU_DECLARE_SEMIRING, CBGPU_HD

namespace cbgpu_overlay {

// library semiring -> cbgpu_semiring id. A driver's own semiring struct reaches the device either by specialising this
// trait with the library id whose arithmetic it matches (an OR-AND struct -> CBGPU_SR_OR_AND_BOOL), or -- any arithmetic --
// through its own device instantiation: CBGPU_DEFINE_SEMIRING in a .cu of the application + CBGPU_DECLARE_SEMIRING here
// (combblas_b200/semiring_decl.h). Anything else stays on the reference's CPU path.
template <class SR> struct semiring_id { static constexpr int value = -1; };
template <> struct semiring_id<combblas::PlusTimesSRing<double, double>> { static constexpr int value = CBGPU_SR_PLUS_TIMES_F64; };
template <> struct semiring_id<combblas::PlusTimesSRing<float, float>> { static constexpr int value = CBGPU_SR_PLUS_TIMES_F32; };
template <> struct semiring_id<combblas::PlusTimesSRing<int64_t, int64_t>> { static constexpr int value = CBGPU_SR_PLUS_TIMES_I64; };
template <> struct semiring_id<combblas::PlusTimesSRing<int32_t, int32_t>> { static constexpr int value = CBGPU_SR_PLUS_TIMES_I32; };
template <> struct semiring_id<combblas::PlusTimesSRing<bool, double>> { static constexpr int value = CBGPU_SR_PLUS_TIMES_BOOL_F64; };
template <> struct semiring_id<combblas::SelectMaxSRing<bool, int64_t>> { static constexpr int value = CBGPU_SR_SELECT_MAX_BOOL_I64; };
template <> struct semiring_id<combblas::SelectMaxSRing<int64_t, int64_t>> { static constexpr int value = CBGPU_SR_SELECT_MAX_I64; };
template <> struct semiring_id<combblas::MinPlusSRing<double, double>> { static constexpr int value = CBGPU_SR_MIN_PLUS_F64; };
// the indexing pair of SpParMat::SubsRef_SR (SpParMat.cpp:2515-2566): add must not happen, the device path checks it
template <> struct semiring_id<combblas::BoolCopy2ndSRing<double>> { static constexpr int value = CBGPU_SR_BOOL_COPY_2ND_F64; };
template <> struct semiring_id<combblas::BoolCopy1stSRing<double>> { static constexpr int value = CBGPU_SR_BOOL_COPY_1ST_F64; };
template <> struct semiring_id<combblas::BoolCopy2ndSRing<int64_t>> { static constexpr int value = CBGPU_SR_BOOL_COPY_2ND_I64; };
template <> struct semiring_id<combblas::BoolCopy1stSRing<int64_t>> { static constexpr int value = CBGPU_SR_BOOL_COPY_1ST_I64; };
template <> struct semiring_id<combblas::BoolCopy2ndSRing<bool>> { static constexpr int value = CBGPU_SR_BOOL_COPY_2ND_BOOL; };
template <> struct semiring_id<combblas::BoolCopy1stSRing<bool>> { static constexpr int value = CBGPU_SR_BOOL_COPY_1ST_BOOL; };

constexpr int sr_types[CBGPU_SR_COUNT][3] = {
    {CBGPU_F64, CBGPU_F64, CBGPU_F64}, {CBGPU_F32, CBGPU_F32, CBGPU_F32},  {CBGPU_I64, CBGPU_I64, CBGPU_I64},
    {CBGPU_BOOL, CBGPU_I64, CBGPU_I64}, {CBGPU_F64, CBGPU_F64, CBGPU_F64}, {CBGPU_BOOL, CBGPU_BOOL, CBGPU_BOOL},
    {CBGPU_BOOL, CBGPU_F64, CBGPU_F64}, {CBGPU_I32, CBGPU_I32, CBGPU_I32}, {CBGPU_I64, CBGPU_I64, CBGPU_I64},
    {CBGPU_BOOL, CBGPU_F64, CBGPU_F64}, {CBGPU_F64, CBGPU_BOOL, CBGPU_F64}, {CBGPU_BOOL, CBGPU_I64, CBGPU_I64},
    {CBGPU_I64, CBGPU_BOOL, CBGPU_I64}, {CBGPU_BOOL, CBGPU_BOOL, CBGPU_BOOL}, {CBGPU_BOOL, CBGPU_BOOL, CBGPU_BOOL}};

// value types the device instantiation of SR works on: the library's table, or what CBGPU_DECLARE_SEMIRING stated
template <class SR, int WHICH>
constexpr int sr_dtype() {
  if constexpr (semiring_id<SR>::value >= 0) return sr_types[semiring_id<SR>::value][WHICH];
  else if constexpr (user_semiring<SR>::value)
    return WHICH == 0 ? user_semiring<SR>::a_dtype : (WHICH == 1 ? user_semiring<SR>::b_dtype : user_semiring<SR>::c_dtype);
  else return -2;
}
template <class SR>
inline int runtime_id() {
  if constexpr (semiring_id<SR>::value >= 0) return semiring_id<SR>::value;
  else return user_semiring<SR>::id();
}

template <class SR, class IT, class NT1, class NT2, class NTO>
concept supported = (semiring_id<SR>::value >= 0 || user_semiring<SR>::value) && (sizeof(IT) == 4 || sizeof(IT) == 8) &&
                    std::is_integral_v<IT> && (dtype_of<NT1>::value == sr_dtype<SR, 0>()) &&
                    (dtype_of<NT2>::value == sr_dtype<SR, 1>()) && (dtype_of<NTO>::value == sr_dtype<SR, 2>());
template <class SR, class IT, class NT>
concept mergeable = (semiring_id<SR>::value >= 0 || user_semiring<SR>::value) && (sizeof(IT) == 4 || sizeof(IT) == 8) &&
                    std::is_integral_v<IT> && (dtype_of<NT>::value == sr_dtype<SR, 2>());

inline bool disabled() {
  static const bool off = std::getenv("CBGPU_DISABLE") != nullptr;
  return off;
}

inline cbgpu_ctx *context() {
  static cbgpu_ctx *ctx = nullptr;
  if (!ctx) {
    // one rank per GPU: CBGPU_DEVICE, else the node-local rank exported by the launcher, else rank modulo device count
    int device = 0, ndev = 1, rank = 0;
    cbgpu_device_count(&ndev);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    const char *names[] = {"CBGPU_DEVICE", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID"};
    const char *dev = nullptr;
    for (const char *nm : names)
      if (!dev) dev = std::getenv(nm);
    device = dev ? std::atoi(dev) : (ndev > 0 ? rank % ndev : 0);
    int rc = cbgpu_create(device, nullptr, &ctx);
    if (rc != CBGPU_OK) {
      std::fprintf(stderr, "[cbgpu overlay] no usable GPU (status %d); the device path has no CPU fallback\n", rc);
      MPI_Abort(MPI_COMM_WORLD, INVALIDPARAMS);
    }
    // CBGPU_VALIDATE: every block handed to the device is checked there first (sorted, in-range rows; cbgpu_mat_validate) --
    // reference blocks from the sort=false paths are legal on the host and wrong on the device
    if (std::getenv("CBGPU_VALIDATE")) cbgpu_set_option(ctx, "validate_uploads", 1);
  }
  return ctx;
}

// failures keep the reference's convention: message + MPI_Abort with its code (SpDefs.h:72-78)
inline void check(cbgpu_ctx *ctx, int rc) {
  if (rc == CBGPU_OK) return;
  std::fprintf(stderr, "[cbgpu overlay] %s\n", cbgpu_last_error(ctx));
  MPI_Abort(MPI_COMM_WORLD, rc == CBGPU_ERR_DIMMISMATCH ? DIMMISMATCH : INVALIDPARAMS);
}

// the id a call passes to the library; a user semiring whose device unit failed to register stops the run
template <class SR>
inline int semiring_of() {
  const int id = runtime_id<SR>();
  if (id < 0) {
    std::fprintf(stderr, "[cbgpu overlay] the device instantiation of a user semiring did not register (status %d)\n", id);
    MPI_Abort(MPI_COMM_WORLD, INVALIDPARAMS);
  }
  return id;
}

template <class IT, class NT>
cbgpu_dcsc_view view_of(const combblas::SpDCCols<IT, NT> &M) {
  cbgpu_dcsc_view v{};
  v.m = M.getnrow();
  v.n = M.getncol();
  v.idx_bytes = (int)sizeof(IT);
  v.dtype = dtype_of<NT>::value;
  if (!M.isZero()) {
    combblas::Dcsc<IT, NT> *d = M.GetDCSC();
    v.nnz = d->nz;
    v.nzc = d->nzc;
    v.cp = d->cp;
    v.jc = d->jc;
    v.ir = d->ir;
    v.numx = d->numx;
  }
  return v;
}

// resident result -> the SpTuples the reference's callers expect (column-major, rows ascending, new[] storage)
template <class IT, class NT>
combblas::SpTuples<IT, NT> *tuples_of(cbgpu_ctx *ctx, cbgpu_mat *C) {
  cbgpu_mat_info_t inf;
  check(ctx, cbgpu_mat_info(C, &inf));
  std::vector<IT> rows((size_t)inf.nnz), cols((size_t)inf.nnz);
  typedef std::conditional_t<std::is_same_v<NT, bool>, unsigned char, NT> store_t;
  std::vector<store_t> vals((size_t)inf.nnz);
  check(ctx, cbgpu_mat_download_coo(ctx, C, rows.data(), cols.data(), vals.data(), (int)sizeof(IT)));
  check(ctx, cbgpu_mat_free(ctx, C));
  if (inf.nnz == 0) return new combblas::SpTuples<IT, NT>(0, (IT)inf.m, (IT)inf.n);
  std::tuple<IT, IT, NT> *t = new std::tuple<IT, IT, NT>[inf.nnz];
#ifdef _OPENMP
#pragma omp parallel for
#endif
  for (int64_t i = 0; i < inf.nnz; ++i) t[i] = std::make_tuple(rows[i], cols[i], (NT)vals[i]);
  return new combblas::SpTuples<IT, NT>(inf.nnz, (IT)inf.m, (IT)inf.n, t, true, false);
}

template <class SR, class NTO, class IT, class NT1, class NT2>
combblas::SpTuples<IT, NTO> *multiply(const combblas::SpDCCols<IT, NT1> &A, const combblas::SpDCCols<IT, NT2> &B, bool clearA,
                                      bool clearB) {
  cbgpu_ctx *ctx = context();
  cbgpu_dcsc_view va = view_of(A), vb = view_of(B);
  cbgpu_mat *C = nullptr;
  check(ctx, cbgpu_spgemm_local_host(ctx, semiring_of<SR>(), &va, &vb, &C, nullptr));
  combblas::SpTuples<IT, NTO> *out = tuples_of<IT, NTO>(ctx, C);
  if (clearA) delete const_cast<combblas::SpDCCols<IT, NT1> *>(&A); // mtSpGEMM.h:443-446
  if (clearB) delete const_cast<combblas::SpDCCols<IT, NT2> *>(&B);
  return out;
}

template <class SR, class IT, class NT>
combblas::SpTuples<IT, NT> *merge(std::vector<combblas::SpTuples<IT, NT> *> &lists, IT mdim, IT ndim, bool delarrs) {
  const int k = (int)lists.size();
  if (k == 0) return new combblas::SpTuples<IT, NT>(0, mdim, ndim); // MultiwayMerge.h:433-436
  if (k == 1 && delarrs) return lists[0];                           // steal, MultiwayMerge.h:437-442
  cbgpu_ctx *ctx = context();
  std::vector<cbgpu_mat *> dev(k, nullptr);
  for (int i = 0; i < k; ++i) {
    combblas::SpDCCols<IT, NT> D(*lists[i], false); // tuples -> DCSC on the host (SpDCCols.cpp:110)
    cbgpu_dcsc_view v = view_of(D);
    check(ctx, cbgpu_mat_upload(ctx, &v, &dev[i]));
  }
  cbgpu_mat *C = nullptr;
  check(ctx, cbgpu_merge(ctx, semiring_of<SR>(), k, dev.data(), &C, nullptr));
  for (int i = 0; i < k; ++i) cbgpu_mat_free(ctx, dev[i]);
  combblas::SpTuples<IT, NT> *out = tuples_of<IT, NT>(ctx, C);
  if (delarrs)
    for (int i = 0; i < k; ++i) delete lists[i];
  return out;
}

// ------------------------------------------------------------------------------------------------ distributed seams
// a semiring the overlay does not know: same static members, so the reference's own templates compute the same thing.
// Used to hand a call back to the reference at run time (CBGPU_DISABLE, grids the device library does not map).
template <class SR> struct cpu_only : SR {};

template <class DER> struct der_traits { static constexpr bool ok = false; typedef void index_t; typedef void value_t; };
template <class IT, class NT> struct der_traits<combblas::SpDCCols<IT, NT>> {
  static constexpr bool ok = (sizeof(IT) == 4 || sizeof(IT) == 8) && std::is_integral_v<IT>;
  typedef IT index_t;
  typedef NT value_t;
};
// the blocks are SpDCCols with one local index type, their value types are the SpParMat's, and the local multiply is one
// the device library instantiates
template <class SR, class NUO, class UDERO, class NU1, class NU2, class UDERA, class UDERB>
concept dist_supported = der_traits<UDERA>::ok && der_traits<UDERB>::ok && der_traits<UDERO>::ok &&
                         std::is_same_v<typename der_traits<UDERA>::index_t, typename der_traits<UDERB>::index_t> &&
                         std::is_same_v<typename der_traits<UDERA>::index_t, typename der_traits<UDERO>::index_t> &&
                         std::is_same_v<typename der_traits<UDERA>::value_t, NU1> && std::is_same_v<typename der_traits<UDERB>::value_t, NU2> &&
                         std::is_same_v<typename der_traits<UDERO>::value_t, NUO> &&
                         supported<SR, typename der_traits<UDERA>::index_t, NU1, NU2, NUO>;

// NCCL communicators of this rank for a (world communicator, layers) pair: the unique id travels with the driver's own
// MPI_Bcast, rows / columns / fibers are split inside the library with the reference's colours (CommGrid.cpp:57-58,
// CommGrid3D.h:75-93). Returns nullptr when the library's rank map differs from the grid the driver built.
inline cbgpu_comm *communicator(cbgpu_ctx *ctx, MPI_Comm world, int layers, int my_row, int my_col, int my_layer) {
  static std::map<std::pair<long long, int>, cbgpu_comm *> cache;
  const std::pair<long long, int> key((long long)(intptr_t)world, layers);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int size = 1, rank = 0;
  MPI_Comm_size(world, &size);
  MPI_Comm_rank(world, &rank);
  cbgpu_grid g;
  cbgpu_comm *comm = nullptr;
  int ok = cbgpu_grid_make(size, rank, layers, &g) == CBGPU_OK && g.my_row == my_row && g.my_col == my_col && g.my_layer == my_layer;
  int all_ok = ok;
  MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, world);
  if (all_ok) {
    unsigned char id[128] = {0};
    if (rank == 0) check(ctx, cbgpu_nccl_unique_id(id));
    MPI_Bcast(id, 128, MPI_BYTE, 0, world);
    check(ctx, cbgpu_comm_create(ctx, &g, id, &comm));
  }
  cache[key] = comm;
  return comm;
}

template <class IT, class NT>
cbgpu_mat *resident(cbgpu_ctx *ctx, const combblas::SpDCCols<IT, NT> &M) {
  cbgpu_dcsc_view v = view_of(M);
  cbgpu_mat *d = nullptr;
  check(ctx, cbgpu_mat_upload(ctx, &v, &d));
  return d;
}

// resident result -> a new SpDCCols block (DCSC arrays copied straight into the reference's Dcsc: no tuples, no host sort)
template <class DER>
DER *block_of(cbgpu_ctx *ctx, cbgpu_mat *C) {
  typedef typename der_traits<DER>::index_t IT;
  cbgpu_mat_info_t inf;
  check(ctx, cbgpu_mat_info(C, &inf));
  DER *out = new DER((IT)inf.nnz, (IT)inf.m, (IT)inf.n, (IT)inf.nzc); // SpDCCols.cpp:56: allocates the Dcsc when nnz > 0
  if (inf.nnz > 0) {
    auto *d = out->GetDCSC();
    cbgpu_dcsc_out o{d->cp, d->jc, d->ir, d->numx, (int)sizeof(IT)};
    check(ctx, cbgpu_mat_download(ctx, C, &o));
  }
  check(ctx, cbgpu_mat_free(ctx, C));
  return out;
}

} // namespace cbgpu_overlay

namespace combblas {

// ---- Mult_AnXBn_Synch (ParFriends.h:1447-1556): what PSpGEMM calls. Both local blocks go to HBM once, the SUMMA stages run
// device to device over NCCL (cbgpu_summa2d), one DCSC block comes back.
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
  requires cbgpu_overlay::dist_supported<SR, NUO, UDERO, NU1, NU2, UDERA, UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_Synch(SpParMat<IU, NU1, UDERA> &A, SpParMat<IU, NU2, UDERB> &B, bool clearA = false,
                                          bool clearB = false) {
  namespace ov = cbgpu_overlay;
  if (ov::disabled()) return Mult_AnXBn_Synch<ov::cpu_only<SR>, NUO, UDERO>(A, B, clearA, clearB);
  if (A.getncol() != B.getnrow()) { // CheckSpGEMMCompliance, ParFriends.h:133-147
    std::fprintf(stderr, "Can not multiply, dimensions does not match\n%lld != %lld\n", (long long)A.getncol(), (long long)B.getnrow());
    MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
  }
  int stages, dummy;
  std::shared_ptr<CommGrid> GridC = ProductGrid(A.getcommgrid().get(), B.getcommgrid().get(), stages, dummy, dummy);
  cbgpu_ctx *ctx = ov::context();
  cbgpu_comm *comm = ov::communicator(ctx, GridC->GetWorld(), 1, GridC->GetRankInProcCol(), GridC->GetRankInProcRow(), 0);
  if (!comm) return Mult_AnXBn_Synch<ov::cpu_only<SR>, NUO, UDERO>(A, B, clearA, clearB);
  cbgpu_mat *dA = ov::resident(ctx, A.seq()), *dB = ov::resident(ctx, B.seq()), *dC = nullptr;
  if (clearA) A.FreeMemory(); // ParFriends.h:1531-1540
  if (clearB) B.FreeMemory();
  ov::check(ctx, cbgpu_summa2d(ctx, comm, ov::semiring_of<SR>(), dA, dB, &dC, nullptr));
  cbgpu_mat_free(ctx, dA);
  cbgpu_mat_free(ctx, dB);
  return SpParMat<IU, NUO, UDERO>(ov::block_of<UDERO>(ctx, dC), GridC);
}

// ---- Mult_AnXBn_DoubleBuff (ParFriends.h:1238-1440; what SpParMat::SubsRef_SR calls with the BoolCopy semirings,
// SpParMat.cpp:2515-2566) and Mult_AnXBn_Overlap (:1562-1690, "not stable" in the reference). Both compute the same product
// as Mult_AnXBn_Synch and differ only in how the host overlaps its broadcasts with its multiplies (split halves / Ibcast).
// On the device the stages of a SUMMA are one stacked multiply after device-to-device broadcasts that cost 0.5-2 % of a
// step, so all three names reach the same entry.
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
  requires cbgpu_overlay::dist_supported<SR, NUO, UDERO, NU1, NU2, UDERA, UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_DoubleBuff(SpParMat<IU, NU1, UDERA> &A, SpParMat<IU, NU2, UDERB> &B, bool clearA = false,
                                               bool clearB = false) {
  if (cbgpu_overlay::disabled()) return Mult_AnXBn_DoubleBuff<cbgpu_overlay::cpu_only<SR>, NUO, UDERO>(A, B, clearA, clearB);
  return Mult_AnXBn_Synch<SR, NUO, UDERO>(A, B, clearA, clearB);
}
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
  requires cbgpu_overlay::dist_supported<SR, NUO, UDERO, NU1, NU2, UDERA, UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_Overlap(SpParMat<IU, NU1, UDERA> &A, SpParMat<IU, NU2, UDERB> &B, bool clearA = false,
                                            bool clearB = false) {
  if (cbgpu_overlay::disabled()) return Mult_AnXBn_Overlap<cbgpu_overlay::cpu_only<SR>, NUO, UDERO>(A, B, clearA, clearB);
  return Mult_AnXBn_Synch<SR, NUO, UDERO>(A, B, clearA, clearB);
}

// ---- MemEfficientSpGEMM (ParFriends.h:452-777): HipMCL's expansion. Column phases of B, every finished piece of C pruned on the
// device over whole distributed columns (MCLPruneRecoverySelect :186-354, called at :744) before the next phase multiplies;
// only the pruned block leaves the GPUs. perProcessMemory > 0 asks for an automatic phase count (:504-551): it comes from
// the exact distributed symbolic pass and the memory of the GPU. kselectVersion / computationKernel pick between
// host algorithms with identical results and have no counterpart here.
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
  requires cbgpu_overlay::dist_supported<SR, NUO, UDERO, NU1, NU2, UDERA, UDERB> && std::is_floating_point_v<NUO>
SpParMat<IU, NUO, UDERO> MemEfficientSpGEMM(SpParMat<IU, NU1, UDERA> &A, SpParMat<IU, NU2, UDERB> &B, int phases, NUO hardThreshold,
                                            IU selectNum, IU recoverNum, NUO recoverPct, int kselectVersion, int computationKernel,
                                            int64_t perProcessMemory) {
  namespace ov = cbgpu_overlay;
  if (ov::disabled())
    return MemEfficientSpGEMM<ov::cpu_only<SR>, NUO, UDERO>(A, B, phases, hardThreshold, selectNum, recoverNum, recoverPct,
                                                             kselectVersion, computationKernel, perProcessMemory);
  if (A.getncol() != B.getnrow()) {
    std::fprintf(stderr, "Can not multiply, dimensions does not match\n%lld != %lld\n", (long long)A.getncol(), (long long)B.getnrow());
    MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
  }
  if (phases < 1 || phases >= A.getncol()) phases = 1; // ParFriends.h:474-478
  if (perProcessMemory > 0) phases = 0;                // automatic
  int stages, dummy;
  std::shared_ptr<CommGrid> GridC = ProductGrid(A.getcommgrid().get(), B.getcommgrid().get(), stages, dummy, dummy);
  cbgpu_ctx *ctx = ov::context();
  cbgpu_comm *comm = ov::communicator(ctx, GridC->GetWorld(), 1, GridC->GetRankInProcCol(), GridC->GetRankInProcRow(), 0);
  if (!comm)
    return MemEfficientSpGEMM<ov::cpu_only<SR>, NUO, UDERO>(A, B, phases < 1 ? 1 : phases, hardThreshold, selectNum, recoverNum,
                                                             recoverPct, kselectVersion, computationKernel, perProcessMemory);
  cbgpu_mat *dA = ov::resident(ctx, A.seq()), *dB = ov::resident(ctx, B.seq()), *dC = nullptr;
  ov::check(ctx, cbgpu_memefficient_spgemm_dist(ctx, comm, ov::semiring_of<SR>(), dA, dB, phases, (double)hardThreshold,
                                                (int64_t)selectNum, (int64_t)recoverNum, (double)recoverPct, &dC, nullptr, nullptr));
  cbgpu_mat_free(ctx, dA);
  cbgpu_mat_free(ctx, dB);
  (void)kselectVersion;
  (void)computationKernel;
  return SpParMat<IU, NUO, UDERO>(ov::block_of<UDERO>(ctx, dC), GridC);
}

// ---- Mult_AnXBn_SUMMA3D (ParFriends.h:3374-3667): A column-split, B row-split over the layers, C column-split as A.
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDER1, typename UDER2>
  requires cbgpu_overlay::dist_supported<SR, NUO, UDERO, NU1, NU2, UDER1, UDER2>
SpParMat3D<IU, NUO, UDERO> Mult_AnXBn_SUMMA3D(SpParMat3D<IU, NU1, UDER1> &A, SpParMat3D<IU, NU2, UDER2> &B) {
  namespace ov = cbgpu_overlay;
  std::shared_ptr<CommGrid3D> g3 = A.getcommgrid3D();
  std::shared_ptr<CommGrid> layer = g3->GetCommGridLayer();
  cbgpu_ctx *ctx = ov::disabled() ? nullptr : ov::context();
  cbgpu_comm *comm = nullptr;
  if (ctx && !A.isSpecial() && A.isColSplit() && !B.isColSplit())
    comm = ov::communicator(ctx, g3->GetWorld(), g3->GetGridLayers(), layer->GetRankInProcCol(), layer->GetRankInProcRow(), g3->GetRankInFiber());
  if (!comm) return Mult_AnXBn_SUMMA3D<ov::cpu_only<SR>, NUO, UDERO>(A, B);
  if (A.getncol() != B.getnrow()) {
    std::fprintf(stderr, "Can not multiply, dimensions does not match\n%lld != %lld\n", (long long)A.getncol(), (long long)B.getnrow());
    MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
  }
  cbgpu_mat *dA = ov::resident(ctx, A.GetLayerMat()->seq()), *dB = ov::resident(ctx, B.GetLayerMat()->seq()), *dC = nullptr;
  ov::check(ctx, cbgpu_summa3d(ctx, comm, ov::semiring_of<SR>(), dA, dB, &dC, nullptr));
  cbgpu_mat_free(ctx, dA);
  cbgpu_mat_free(ctx, dB);
  std::shared_ptr<CommGrid3D> grid3d(new CommGrid3D(g3->GetWorld(), g3->GetGridLayers(), g3->GetGridRows(), g3->GetGridCols(), A.isSpecial()));
  return SpParMat3D<IU, NUO, UDERO>(ov::block_of<UDERO>(ctx, dC), grid3d, A.isColSplit(), A.isSpecial()); // ParFriends.h:3663-3665
}

// ---- MemEfficientSpGEMM3D (ParFriends.h:3673-4170): the phased 3D multiply with the pruning of every piece (:4148).
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
  requires cbgpu_overlay::dist_supported<SR, NUO, UDERO, NU1, NU2, UDERA, UDERB> && std::is_floating_point_v<NUO>
SpParMat3D<IU, NUO, UDERO> MemEfficientSpGEMM3D(SpParMat3D<IU, NU1, UDERA> &A, SpParMat3D<IU, NU2, UDERB> &B, int phases,
                                                NUO hardThreshold, IU selectNum, IU recoverNum, NUO recoverPct, int kselectVersion,
                                                int computationKernel, int64_t perProcessMemory) {
  namespace ov = cbgpu_overlay;
  std::shared_ptr<CommGrid3D> g3 = A.getcommgrid3D();
  std::shared_ptr<CommGrid> layer = g3->GetCommGridLayer();
  cbgpu_ctx *ctx = ov::disabled() ? nullptr : ov::context();
  cbgpu_comm *comm = nullptr;
  if (ctx && !A.isSpecial() && A.isColSplit() && !B.isColSplit())
    comm = ov::communicator(ctx, g3->GetWorld(), g3->GetGridLayers(), layer->GetRankInProcCol(), layer->GetRankInProcRow(), g3->GetRankInFiber());
  if (!comm)
    return MemEfficientSpGEMM3D<ov::cpu_only<SR>, NUO, UDERO>(A, B, phases, hardThreshold, selectNum, recoverNum, recoverPct,
                                                               kselectVersion, computationKernel, perProcessMemory);
  if (A.getncol() != B.getnrow()) {
    std::fprintf(stderr, "Can not multiply, dimensions does not match\n%lld != %lld\n", (long long)A.getncol(), (long long)B.getnrow());
    MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
  }
  if (phases < 1 || phases >= B.getncol()) phases = 1; // ParFriends.h:3694-3697
  if (perProcessMemory > 0) phases = 0;
  cbgpu_mat *dA = ov::resident(ctx, A.GetLayerMat()->seq()), *dB = ov::resident(ctx, B.GetLayerMat()->seq()), *dC = nullptr;
  ov::check(ctx, cbgpu_memefficient_spgemm_dist(ctx, comm, ov::semiring_of<SR>(), dA, dB, phases, (double)hardThreshold,
                                                (int64_t)selectNum, (int64_t)recoverNum, (double)recoverPct, &dC, nullptr, nullptr));
  cbgpu_mat_free(ctx, dA);
  cbgpu_mat_free(ctx, dB);
  (void)kselectVersion;
  (void)computationKernel;
  std::shared_ptr<CommGrid3D> grid3d(new CommGrid3D(g3->GetWorld(), g3->GetGridLayers(), g3->GetGridRows(), g3->GetGridCols(), A.isSpecial()));
  return SpParMat3D<IU, NUO, UDERO>(ov::block_of<UDERO>(ctx, dC), grid3d, A.isColSplit(), A.isSpecial());
}

} // namespace combblas

namespace cbgpu_overlay {
} // namespace cbgpu_overlay

namespace combblas {

template <typename SR, typename NTO, typename IT, typename NT1, typename NT2>
  requires cbgpu_overlay::supported<SR, IT, NT1, NT2, NTO>
SpTuples<IT, NTO> *LocalHybridSpGEMM(const SpDCCols<IT, NT1> &A, const SpDCCols<IT, NT2> &B, bool clearA, bool clearB,
                                     IT *aux = nullptr) {
  if (cbgpu_overlay::disabled()) return LocalHybridSpGEMM<cbgpu_overlay::cpu_only<SR>, NTO>(A, B, clearA, clearB, aux);
  return cbgpu_overlay::multiply<SR, NTO>(A, B, clearA, clearB);
}

template <typename SR, typename NTO, typename IT, typename NT1, typename NT2>
  requires cbgpu_overlay::supported<SR, IT, NT1, NT2, NTO>
SpTuples<IT, NTO> *LocalSpGEMMHash(const SpDCCols<IT, NT1> &A, const SpDCCols<IT, NT2> &B, bool clearA, bool clearB,
                                   bool sort = true) {
  // the device path always emits sorted columns (callers with sort=false accept any order)
  if (cbgpu_overlay::disabled()) return LocalSpGEMMHash<cbgpu_overlay::cpu_only<SR>, NTO>(A, B, clearA, clearB, sort);
  return cbgpu_overlay::multiply<SR, NTO>(A, B, clearA, clearB);
}

template <typename SR, typename NTO, typename IT, typename NT1, typename NT2>
  requires cbgpu_overlay::supported<SR, IT, NT1, NT2, NTO>
SpTuples<IT, NTO> *LocalSpGEMM(const SpDCCols<IT, NT1> &A, const SpDCCols<IT, NT2> &B, bool clearA, bool clearB) {
  if (cbgpu_overlay::disabled()) return LocalSpGEMM<cbgpu_overlay::cpu_only<SR>, NTO>(A, B, clearA, clearB);
  return cbgpu_overlay::multiply<SR, NTO>(A, B, clearA, clearB);
}

template <class SR, class IT, class NT>
  requires cbgpu_overlay::mergeable<SR, IT, NT>
SpTuples<IT, NT> *MultiwayMerge(std::vector<SpTuples<IT, NT> *> &ArrSpTups, IT mdim = 0, IT ndim = 0, bool delarrs = false) {
  if (cbgpu_overlay::disabled()) return MultiwayMerge<cbgpu_overlay::cpu_only<SR>>(ArrSpTups, mdim, ndim, delarrs);
  return cbgpu_overlay::merge<SR>(ArrSpTups, mdim, ndim, delarrs);
}

template <class SR, class IT, class NT>
  requires cbgpu_overlay::mergeable<SR, IT, NT>
SpTuples<IT, NT> *MultiwayMergeHash(std::vector<SpTuples<IT, NT> *> &ArrSpTups, IT mdim = 0, IT ndim = 0, bool delarrs = false,
                                    bool sorted = true) {
  if (cbgpu_overlay::disabled()) return MultiwayMergeHash<cbgpu_overlay::cpu_only<SR>>(ArrSpTups, mdim, ndim, delarrs, sorted);
  return cbgpu_overlay::merge<SR>(ArrSpTups, mdim, ndim, delarrs);
}

} // namespace combblas
#endif
