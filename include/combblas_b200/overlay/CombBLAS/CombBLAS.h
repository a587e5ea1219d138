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
//                              MultiwayMerge MultiwayMerge.h:428-429, MultiwayMergeHash :553-554.
#ifndef CBGPU_OVERLAY_COMBBLAS_H
#define CBGPU_OVERLAY_COMBBLAS_H
#include_next "CombBLAS/CombBLAS.h"

#include <cstdio>
#include <cstdlib>
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

constexpr int sr_types[CBGPU_SR_COUNT][3] = {
    {CBGPU_F64, CBGPU_F64, CBGPU_F64}, {CBGPU_F32, CBGPU_F32, CBGPU_F32},  {CBGPU_I64, CBGPU_I64, CBGPU_I64},
    {CBGPU_BOOL, CBGPU_I64, CBGPU_I64}, {CBGPU_F64, CBGPU_F64, CBGPU_F64}, {CBGPU_BOOL, CBGPU_BOOL, CBGPU_BOOL},
    {CBGPU_BOOL, CBGPU_F64, CBGPU_F64}, {CBGPU_I32, CBGPU_I32, CBGPU_I32}, {CBGPU_I64, CBGPU_I64, CBGPU_I64}};

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
    const char *dev = std::getenv("CBGPU_DEVICE");
    int rc = cbgpu_create(dev ? std::atoi(dev) : 0, nullptr, &ctx);
    if (rc != CBGPU_OK) {
      std::fprintf(stderr, "[cbgpu overlay] no usable GPU (status %d); the device path has no CPU fallback\n", rc);
      MPI_Abort(MPI_COMM_WORLD, INVALIDPARAMS);
    }
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

} // namespace cbgpu_overlay

namespace combblas {

template <typename SR, typename NTO, typename IT, typename NT1, typename NT2>
  requires cbgpu_overlay::supported<SR, IT, NT1, NT2, NTO>
SpTuples<IT, NTO> *LocalHybridSpGEMM(const SpDCCols<IT, NT1> &A, const SpDCCols<IT, NT2> &B, bool clearA, bool clearB,
                                     IT *aux = nullptr) {
  (void)aux;
  return cbgpu_overlay::multiply<SR, NTO>(A, B, clearA, clearB);
}

template <typename SR, typename NTO, typename IT, typename NT1, typename NT2>
  requires cbgpu_overlay::supported<SR, IT, NT1, NT2, NTO>
SpTuples<IT, NTO> *LocalSpGEMMHash(const SpDCCols<IT, NT1> &A, const SpDCCols<IT, NT2> &B, bool clearA, bool clearB,
                                   bool sort = true) {
  (void)sort; // the device path always emits sorted columns (callers with sort=false accept any order)
  return cbgpu_overlay::multiply<SR, NTO>(A, B, clearA, clearB);
}

template <typename SR, typename NTO, typename IT, typename NT1, typename NT2>
  requires cbgpu_overlay::supported<SR, IT, NT1, NT2, NTO>
SpTuples<IT, NTO> *LocalSpGEMM(const SpDCCols<IT, NT1> &A, const SpDCCols<IT, NT2> &B, bool clearA, bool clearB) {
  return cbgpu_overlay::multiply<SR, NTO>(A, B, clearA, clearB);
}

template <class SR, class IT, class NT>
  requires cbgpu_overlay::mergeable<SR, IT, NT>
SpTuples<IT, NT> *MultiwayMerge(std::vector<SpTuples<IT, NT> *> &ArrSpTups, IT mdim = 0, IT ndim = 0, bool delarrs = false) {
  return cbgpu_overlay::merge<SR>(ArrSpTups, mdim, ndim, delarrs);
}

template <class SR, class IT, class NT>
  requires cbgpu_overlay::mergeable<SR, IT, NT>
SpTuples<IT, NT> *MultiwayMergeHash(std::vector<SpTuples<IT, NT> *> &ArrSpTups, IT mdim = 0, IT ndim = 0, bool delarrs = false,
                                    bool sorted = true) {
  (void)sorted;
  return cbgpu_overlay::merge<SR>(ArrSpTups, mdim, ndim, delarrs);
}

} // namespace combblas
#endif
