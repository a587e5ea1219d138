/*
 * Parity oracle, "reference" kind: thin C entry points around the UNMODIFIED CombBLAS reference,
 * compiled from the sources where they lie under /root/reference (see oracle/Makefile) against the
 * single-rank mpi.h stand-in in oracle/mpi_shim/. Output goes to oracle/_ref/ only.
 *
 * TEST INFRASTRUCTURE ONLY. Used by tests/ (as the checker) and by bench.py's reference arm /
 * cpu_baseline leg (as the CPU implementation being timed). Never on the product path.
 *
 * One translation unit is compiled per semiring (-DCB_SR=<id>) so the heavy template instantiation
 * parallelises; CB_SR == -1 builds the dispatcher.
 */
#include <cstdint>
#include <cstring>
#include <vector>
#include <tuple>
#include <limits>
#include <chrono>
#include "ref_oracle.h"

struct ref_result {
  std::vector<int64_t> rows, cols;
  std::vector<unsigned char> vals;
  int value_bytes = 0;
};

#ifndef CB_SR
#error "compile with -DCB_SR=<semiring id> or -DCB_SR=-1 for the dispatcher"
#endif

#if CB_SR >= 0
#include "CombBLAS/CombBLAS.h"
using namespace combblas;

/* user-defined boolean OR-AND semiring in the style of ReleaseTests/KTipsTest.cpp:12-20 */
struct OrAndBoolSR {
  static bool id() { return false; }
  static bool returnedSAID() { return false; }
  static MPI_Op mpi_op() { return MPI_LOR; }
  static bool add(const bool &a, const bool &b) { return a || b; }
  static bool multiply(const bool &a, const bool &b) { return a && b; }
  static void axpy(bool a, const bool &x, bool &y) { y = add(y, multiply(a, x)); }
};

#if CB_SR == 0
typedef double NT1; typedef double NT2; typedef double NTO; typedef PlusTimesSRing<double, double> SR;
#elif CB_SR == 1
typedef float NT1; typedef float NT2; typedef float NTO; typedef PlusTimesSRing<float, float> SR;
#elif CB_SR == 2
typedef int64_t NT1; typedef int64_t NT2; typedef int64_t NTO; typedef PlusTimesSRing<int64_t, int64_t> SR;
#elif CB_SR == 3
typedef bool NT1; typedef int64_t NT2; typedef int64_t NTO; typedef SelectMaxSRing<bool, int64_t> SR;
#elif CB_SR == 4
typedef double NT1; typedef double NT2; typedef double NTO; typedef MinPlusSRing<double, double> SR;
#elif CB_SR == 5
typedef bool NT1; typedef bool NT2; typedef bool NTO; typedef OrAndBoolSR SR;
#elif CB_SR == 6
typedef bool NT1; typedef double NT2; typedef double NTO; typedef PlusTimesSRing<bool, double> SR;
#elif CB_SR == 7
typedef int32_t NT1; typedef int32_t NT2; typedef int32_t NTO; typedef PlusTimesSRing<int32_t, int32_t> SR;
#elif CB_SR == 8
typedef int64_t NT1; typedef int64_t NT2; typedef int64_t NTO; typedef SelectMaxSRing<int64_t, int64_t> SR;
#elif CB_SR == 9 /* the indexing pair of SpParMat::SubsRef_SR (SpParMat.cpp:2515-2566); their add throws */
typedef bool NT1; typedef double NT2; typedef double NTO; typedef BoolCopy2ndSRing<double> SR;
#elif CB_SR == 10
typedef double NT1; typedef bool NT2; typedef double NTO; typedef BoolCopy1stSRing<double> SR;
#elif CB_SR == 11
typedef bool NT1; typedef int64_t NT2; typedef int64_t NTO; typedef BoolCopy2ndSRing<int64_t> SR;
#elif CB_SR == 12
typedef int64_t NT1; typedef bool NT2; typedef int64_t NTO; typedef BoolCopy1stSRing<int64_t> SR;
#elif CB_SR == 13
typedef bool NT1; typedef bool NT2; typedef bool NTO; typedef BoolCopy2ndSRing<bool> SR;
#elif CB_SR == 14
typedef bool NT1; typedef bool NT2; typedef bool NTO; typedef BoolCopy1stSRing<bool> SR;
#else
#error "unknown CB_SR"
#endif

typedef int64_t IT;

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class NT>
SpDCCols<IT, NT> *make_dcsc(const ref_csc *M) {
  if (M->nnz == 0) return new SpDCCols<IT, NT>(0, M->m, M->n, 0);
  std::tuple<IT, IT, NT> *t = new std::tuple<IT, IT, NT>[M->nnz];
  const NT *v = static_cast<const NT *>(M->vals);
  for (int64_t j = 0; j < M->n; ++j)
    for (int64_t p = M->colptr[j]; p < M->colptr[j + 1]; ++p) t[p] = std::make_tuple((IT)M->rows[p], (IT)j, v[p]);
  SpTuples<IT, NT> tup(M->nnz, M->m, M->n, t, false, false); /* sorts column-major; owns t (new[]) */
  return new SpDCCols<IT, NT>(tup, false);
}

template <class NT>
SpTuples<IT, NT> *make_tuples(const ref_csc *M) {
  std::tuple<IT, IT, NT> *t = new std::tuple<IT, IT, NT>[M->nnz > 0 ? M->nnz : 1];
  const NT *v = static_cast<const NT *>(M->vals);
  for (int64_t j = 0; j < M->n; ++j)
    for (int64_t p = M->colptr[j]; p < M->colptr[j + 1]; ++p) t[p] = std::make_tuple((IT)M->rows[p], (IT)j, v[p]);
  return new SpTuples<IT, NT>(M->nnz, M->m, M->n, t, true, false); /* keep the caller's order */
}

template <class NT>
ref_result *to_result(SpTuples<IT, NT> &t, int canonical) {
  if (canonical) t.SortColBased();
  ref_result *r = new ref_result;
  int64_t nnz = t.getnnz();
  r->value_bytes = (int)sizeof(NT);
  r->rows.resize(nnz); r->cols.resize(nnz); r->vals.resize((size_t)nnz * sizeof(NT));
  NT *v = reinterpret_cast<NT *>(r->vals.data());
  for (int64_t i = 0; i < nnz; ++i) { r->rows[i] = t.rowindex(i); r->cols[i] = t.colindex(i); v[i] = t.numvalue(i); }
  return r;
}

} // namespace

#define CB_CAT2(a, b) a##b
#define CB_CAT(a, b) CB_CAT2(a, b)
#define CB_FN(name) CB_CAT(name, CB_SR)

extern "C" int CB_FN(ref_spgemm_sr)(int routine, const ref_csc *A, const ref_csc *B, int phases, int canonical,
                                     ref_result **out, double *seconds) {
  typedef SpDCCols<IT, NT1> DA; typedef SpDCCols<IT, NT2> DB; typedef SpDCCols<IT, NTO> DC;
  DA *a = make_dcsc<NT1>(A);
  DB *b = make_dcsc<NT2>(B);
  double t0 = 0, t1 = 0;
  if (routine < 10) {
    SpTuples<IT, NTO> *c = nullptr;
    t0 = now_s();
    switch (routine) {
      case REF_LOCAL_HYBRID: c = LocalHybridSpGEMM<SR, NTO>(*a, *b, false, false); break;
      case REF_LOCAL_HASH_SORTED: c = LocalSpGEMMHash<SR, NTO>(*a, *b, false, false, true); break;
      case REF_LOCAL_HASH_UNSORTED: c = LocalSpGEMMHash<SR, NTO>(*a, *b, false, false, false); break;
      case REF_LOCAL_HEAP: c = LocalSpGEMM<SR, NTO>(*a, *b, false, false); break;
      default: delete a; delete b; return -1;
    }
    t1 = now_s();
    *out = to_result(*c, canonical);
    delete c; delete a; delete b;
  } else {
#if CB_SR == 0 || CB_SR == 1
    std::shared_ptr<CommGrid> grid;
    grid.reset(new CommGrid(MPI_COMM_WORLD, 0, 0));
    SpParMat<IT, NT1, DA> PA(a, grid);
    SpParMat<IT, NT2, DB> PB(b, grid);
    if (routine == REF_DIST_SUMMA3D) {
      SpParMat3D<IT, NT1, DA> A3(PA, 1, true, false);
      SpParMat3D<IT, NT2, DB> B3(PB, 1, false, false);
      t0 = now_s();
      SpParMat3D<IT, NTO, DC> C3 = Mult_AnXBn_SUMMA3D<SR, NTO, DC>(A3, B3);
      t1 = now_s();
      SpParMat<IT, NTO, DC> C2 = C3.Convert2D();
      SpTuples<IT, NTO> ct(*C2.seqptr());
      *out = to_result(ct, canonical);
    } else {
      t0 = now_s();
      SpParMat<IT, NTO, DC> PC =
          routine == REF_DIST_SYNCH ? Mult_AnXBn_Synch<SR, NTO, DC>(PA, PB)
          : routine == REF_DIST_DOUBLEBUFF
              ? Mult_AnXBn_DoubleBuff<SR, NTO, DC>(PA, PB)
              : MemEfficientSpGEMM<SR, NTO, DC>(PA, PB, phases, std::numeric_limits<NTO>::lowest(),
                                                 (IT)std::numeric_limits<int32_t>::max(), (IT)0, (NTO)0, 1,
                                                 routine == REF_DIST_MEMEFF_HASH ? 1 : 2, (int64_t)0);
      t1 = now_s();
      SpTuples<IT, NTO> ct(*PC.seqptr());
      *out = to_result(ct, canonical);
    }
#else
    (void)phases; delete a; delete b; return -2; /* distributed drivers are instantiated for PlusTimes float types */
#endif
  }
  if (seconds) *seconds = t1 - t0;
  return 0;
}

extern "C" int CB_FN(ref_merge_sr)(int hash, int k, const ref_csc *lists, int sorted, int canonical, ref_result **out,
                                    double *seconds) {
#if CB_SR == 5
  /* the reference's MultiwayMergeHash does not terminate for NT=bool (observed in the build container);
   * the heap MultiwayMerge is the bool oracle. */
  if (hash) return -4;
#endif
  std::vector<SpTuples<IT, NTO> *> arr;
  IT m = 0, n = 0;
  for (int i = 0; i < k; ++i) { arr.push_back(make_tuples<NTO>(&lists[i])); m = lists[i].m; n = lists[i].n; }
  double t0 = now_s();
  SpTuples<IT, NTO> *c = hash ? MultiwayMergeHash<SR>(arr, m, n, false, sorted != 0) : MultiwayMerge<SR>(arr, m, n, false);
  double t1 = now_s();
  *out = to_result(*c, canonical);
  delete c;
  for (auto p : arr) delete p;
  if (seconds) *seconds = t1 - t0;
  return 0;
}

extern "C" int CB_FN(ref_symbolic_sr)(const ref_csc *A, const ref_csc *B, int64_t *nzc, int64_t **flop, int64_t **nnz) {
  SpDCCols<IT, NT1> *a = make_dcsc<NT1>(A);
  SpDCCols<IT, NT2> *b = make_dcsc<NT2>(B);
  *nzc = 0; *flop = nullptr; *nnz = nullptr;
  if (!a->isZero() && !b->isZero()) {
    IT *aux = nullptr;
    a->GetDCSC()->ConstructAux(a->getncol(), aux);
    *nzc = b->GetDCSC()->nzc;
    *flop = estimateFLOP(*a, *b, aux);
    *nnz = estimateNNZ_Hash(*a, *b, *flop, aux);
    delete[] aux;
  }
  delete a; delete b;
  return 0;
}

/* MCLPruneRecoverySelect (ParFriends.h:186-354) on a P=1 SpParMat, and MemEfficientSpGEMM (ParFriends.h:453-777) with its
 * pruning parameters. Instantiated for the floating-point PlusTimes semirings only (what MCL.cpp uses). */
extern "C" int CB_FN(ref_mcl_prune_sr)(const ref_csc *A, double hard, int64_t select, int64_t recover, double pct,
                                        int kselect_version, ref_result **out) {
#if CB_SR == 0 || CB_SR == 1
  typedef SpDCCols<IT, NTO> DC;
  std::shared_ptr<CommGrid> grid;
  grid.reset(new CommGrid(MPI_COMM_WORLD, 0, 0));
  SpParMat<IT, NTO, DC> PA(make_dcsc<NTO>(A), grid);
  MCLPruneRecoverySelect(PA, (NTO)hard, (IT)select, (IT)recover, (NTO)pct, kselect_version);
  SpTuples<IT, NTO> ct(*PA.seqptr());
  *out = to_result(ct, 1);
  return 0;
#else
  (void)A; (void)hard; (void)select; (void)recover; (void)pct; (void)kselect_version; (void)out;
  return -2;
#endif
}

extern "C" int CB_FN(ref_memeff_prune_sr)(const ref_csc *A, const ref_csc *B, int phases, double hard, int64_t select,
                                           int64_t recover, double pct, int kselect_version, int kernel, ref_result **out) {
#if CB_SR == 0 || CB_SR == 1
  typedef SpDCCols<IT, NT1> DA; typedef SpDCCols<IT, NT2> DB; typedef SpDCCols<IT, NTO> DC;
  std::shared_ptr<CommGrid> grid;
  grid.reset(new CommGrid(MPI_COMM_WORLD, 0, 0));
  SpParMat<IT, NT1, DA> PA(make_dcsc<NT1>(A), grid);
  SpParMat<IT, NT2, DB> PB(make_dcsc<NT2>(B), grid);
  SpParMat<IT, NTO, DC> PC = MemEfficientSpGEMM<SR, NTO, DC>(PA, PB, phases, (NTO)hard, (IT)select, (IT)recover, (NTO)pct,
                                                             kselect_version, kernel, (int64_t)0);
  SpTuples<IT, NTO> ct(*PC.seqptr());
  *out = to_result(ct, 1);
  return 0;
#else
  (void)A; (void)B; (void)phases; (void)hard; (void)select; (void)recover; (void)pct; (void)kselect_version; (void)kernel; (void)out;
  return -2;
#endif
}

#else /* ------------------------------- dispatcher ------------------------------- */
#include <omp.h>
int cblas_splits = 1; /* every CombBLAS program defines this (CombBLAS.h:76) */

#define DECL(i)                                                                                                        \
  extern "C" int ref_spgemm_sr##i(int, const ref_csc *, const ref_csc *, int, int, ref_result **, double *);           \
  extern "C" int ref_merge_sr##i(int, int, const ref_csc *, int, int, ref_result **, double *);                        \
  extern "C" int ref_symbolic_sr##i(const ref_csc *, const ref_csc *, int64_t *, int64_t **, int64_t **);                \
  extern "C" int ref_mcl_prune_sr##i(const ref_csc *, double, int64_t, int64_t, double, int, ref_result **);              \
  extern "C" int ref_memeff_prune_sr##i(const ref_csc *, const ref_csc *, int, double, int64_t, int64_t, double, int, int, \
                                        ref_result **);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5) DECL(6) DECL(7) DECL(8)
DECL(9) DECL(10) DECL(11) DECL(12) DECL(13) DECL(14)
#define CASE(i, call) case i: return call;
#define ALL(fn, ...)                                                                                                   \
  switch (semiring) {                                                                                                  \
    case 0: return fn##0(__VA_ARGS__); case 1: return fn##1(__VA_ARGS__); case 2: return fn##2(__VA_ARGS__);           \
    case 3: return fn##3(__VA_ARGS__); case 4: return fn##4(__VA_ARGS__); case 5: return fn##5(__VA_ARGS__);           \
    case 6: return fn##6(__VA_ARGS__); case 7: return fn##7(__VA_ARGS__); case 8: return fn##8(__VA_ARGS__);           \
    case 9: return fn##9(__VA_ARGS__); case 10: return fn##10(__VA_ARGS__); case 11: return fn##11(__VA_ARGS__);       \
    case 12: return fn##12(__VA_ARGS__); case 13: return fn##13(__VA_ARGS__); case 14: return fn##14(__VA_ARGS__);     \
    default: return -3;                                                                                                \
  }

extern "C" {
int ref_spgemm(int routine, int semiring, const ref_csc *A, const ref_csc *B, int phases, int canonical,
               ref_result **out, double *seconds) {
  ALL(ref_spgemm_sr, routine, A, B, phases, canonical, out, seconds)
}
int ref_merge(int hash, int semiring, int k, const ref_csc *lists, int sorted, int canonical, ref_result **out,
              double *seconds) {
  ALL(ref_merge_sr, hash, k, lists, sorted, canonical, out, seconds)
}
int ref_symbolic(int semiring, const ref_csc *A, const ref_csc *B, int64_t *nzc, int64_t **flop, int64_t **nnz) {
  ALL(ref_symbolic_sr, A, B, nzc, flop, nnz)
}
int ref_mcl_prune(int semiring, const ref_csc *A, double hard, int64_t select, int64_t recover, double pct,
                  int kselect_version, ref_result **out) {
  ALL(ref_mcl_prune_sr, A, hard, select, recover, pct, kselect_version, out)
}
int ref_memeff_prune(int semiring, const ref_csc *A, const ref_csc *B, int phases, double hard, int64_t select,
                     int64_t recover, double pct, int kselect_version, int kernel, ref_result **out) {
  ALL(ref_memeff_prune_sr, A, B, phases, hard, select, recover, pct, kselect_version, kernel, out)
}
void ref_free(void *p) { delete[] static_cast<int64_t *>(p); }
int64_t ref_result_nnz(const ref_result *r) { return (int64_t)r->rows.size(); }
int ref_result_value_bytes(const ref_result *r) { return r->value_bytes; }
void ref_result_copy(const ref_result *r, int64_t *rows, int64_t *cols, void *vals) {
  if (r->rows.empty()) return;
  std::memcpy(rows, r->rows.data(), r->rows.size() * sizeof(int64_t));
  std::memcpy(cols, r->cols.data(), r->cols.size() * sizeof(int64_t));
  std::memcpy(vals, r->vals.data(), r->vals.size());
}
void ref_result_free(ref_result *r) { delete r; }
int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }
}
#endif
