/*
 * Plain-C restatement ("port") of the reference's local semiring SpGEMM hot path.
 *
 * TEST INFRASTRUCTURE ONLY: the checker used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg. It is never linked into, loaded by, or called from the product.
 *
 * Parity pinning (see spgemm_oracle.h): checked against the unmodified reference compiled into
 * oracle/_ref (tests/test_oracle.py) and against the reference tree's known-answer vector
 * 3DSpGEMM/matlab/{bcsstk01,C}.mtx (fixtures under tests/golden/).
 *
 * What is restated, with the reference lines each part follows (paths relative to the reference root):
 *   column lookup            dcsc.cpp:1363 FillColInds  -> here a dense colptr lookup (same ranges)
 *   flop per column          include/CombBLAS/mtSpGEMM.h:1112-1130  estimateFLOP
 *   symbolic nnz per column  include/CombBLAS/mtSpGEMM.h:861-928    estimateNNZ_Hash
 *   exclusive scan           include/CombBLAS/mtSpGEMM.h:24-70      prefixsum
 *   numeric hash column      include/CombBLAS/mtSpGEMM.h:552-634    LocalSpGEMMHash (== :362-440 hybrid hash branch)
 *   merge column             include/CombBLAS/MultiwayMerge.h:338-422 SerialMergeHash (driver :554-701)
 *   semirings                include/CombBLAS/Semirings.h:143-255, ReleaseTests/KTipsTest.cpp:12-20
 *
 * Hash tables: size = smallest power of two >= max(16, count); slot = (key*107) & (size-1); linear probing;
 * empty key = -1; the first product is stored as is, later ones as add(product, stored) -- the same argument
 * order as the reference, so non-commutative add functions behave identically; the floating point
 * accumulation order is the storage order of B(:,j), as in the reference.
 */
#define _POSIX_C_SOURCE 200809L
#include "spgemm_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct port_result {
  int64_t nnz;
  int vbytes;
  int64_t *rows, *cols;
  void *vals;
};

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static size_t table_size(int64_t count) { /* mtSpGEMM.h:884-888 / :370-374 */
  size_t s = 16;
  while ((int64_t)s < count) s <<= 1;
  return s;
}

/* mtSpGEMM.h:1112-1130 */
static void flops_per_column(const port_csc *A, const port_csc *B, int64_t *flop) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t j = 0; j < B->n; ++j) {
    int64_t f = 0;
    for (int64_t p = B->colptr[j]; p < B->colptr[j + 1]; ++p) {
      int64_t k = B->rows[p];
      f += A->colptr[k + 1] - A->colptr[k];
    }
    flop[j] = f;
  }
}

/* mtSpGEMM.h:861-928: distinct row ids among the products of one column */
int64_t port_symbolic(const port_csc *A, const port_csc *B, int64_t *flop, int64_t *nnz) {
  flops_per_column(A, B, flop);
  int64_t total = 0;
  for (int64_t j = 0; j < B->n; ++j) total += flop[j];
#pragma omp parallel
  {
    int64_t *tab = NULL;
    size_t cap = 0;
#pragma omp for schedule(dynamic, 64)
    for (int64_t j = 0; j < B->n; ++j) {
      if (flop[j] == 0) { nnz[j] = 0; continue; }
      size_t sz = table_size(flop[j]);
      if (sz > cap) { free(tab); tab = (int64_t *)malloc(sz * sizeof(int64_t)); cap = sz; }
      for (size_t t = 0; t < sz; ++t) tab[t] = -1;
      int64_t cnt = 0;
      for (int64_t p = B->colptr[j]; p < B->colptr[j + 1]; ++p) {
        int64_t k = B->rows[p];
        for (int64_t q = A->colptr[k]; q < A->colptr[k + 1]; ++q) {
          int64_t key = A->rows[q];
          size_t h = (size_t)(key * 107) & (sz - 1);
          for (;;) {
            if (tab[h] == key) break;
            if (tab[h] == -1) { tab[h] = key; ++cnt; break; }
            h = (h + 1) & (sz - 1);
          }
        }
      }
      nnz[j] = cnt;
    }
    free(tab);
  }
  return total;
}

/* ---- semiring arithmetic (Semirings.h:143-255) ---- */
static inline double f64_mul(double a, double b) { return a * b; }
static inline double f64_add(double a, double b) { return a + b; }
static inline float f32_mul(float a, float b) { return a * b; }
static inline float f32_add(float a, float b) { return a + b; }
static inline int64_t i64_mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
static inline int64_t i64_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int32_t i32_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t i32_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int64_t i64_max(int64_t a, int64_t b) { return a > b ? a : b; }          /* std::max(arg1,arg2) */
static inline int64_t selmax_bool_mul(uint8_t a, int64_t b) { (void)a; return b; }     /* Semirings.h:200-203 */
static inline double minplus_mul(double a, double b) {                                 /* inf_plus, Semirings.h:41-47 */
  const double inf = 1.7976931348623157e308;
  return (a == inf || b == inf) ? inf : a + b;
}
static inline double f64_min(double a, double b) { return b < a ? b : a; }             /* std::min(arg1,arg2) */
static inline uint8_t bool_and(uint8_t a, uint8_t b) { return (uint8_t)(a && b); }
static inline uint8_t bool_or(uint8_t a, uint8_t b) { return (uint8_t)(a || b); }
static inline double boolf64_mul(uint8_t a, double b) { return (double)(a != 0) * b; }
/* BoolCopy2ndSRing<OUT> / BoolCopy1stSRing<OUT> (Semirings.h:51-138): multiply copies the non-boolean operand; add prints
 * "Add should not happen" and throws. Here add raises a flag and the entry point returns -2 (the reference aborts). */
static volatile int bool_copy_add_happened = 0;
static inline double copy2nd_f64(uint8_t a, double b) { (void)a; return b; }
static inline double copy1st_f64(double a, uint8_t b) { (void)b; return a; }
static inline int64_t copy2nd_i64(uint8_t a, int64_t b) { (void)a; return b; }
static inline int64_t copy1st_i64(int64_t a, uint8_t b) { (void)b; return a; }
static inline uint8_t copy2nd_u8(uint8_t a, uint8_t b) { (void)a; return b; }
static inline uint8_t copy1st_u8(uint8_t a, uint8_t b) { (void)b; return a; }
static inline double forbidden_f64(double a, double b) { (void)a; bool_copy_add_happened = 1; return b; }
static inline int64_t forbidden_i64(int64_t a, int64_t b) { (void)a; bool_copy_add_happened = 1; return b; }
static inline uint8_t forbidden_u8(uint8_t a, uint8_t b) { (void)a; bool_copy_add_happened = 1; return b; }

typedef struct { int64_t key; int64_t slot; } keyslot;
static int cmp_keyslot(const void *x, const void *y) {
  int64_t a = ((const keyslot *)x)->key, b = ((const keyslot *)y)->key;
  return (a > b) - (a < b);
}

/*
 * One generic body per (TA, TB, TC, MUL, ADD). Pass 1 = symbolic (above), pass 2 = numeric hash
 * (mtSpGEMM.h:552-634) writing each column at colptrC[j].
 */
#define DEFINE_SPGEMM(NAME, TA, TB, TC, MUL, ADD)                                                                      \
  static void spgemm_##NAME(const port_csc *A, const port_csc *B, int sort, const int64_t *colptrC,                    \
                            int64_t *crows, int64_t *ccols, TC *cvals) {                                               \
    const TA *av = (const TA *)A->vals;                                                                                \
    const TB *bv = (const TB *)B->vals;                                                                                \
    _Pragma("omp parallel") {                                                                                          \
      int64_t *keys = NULL; TC *vals = NULL; keyslot *ks = NULL; size_t cap = 0;                                       \
      _Pragma("omp for schedule(dynamic, 64)")                                                                         \
      for (int64_t j = 0; j < B->n; ++j) {                                                                             \
        int64_t nnzc = colptrC[j + 1] - colptrC[j];                                                                    \
        if (nnzc == 0) continue;                                                                                       \
        size_t sz = table_size(nnzc);                                                                                  \
        if (sz > cap) {                                                                                                \
          free(keys); free(vals); free(ks);                                                                            \
          keys = (int64_t *)malloc(sz * sizeof(int64_t)); vals = (TC *)malloc(sz * sizeof(TC));                        \
          ks = (keyslot *)malloc(sz * sizeof(keyslot)); cap = sz;                                                      \
        }                                                                                                              \
        for (size_t t = 0; t < sz; ++t) keys[t] = -1;                                                                  \
        for (int64_t p = B->colptr[j]; p < B->colptr[j + 1]; ++p) {                                                    \
          int64_t k = B->rows[p];                                                                                      \
          TB bval = bv[p];                                                                                             \
          for (int64_t q = A->colptr[k]; q < A->colptr[k + 1]; ++q) {                                                  \
            TC mrhs = MUL(av[q], bval);                                                                                \
            int64_t key = A->rows[q];                                                                                  \
            size_t h = (size_t)(key * 107) & (sz - 1);                                                                 \
            for (;;) {                                                                                                 \
              if (keys[h] == key) { vals[h] = ADD(mrhs, vals[h]); break; }                                             \
              if (keys[h] == -1) { keys[h] = key; vals[h] = mrhs; break; }                                             \
              h = (h + 1) & (sz - 1);                                                                                  \
            }                                                                                                          \
          }                                                                                                            \
        }                                                                                                              \
        size_t cnt = 0;                                                                                                \
        for (size_t t = 0; t < sz; ++t) if (keys[t] != -1) { ks[cnt].key = keys[t]; ks[cnt].slot = (int64_t)t; ++cnt; }\
        if (sort) qsort(ks, cnt, sizeof(keyslot), cmp_keyslot);                                                        \
        int64_t o = colptrC[j];                                                                                        \
        for (size_t t = 0; t < cnt; ++t, ++o) { crows[o] = ks[t].key; ccols[o] = j; cvals[o] = vals[ks[t].slot]; }     \
      }                                                                                                                \
      free(keys); free(vals); free(ks);                                                                                \
    }                                                                                                                  \
  }                                                                                                                    \
  /* MultiwayMerge.h:338-422: per column, lists visited in order, add(current, stored) */                              \
  static int64_t merge_##NAME(int k, const port_csc *L, int sort, int count_only, const int64_t *colptrC,              \
                              int64_t *colnnz, int64_t *crows, int64_t *ccols, TC *cvals) {                            \
    int64_t n = L[0].n, total = 0;                                                                                     \
    _Pragma("omp parallel reduction(+ : total)") {                                                                     \
      int64_t *keys = NULL; TC *vals = NULL; keyslot *ks = NULL; size_t cap = 0;                                       \
      _Pragma("omp for schedule(dynamic, 64)")                                                                         \
      for (int64_t j = 0; j < n; ++j) {                                                                                \
        int64_t ub = 0;                                                                                                \
        for (int i = 0; i < k; ++i) ub += L[i].colptr[j + 1] - L[i].colptr[j];                                         \
        if (ub == 0) { if (count_only) colnnz[j] = 0; continue; }                                                      \
        size_t sz = table_size(count_only ? ub : colptrC[j + 1] - colptrC[j]);                                         \
        if (sz > cap) {                                                                                                \
          free(keys); free(vals); free(ks);                                                                            \
          keys = (int64_t *)malloc(sz * sizeof(int64_t)); vals = (TC *)malloc(sz * sizeof(TC));                        \
          ks = (keyslot *)malloc(sz * sizeof(keyslot)); cap = sz;                                                      \
        }                                                                                                              \
        for (size_t t = 0; t < sz; ++t) keys[t] = -1;                                                                  \
        size_t cnt = 0;                                                                                                \
        for (int i = 0; i < k; ++i) {                                                                                  \
          const TC *lv = (const TC *)L[i].vals;                                                                        \
          for (int64_t p = L[i].colptr[j]; p < L[i].colptr[j + 1]; ++p) {                                              \
            int64_t key = L[i].rows[p];                                                                                \
            TC cur = lv[p];                                                                                            \
            size_t h = (size_t)(key * 107) & (sz - 1);                                                                 \
            for (;;) {                                                                                                 \
              if (keys[h] == key) { vals[h] = ADD(cur, vals[h]); break; }                                              \
              if (keys[h] == -1) { keys[h] = key; vals[h] = cur; ++cnt; break; }                                       \
              h = (h + 1) & (sz - 1);                                                                                  \
            }                                                                                                          \
          }                                                                                                            \
        }                                                                                                              \
        total += (int64_t)cnt;                                                                                         \
        if (count_only) { colnnz[j] = (int64_t)cnt; continue; }                                                        \
        size_t c2 = 0;                                                                                                 \
        for (size_t t = 0; t < sz; ++t) if (keys[t] != -1) { ks[c2].key = keys[t]; ks[c2].slot = (int64_t)t; ++c2; }   \
        if (sort) qsort(ks, c2, sizeof(keyslot), cmp_keyslot);                                                         \
        int64_t o = colptrC[j];                                                                                        \
        for (size_t t = 0; t < c2; ++t, ++o) { crows[o] = ks[t].key; ccols[o] = j; cvals[o] = vals[ks[t].slot]; }      \
      }                                                                                                                \
      free(keys); free(vals); free(ks);                                                                                \
    }                                                                                                                  \
    return total;                                                                                                      \
  }

DEFINE_SPGEMM(sr0, double, double, double, f64_mul, f64_add)
DEFINE_SPGEMM(sr1, float, float, float, f32_mul, f32_add)
DEFINE_SPGEMM(sr2, int64_t, int64_t, int64_t, i64_mul, i64_add)
DEFINE_SPGEMM(sr3, uint8_t, int64_t, int64_t, selmax_bool_mul, i64_max)
DEFINE_SPGEMM(sr4, double, double, double, minplus_mul, f64_min)
DEFINE_SPGEMM(sr5, uint8_t, uint8_t, uint8_t, bool_and, bool_or)
DEFINE_SPGEMM(sr6, uint8_t, double, double, boolf64_mul, f64_add)
DEFINE_SPGEMM(sr7, int32_t, int32_t, int32_t, i32_mul, i32_add)
DEFINE_SPGEMM(sr8, int64_t, int64_t, int64_t, i64_mul, i64_max)
DEFINE_SPGEMM(sr9, uint8_t, double, double, copy2nd_f64, forbidden_f64)
DEFINE_SPGEMM(sr10, double, uint8_t, double, copy1st_f64, forbidden_f64)
DEFINE_SPGEMM(sr11, uint8_t, int64_t, int64_t, copy2nd_i64, forbidden_i64)
DEFINE_SPGEMM(sr12, int64_t, uint8_t, int64_t, copy1st_i64, forbidden_i64)
DEFINE_SPGEMM(sr13, uint8_t, uint8_t, uint8_t, copy2nd_u8, forbidden_u8)
DEFINE_SPGEMM(sr14, uint8_t, uint8_t, uint8_t, copy1st_u8, forbidden_u8)

#define PORT_SR_COUNT 15
static const int OUT_BYTES[PORT_SR_COUNT] = {8, 4, 8, 8, 8, 1, 8, 4, 8, 8, 8, 8, 8, 1, 1};

static port_result *alloc_result(int64_t nnz, int vbytes) {
  port_result *r = (port_result *)calloc(1, sizeof(port_result));
  r->nnz = nnz; r->vbytes = vbytes;
  size_t cnt = nnz > 0 ? (size_t)nnz : 1;
  r->rows = (int64_t *)malloc(cnt * sizeof(int64_t));
  r->cols = (int64_t *)malloc(cnt * sizeof(int64_t));
  r->vals = malloc(cnt * (size_t)vbytes);
  return r;
}

/* mtSpGEMM.h:24-70 prefixsum: returns size+1 entries */
static int64_t *exclusive_scan(const int64_t *in, int64_t n) {
  int64_t *out = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
  out[0] = 0;
  for (int64_t i = 0; i < n; ++i) out[i + 1] = out[i] + in[i];
  return out;
}

int port_spgemm(int semiring, const port_csc *A, const port_csc *B, int sort, port_result **out, double *seconds) {
  if (semiring < 0 || semiring >= PORT_SR_COUNT || A->n != B->m) return -1;
  bool_copy_add_happened = 0;
  double t0 = now_s();
  int64_t *flop = (int64_t *)malloc((size_t)(B->n + 1) * sizeof(int64_t));
  int64_t *nnz = (int64_t *)malloc((size_t)(B->n + 1) * sizeof(int64_t));
  port_symbolic(A, B, flop, nnz);
  int64_t *colptrC = exclusive_scan(nnz, B->n);
  port_result *r = alloc_result(colptrC[B->n], OUT_BYTES[semiring]);
  switch (semiring) {
    case 0: spgemm_sr0(A, B, sort, colptrC, r->rows, r->cols, (double *)r->vals); break;
    case 1: spgemm_sr1(A, B, sort, colptrC, r->rows, r->cols, (float *)r->vals); break;
    case 2: spgemm_sr2(A, B, sort, colptrC, r->rows, r->cols, (int64_t *)r->vals); break;
    case 3: spgemm_sr3(A, B, sort, colptrC, r->rows, r->cols, (int64_t *)r->vals); break;
    case 4: spgemm_sr4(A, B, sort, colptrC, r->rows, r->cols, (double *)r->vals); break;
    case 5: spgemm_sr5(A, B, sort, colptrC, r->rows, r->cols, (uint8_t *)r->vals); break;
    case 6: spgemm_sr6(A, B, sort, colptrC, r->rows, r->cols, (double *)r->vals); break;
    case 7: spgemm_sr7(A, B, sort, colptrC, r->rows, r->cols, (int32_t *)r->vals); break;
    case 8: spgemm_sr8(A, B, sort, colptrC, r->rows, r->cols, (int64_t *)r->vals); break;
    case 9: spgemm_sr9(A, B, sort, colptrC, r->rows, r->cols, (double *)r->vals); break;
    case 10: spgemm_sr10(A, B, sort, colptrC, r->rows, r->cols, (double *)r->vals); break;
    case 11: spgemm_sr11(A, B, sort, colptrC, r->rows, r->cols, (int64_t *)r->vals); break;
    case 12: spgemm_sr12(A, B, sort, colptrC, r->rows, r->cols, (int64_t *)r->vals); break;
    case 13: spgemm_sr13(A, B, sort, colptrC, r->rows, r->cols, (uint8_t *)r->vals); break;
    case 14: spgemm_sr14(A, B, sort, colptrC, r->rows, r->cols, (uint8_t *)r->vals); break;
  }
  free(flop); free(nnz); free(colptrC);
  if (seconds) *seconds = now_s() - t0;
  if (bool_copy_add_happened) { port_result_free(r); return -2; } /* "Add should not happen" */
  *out = r;
  return 0;
}

#define MERGE_CALL(NAME, T, count_only, cp, cn, r)                                                                     \
  merge_##NAME(k, lists, sort, count_only, cp, cn, (r) ? (r)->rows : NULL, (r) ? (r)->cols : NULL,                     \
               (r) ? (T *)(r)->vals : NULL)

static int64_t merge_dispatch(int semiring, int k, const port_csc *lists, int sort, int count_only, const int64_t *cp,
                              int64_t *cn, port_result *r) {
  switch (semiring) {
    case 0: return MERGE_CALL(sr0, double, count_only, cp, cn, r);
    case 1: return MERGE_CALL(sr1, float, count_only, cp, cn, r);
    case 2: return MERGE_CALL(sr2, int64_t, count_only, cp, cn, r);
    case 3: return MERGE_CALL(sr3, int64_t, count_only, cp, cn, r);
    case 4: return MERGE_CALL(sr4, double, count_only, cp, cn, r);
    case 5: return MERGE_CALL(sr5, uint8_t, count_only, cp, cn, r);
    case 6: return MERGE_CALL(sr6, double, count_only, cp, cn, r);
    case 7: return MERGE_CALL(sr7, int32_t, count_only, cp, cn, r);
    case 8: return MERGE_CALL(sr8, int64_t, count_only, cp, cn, r);
    case 9: return MERGE_CALL(sr9, double, count_only, cp, cn, r);
    case 10: return MERGE_CALL(sr10, double, count_only, cp, cn, r);
    case 11: return MERGE_CALL(sr11, int64_t, count_only, cp, cn, r);
    case 12: return MERGE_CALL(sr12, int64_t, count_only, cp, cn, r);
    case 13: return MERGE_CALL(sr13, uint8_t, count_only, cp, cn, r);
    case 14: return MERGE_CALL(sr14, uint8_t, count_only, cp, cn, r);
  }
  return -1;
}

int port_merge(int semiring, int k, const port_csc *lists, int sort, port_result **out, double *seconds) {
  if (semiring < 0 || semiring >= PORT_SR_COUNT || k < 1) return -1;
  bool_copy_add_happened = 0;
  double t0 = now_s();
  int64_t n = lists[0].n;
  int64_t *colnnz = (int64_t *)malloc((size_t)(n + 1) * sizeof(int64_t));
  merge_dispatch(semiring, k, lists, sort, 1, NULL, colnnz, NULL); /* SerialMergeNNZHash, MultiwayMerge.h:255-329 */
  int64_t *cp = exclusive_scan(colnnz, n);
  port_result *r = alloc_result(cp[n], OUT_BYTES[semiring]);
  merge_dispatch(semiring, k, lists, sort, 0, cp, NULL, r);
  free(colnnz); free(cp);
  if (seconds) *seconds = now_s() - t0;
  if (bool_copy_add_happened) { port_result_free(r); return -2; } /* "Add should not happen" */
  *out = r;
  return 0;
}

int64_t port_result_nnz(const port_result *r) { return r->nnz; }
void port_result_copy(const port_result *r, int64_t *rows, int64_t *cols, void *vals) {
  if (r->nnz <= 0) return;
  memcpy(rows, r->rows, (size_t)r->nnz * sizeof(int64_t));
  memcpy(cols, r->cols, (size_t)r->nnz * sizeof(int64_t));
  memcpy(vals, r->vals, (size_t)r->nnz * (size_t)r->vbytes);
}
void port_result_free(port_result *r) {
  if (!r) return;
  free(r->rows); free(r->cols); free(r->vals); free(r);
}
int port_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void port_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------------------------
 * Seeded R-MAT edge stream of the benchmark inputs, on the host cores (threaded). Same integer arithmetic as the
 * library's device generator (combblas_b200/csrc/gen.cu: splitmix64 per level, thresholds scaled to 2^53, seeded
 * bijective vertex scramble); tests assert that both produce identical edges. The checkers and the reference arm of
 * bench.py build their matrices from this copy, so that they never touch the product library.
 * The reference's counterpart is DistEdgeList::GenGraph500Data (DistEdgeList.cpp:223-279). */
static inline uint64_t gen_splitmix64(uint64_t *s) {
  *s += 0x9E3779B97F4A7C15ULL;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

int port_rmat_edges(int scale, int64_t nedges, uint64_t seed, double a, double b, double c, int scramble, int64_t *rows,
                    int64_t *cols) {
  if (scale < 1 || scale > 31 || nedges < 0 || !rows || !cols) return -1;
  const double two53 = 9007199254740992.0;
  const uint64_t ta = (uint64_t)(a * two53), tab = (uint64_t)((a + b) * two53), tabc = (uint64_t)((a + b + c) * two53);
  const uint64_t mask = (1ULL << scale) - 1;
  uint64_t ss = seed ^ 0xD1B54A32D192ED03ULL;
  const uint64_t m1 = gen_splitmix64(&ss) | 1ULL, c1 = gen_splitmix64(&ss), m2 = gen_splitmix64(&ss) | 1ULL;
  const int sh = scale / 2 > 0 ? scale / 2 : 1;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nedges; ++e) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ULL + (uint64_t)e * 0xD6E8FEB86659FD93ULL + 0x2545F4914F6CDD1DULL;
    uint64_t r = 0, cc = 0;
    for (int l = 0; l < scale; ++l) {
      const uint64_t u = gen_splitmix64(&s) >> 11;
      int rb, cb;
      if (u < ta) { rb = 0; cb = 0; }
      else if (u < tab) { rb = 0; cb = 1; }
      else if (u < tabc) { rb = 1; cb = 0; }
      else { rb = 1; cb = 1; }
      r = (r << 1) | (uint64_t)rb;
      cc = (cc << 1) | (uint64_t)cb;
    }
    if (scramble) {
      r = (r * m1 + c1) & mask; r ^= r >> sh; r = (r * m2) & mask;
      cc = (cc * m1 + c1) & mask; cc ^= cc >> sh; cc = (cc * m2) & mask;
    }
    rows[e] = (int64_t)r;
    cols[e] = (int64_t)cc;
  }
  return 0;
}
