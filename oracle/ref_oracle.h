/*
 * C interface of the parity oracle built from the UNMODIFIED CombBLAS reference sources
 * (read in place from /root/reference; see oracle/Makefile). TEST INFRASTRUCTURE ONLY:
 * nothing in the product (libcbgpu.so, combblas_b200/) links or loads this.
 *
 * Matrices cross this interface as plain CSC (colptr over all n columns). Values are
 * typed by the semiring id (see cbgpu_semiring in include/cbgpu.h — same numbering).
 */
#ifndef CBGPU_REF_ORACLE_H
#define CBGPU_REF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int64_t m, n, nnz;
  const int64_t *colptr; /* n+1 */
  const int64_t *rows;   /* nnz */
  const void *vals;      /* nnz elements of the semiring's operand type */
} ref_csc;

typedef struct ref_result ref_result;

/* which reference routine to run */
enum {
  REF_LOCAL_HYBRID = 0,        /* LocalHybridSpGEMM      mtSpGEMM.h:214 */
  REF_LOCAL_HASH_SORTED = 1,   /* LocalSpGEMMHash(sort=true)  mtSpGEMM.h:464 */
  REF_LOCAL_HASH_UNSORTED = 2, /* LocalSpGEMMHash(sort=false) */
  REF_LOCAL_HEAP = 3,          /* LocalSpGEMM            mtSpGEMM.h:75 */
  REF_DIST_SYNCH = 10,         /* Mult_AnXBn_Synch at P=1      ParFriends.h:1448 */
  REF_DIST_DOUBLEBUFF = 11,    /* Mult_AnXBn_DoubleBuff at P=1 ParFriends.h:1239 */
  REF_DIST_MEMEFF_HASH = 12,   /* MemEfficientSpGEMM(kernel=1, no pruning) ParFriends.h:453 */
  REF_DIST_MEMEFF_HEAP = 13,   /* MemEfficientSpGEMM(kernel=2, no pruning) */
  REF_DIST_SUMMA3D = 14        /* Mult_AnXBn_SUMMA3D layers=1 + Convert2D  ParFriends.h:3375 */
};

/* C = A (x) B over `semiring`. If canonical != 0 the result is re-sorted column-major with rows
 * ascending before being handed back. `phases` is used by the MemEfficient variants only.
 * seconds (may be NULL) receives the wall time of the multiply call alone. Returns 0 on success. */
int ref_spgemm(int routine, int semiring, const ref_csc *A, const ref_csc *B, int phases, int canonical,
               ref_result **out, double *seconds);

/* k-way merge of column-sorted lists with SR::add. hash=0: MultiwayMerge (MultiwayMerge.h:429),
 * hash=1: MultiwayMergeHash(sorted) (MultiwayMerge.h:554). Values are the semiring's OUTPUT type. */
int ref_merge(int hash, int semiring, int k, const ref_csc *lists, int sorted, int canonical, ref_result **out,
              double *seconds);

/* reference symbolic pass: per-non-empty-B-column flop (estimateFLOP, mtSpGEMM.h:1058) and nnz
 * (estimateNNZ_Hash, mtSpGEMM.h:807); arrays sized nzc(B), *nzc receives the count. Caller frees with ref_free. */
int ref_symbolic(int semiring, const ref_csc *A, const ref_csc *B, int64_t *nzc, int64_t **flop, int64_t **nnz);

/* MCLPruneRecoverySelect (ParFriends.h:186-354) applied to A as a P=1 SpParMat; semiring 0 (double) or 1 (float) names
 * the value type. Result canonical (column-major, rows ascending). */
int ref_mcl_prune(int semiring, const ref_csc *A, double hard, int64_t select, int64_t recover, double pct,
                  int kselect_version, ref_result **out);
/* MemEfficientSpGEMM (ParFriends.h:453-777) at P=1 WITH its pruning parameters; kernel 1 = hash, 2 = heap */
int ref_memeff_prune(int semiring, const ref_csc *A, const ref_csc *B, int phases, double hard, int64_t select,
                     int64_t recover, double pct, int kselect_version, int kernel, ref_result **out);
void ref_free(void *p);

int64_t ref_result_nnz(const ref_result *r);
int ref_result_value_bytes(const ref_result *r);
/* copies COO in the order the reference produced it (or canonical order if requested above) */
void ref_result_copy(const ref_result *r, int64_t *rows, int64_t *cols, void *vals);
void ref_result_free(ref_result *r);

int ref_num_threads(void);
void ref_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
