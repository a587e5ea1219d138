/*
 * Plain-C restatement ("port") of the reference's local semiring SpGEMM and k-way merge.
 * TEST INFRASTRUCTURE ONLY -- the checker for tests/, smoke() and bench.py's cpu_baseline leg.
 * Never linked into or called from the product (libcbgpu.so / combblas_b200).
 *
 * Parity pinning: validated against (a) the unmodified reference built in oracle/_ref (tests/test_oracle.py,
 * run in the build container) and (b) the reference tree's own known-answer vector bcsstk01^2 == C.mtx
 * (3DSpGEMM/matlab), committed as tests/golden/bcsstk01_*.npz by tests/golden/make_golden.py.
 */
#ifndef CBGPU_SPGEMM_ORACLE_H
#define CBGPU_SPGEMM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int64_t m, n, nnz;
  const int64_t *colptr; /* n+1 */
  const int64_t *rows;   /* nnz; ascending within a column for operands */
  const void *vals;      /* nnz */
} port_csc;

typedef struct port_result port_result;

/* C = A (x) B with the reference's hash algorithm (mtSpGEMM.h:464-656 LocalSpGEMMHash; identical arithmetic
 * and accumulation order to the hash branch of LocalHybridSpGEMM, mtSpGEMM.h:362-440).
 * sort != 0: rows ascending inside each column; sort == 0: the reference's hash-table order. */
int port_spgemm(int semiring, const port_csc *A, const port_csc *B, int sort, port_result **out, double *seconds);

/* k-way merge, per column hash accumulation in list order (MultiwayMergeHash, MultiwayMerge.h:554-701;
 * SerialMergeHash :338-422). Values are of the semiring's output type. */
int port_merge(int semiring, int k, const port_csc *lists, int sort, port_result **out, double *seconds);

/* symbolic pass over ALL n columns of B: flop[j] (estimateFLOP, mtSpGEMM.h:1058-1134) and nnz[j]
 * (estimateNNZ_Hash, mtSpGEMM.h:807-933). Returns total flop. */
int64_t port_symbolic(const port_csc *A, const port_csc *B, int64_t *flop, int64_t *nnz);

int64_t port_result_nnz(const port_result *r);
void port_result_copy(const port_result *r, int64_t *rows, int64_t *cols, void *vals);
void port_result_free(port_result *r);
int port_num_threads(void);
/* seeded R-MAT edge stream of the benchmark inputs (same arithmetic as the library's generator), threaded */
int port_rmat_edges(int scale, int64_t nedges, uint64_t seed, double a, double b, double c, int scramble, int64_t *rows,
                    int64_t *cols);
void port_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
