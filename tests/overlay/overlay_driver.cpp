// A reference-style driver compiled UNCHANGED against the overlay: it only includes "CombBLAS/CombBLAS.h" and calls the
// reference API (LocalHybridSpGEMM, PSpGEMM -> Mult_AnXBn_Synch, MultiwayMerge). With the overlay first on the include
// path those calls reach libcbgpu.so; a structurally identical user semiring that the overlay does not know keeps the
// reference's CPU path, which serves as the in-process comparison.
#include <mpi.h>
#include <cstdio>
#include <random>
#include "CombBLAS/CombBLAS.h"

using namespace combblas;
int cblas_splits = 1;

// same arithmetic as PlusTimesSRing<double,double>, unknown to the overlay -> stays on the reference path
struct MyPlusTimes {
  static double id() { return 0; }
  static bool returnedSAID() { return false; }
  static MPI_Op mpi_op() { return MPI_SUM; }
  static double add(const double &a, const double &b) { return a + b; }
  static double multiply(const double &a, const double &b) { return a * b; }
  static void axpy(double a, const double &x, double &y) { y += a * x; }
};

typedef SpDCCols<int64_t, double> DCD;
typedef PlusTimesSRing<double, double> PTDD;

static DCD *random_block(int64_t m, int64_t n, int64_t nnz, unsigned seed) {
  std::mt19937_64 g(seed);
  std::tuple<int64_t, int64_t, double> *t = new std::tuple<int64_t, int64_t, double>[nnz];
  for (int64_t i = 0; i < nnz; ++i) t[i] = std::make_tuple((int64_t)(g() % m), (int64_t)(g() % n), 1.0 + (double)(g() % 7));
  SpTuples<int64_t, double> tup(nnz, m, n, t, false, false);
  tup.RemoveDuplicates(std::plus<double>());
  return new DCD(tup, false);
}

static bool same(SpTuples<int64_t, double> &x, SpTuples<int64_t, double> &y) {
  x.SortColBased();
  y.SortColBased();
  if (x.getnnz() != y.getnnz()) return false;
  for (int64_t i = 0; i < x.getnnz(); ++i) {
    if (x.rowindex(i) != y.rowindex(i) || x.colindex(i) != y.colindex(i)) return false;
    double a = x.numvalue(i), b = y.numvalue(i);
    if (std::abs(a - b) > 1e-12 * std::max(std::abs(a), std::abs(b))) return false;
  }
  return true;
}

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  int fails = 0;
  {
    DCD *A = random_block(3000, 2500, 40000, 1), *B = random_block(2500, 2800, 35000, 2);
    SpTuples<int64_t, double> *gpu = LocalHybridSpGEMM<PTDD, double>(*A, *B, false, false);     // -> libcbgpu.so
    SpTuples<int64_t, double> *cpu = LocalHybridSpGEMM<MyPlusTimes, double>(*A, *B, false, false); // -> reference
    bool ok = same(*gpu, *cpu);
    std::printf("%s LocalHybridSpGEMM overlay vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)gpu->getnnz());
    fails += !ok;
    std::vector<SpTuples<int64_t, double> *> lists;
    for (int i = 0; i < 3; ++i) {
      DCD *Ai = random_block(3000, 2500, 30000, 10 + i);
      lists.push_back(LocalSpGEMMHash<PTDD, double>(*Ai, *B, false, false, false));
      delete Ai;
    }
    SpTuples<int64_t, double> *mg = MultiwayMerge<PTDD>(lists, (int64_t)3000, (int64_t)2800, false); // -> libcbgpu.so
    SpTuples<int64_t, double> *mc = MultiwayMerge<MyPlusTimes>(lists, (int64_t)3000, (int64_t)2800, false); // -> reference
    ok = same(*mg, *mc);
    std::printf("%s MultiwayMerge overlay vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)mg->getnnz());
    fails += !ok;
    for (auto p : lists) delete p;
    delete gpu; delete cpu; delete mg; delete mc; delete A; delete B;
  }
  {
    // the distributed driver, unchanged: PSpGEMM -> Mult_AnXBn_Synch (P = 1) -> our LocalHybridSpGEMM / MultiwayMerge
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    SpParMat<int64_t, double, DCD> A(random_block(4000, 4000, 60000, 5), grid), B(random_block(4000, 4000, 60000, 6), grid);
    SpParMat<int64_t, double, DCD> Cg = PSpGEMM<PTDD>(A, B);
    SpParMat<int64_t, double, DCD> Cc = PSpGEMM<MyPlusTimes>(A, B);
    bool ok = (Cg == Cc); // the reference's own equality (Dcsc::operator==, dcsc.cpp:522-568)
    std::printf("%s PSpGEMM (Mult_AnXBn_Synch) overlay vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)Cg.getnnz());
    fails += !ok;
  }
  MPI_Finalize();
  return fails;
}
