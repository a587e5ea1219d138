// A reference-style driver with its OWN semiring structs (the KTipsTest.cpp:12-20 situation): the structs live in a header of
// the application with CBGPU_HD on their members, the application's .cu instantiates them for the device
// (tests/user_semiring/my_semirings.cu), and after CBGPU_DECLARE_SEMIRING the unchanged calls PSpGEMM<KTipsOrAnd>(A, B),
// LocalHybridSpGEMM<MaxTimesF64, double>(...) take the device path. Structs with the same arithmetic that are NOT declared
// keep the reference's CPU templates and serve as the in-process comparison.
#include <mpi.h>
#include <cstdio>
#include <random>
#include "CombBLAS/CombBLAS.h"
#include "../user_semiring/my_semirings.h"

CBGPU_DECLARE_SEMIRING(ktips_or_and_id, KTipsOrAnd, bool, bool, bool)
CBGPU_DECLARE_SEMIRING(max_times_f64_id, MaxTimesF64, double, double, double)

using namespace combblas;
int cblas_splits = 1;

struct CpuOrAnd : KTipsOrAnd {};     // same members, unknown to the overlay -> reference path
struct CpuMaxTimes : MaxTimesF64 {};

typedef SpDCCols<int64_t, bool> DCB;
typedef SpDCCols<int64_t, double> DCD;

template <class NT>
static SpDCCols<int64_t, NT> *random_block(int64_t m, int64_t n, int64_t nnz, unsigned seed) {
  std::mt19937_64 g(seed);
  std::tuple<int64_t, int64_t, NT> *t = new std::tuple<int64_t, int64_t, NT>[nnz];
  for (int64_t i = 0; i < nnz; ++i) t[i] = std::make_tuple((int64_t)(g() % m), (int64_t)(g() % n), (NT)(1 + g() % 7));
  SpTuples<int64_t, NT> tup(nnz, m, n, t, false, false);
  tup.RemoveDuplicates([](NT a, NT b) { return a < b ? b : a; });
  return new SpDCCols<int64_t, NT>(tup, false);
}

template <class NT>
static bool same(SpTuples<int64_t, NT> &x, SpTuples<int64_t, NT> &y) {
  x.SortColBased();
  y.SortColBased();
  if (x.getnnz() != y.getnnz()) return false;
  for (int64_t i = 0; i < x.getnnz(); ++i)
    if (x.rowindex(i) != y.rowindex(i) || x.colindex(i) != y.colindex(i) || x.numvalue(i) != y.numvalue(i)) return false;
  return true;
}

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  int fails = 0;
  static_assert(cbgpu_overlay::supported<KTipsOrAnd, int64_t, bool, bool, bool>, "declared user semiring takes the device path");
  static_assert(!cbgpu_overlay::supported<CpuOrAnd, int64_t, bool, bool, bool>, "undeclared struct stays on the reference path");
  static_assert(!cbgpu_overlay::supported<KTipsOrAnd, int64_t, double, bool, bool>, "operand types must be the declared ones");
  {
    DCD *A = random_block<double>(3000, 2500, 40000, 1), *B = random_block<double>(2500, 2800, 35000, 2);
    SpTuples<int64_t, double> *gpu = LocalHybridSpGEMM<MaxTimesF64, double>(*A, *B, false, false); // -> device, user add = max
    SpTuples<int64_t, double> *cpu = LocalHybridSpGEMM<CpuMaxTimes, double>(*A, *B, false, false); // -> reference
    bool ok = same(*gpu, *cpu);
    std::printf("%s LocalHybridSpGEMM<MaxTimesF64> device vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)gpu->getnnz());
    fails += !ok;
    delete gpu; delete cpu; delete A; delete B;
  }
  {
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    SpParMat<int64_t, bool, DCB> A(random_block<bool>(4000, 4000, 60000, 5), grid), B(random_block<bool>(4000, 4000, 60000, 6), grid);
    SpParMat<int64_t, bool, DCB> Cg = PSpGEMM<KTipsOrAnd>(A, B); // -> Mult_AnXBn_Synch -> device
    SpParMat<int64_t, bool, DCB> Cc = PSpGEMM<CpuOrAnd>(A, B);   // -> reference
    bool ok = (Cg == Cc);
    std::printf("%s PSpGEMM<KTipsOrAnd> device vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)Cg.getnnz());
    fails += !ok;
  }
  MPI_Finalize();
  return fails;
}
