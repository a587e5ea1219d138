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

// the indexing semirings under names the overlay does not know -> SubsRef_SR stays on the reference path with them
struct MyCopy1st : BoolCopy1stSRing<double> {};
struct MyCopy2nd : BoolCopy2ndSRing<double> {};

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

// where two distributed results differ (printed on failure only): first entries of the local blocks that do not match
static void diff_report(SpParMat<int64_t, double, DCD> &X, SpParMat<int64_t, double, DCD> &Y) {
  SpTuples<int64_t, double> x(X.seq()), y(Y.seq());
  x.SortColBased();
  y.SortColBased();
  std::printf("  local nnz %lld vs %lld, shape %lldx%lld vs %lldx%lld, nzc %lld vs %lld\n", (long long)x.getnnz(), (long long)y.getnnz(),
              (long long)X.seq().getnrow(), (long long)X.seq().getncol(), (long long)Y.seq().getnrow(), (long long)Y.seq().getncol(),
              (long long)X.seq().getnzc(), (long long)Y.seq().getnzc());
  if (X.seq().getnzc() == Y.seq().getnzc() && X.seq().getnnz() > 0 && Y.seq().getnnz() > 0) {
    Dcsc<int64_t, double> *a = X.seq().GetDCSC(), *b = Y.seq().GetDCSC();
    for (int64_t c = 0; c < a->nzc; ++c)
      if (a->jc[c] != b->jc[c] || a->cp[c + 1] != b->cp[c + 1]) {
        std::printf("  column slot %lld: jc %lld vs %lld, cp %lld vs %lld\n", (long long)c, (long long)a->jc[c], (long long)b->jc[c],
                    (long long)a->cp[c + 1], (long long)b->cp[c + 1]);
        break;
      }
    for (int64_t p = 0; p < a->nz; ++p)
      if (a->ir[p] != b->ir[p] || a->numx[p] != b->numx[p]) {
        std::printf("  position %lld: row %lld vs %lld, value %.17g vs %.17g\n", (long long)p, (long long)a->ir[p], (long long)b->ir[p],
                    a->numx[p], b->numx[p]);
        break;
      }
  }
  int shown = 0;
  int64_t i = 0, j = 0;
  while ((i < x.getnnz() || j < y.getnnz()) && shown < 12) {
    const bool hx = i < x.getnnz(), hy = j < y.getnnz();
    const int64_t cx = hx ? x.colindex(i) : INT64_MAX, cy = hy ? y.colindex(j) : INT64_MAX;
    const int64_t rx = hx ? x.rowindex(i) : INT64_MAX, ry = hy ? y.rowindex(j) : INT64_MAX;
    if (cx == cy && rx == ry) {
      if (x.numvalue(i) != y.numvalue(j)) {
        std::printf("  (%lld,%lld): %.17g vs %.17g\n", (long long)rx, (long long)cx, x.numvalue(i), y.numvalue(j));
        ++shown;
      }
      ++i, ++j;
    } else if (cx < cy || (cx == cy && rx < ry)) {
      std::printf("  (%lld,%lld) = %.17g only in the first\n", (long long)rx, (long long)cx, x.numvalue(i));
      ++shown, ++i;
    } else {
      std::printf("  (%lld,%lld) = %.17g only in the second\n", (long long)ry, (long long)cy, y.numvalue(j));
      ++shown, ++j;
    }
  }
}

// equality as matrices. MemEfficientSpGEMM multiplies with LocalSpGEMMHash(sort = false) (ParFriends.h:659), so the reference's
// block keeps its hash-table row order inside every column and Dcsc::operator== (an array comparison, dcsc.cpp:522-568)
// cannot be used against the device result, whose rows are ascending: compare the column-sorted tuples instead.
static bool same_matrix(SpParMat<int64_t, double, DCD> &X, SpParMat<int64_t, double, DCD> &Y) {
  SpTuples<int64_t, double> x(X.seq()), y(Y.seq());
  return X.seq().getnrow() == Y.seq().getnrow() && X.seq().getncol() == Y.seq().getncol() && same(x, y);
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
    // the distributed driver, unchanged: PSpGEMM -> our Mult_AnXBn_Synch overload (blocks resident, cbgpu_summa2d over NCCL)
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    SpParMat<int64_t, double, DCD> A(random_block(4000, 4000, 60000, 5), grid), B(random_block(4000, 4000, 60000, 6), grid);
    SpParMat<int64_t, double, DCD> Cg = PSpGEMM<PTDD>(A, B);
    SpParMat<int64_t, double, DCD> Cc = PSpGEMM<MyPlusTimes>(A, B);
    bool ok = (Cg == Cc); // the reference's own equality (Dcsc::operator==, dcsc.cpp:522-568)
    std::printf("%s PSpGEMM (Mult_AnXBn_Synch) overlay vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)Cg.getnnz());
    fails += !ok;
  }
  {
    // HipMCL's expansion, unchanged: MemEfficientSpGEMM (phases + MCLPruneRecoverySelect per phase) -> one device call
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    SpParMat<int64_t, double, DCD> A(random_block(3000, 3000, 90000, 7), grid), B(random_block(3000, 3000, 90000, 8), grid);
    SpParMat<int64_t, double, DCD> Cg = MemEfficientSpGEMM<PTDD, double, DCD>(A, B, 3, 2.0, (int64_t)20, (int64_t)30, 0.9, 1, 1, (int64_t)0);
    SpParMat<int64_t, double, DCD> Cc = MemEfficientSpGEMM<MyPlusTimes, double, DCD>(A, B, 3, 2.0, (int64_t)20, (int64_t)30, 0.9, 1, 1, (int64_t)0);
    bool ok = same_matrix(Cg, Cc) && Cg.getnnz() > 0;
    std::printf("%s MemEfficientSpGEMM overlay vs reference: nnz %lld (unpruned product would be larger)\n", ok ? "PASS" : "FAIL", (long long)Cg.getnnz());
    if (!ok) diff_report(Cg, Cc);
    fails += !ok;
  }
  {
    // the 3D drivers, unchanged: SpParMat3D operands, Mult_AnXBn_SUMMA3D and MemEfficientSpGEMM3D, compared after Convert2D
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    SpParMat<int64_t, double, DCD> A(random_block(3000, 3000, 80000, 9), grid), B(random_block(3000, 3000, 80000, 10), grid);
    SpParMat3D<int64_t, double, DCD> A3(A, 1, true, false), B3(B, 1, false, false);
    SpParMat3D<int64_t, double, DCD> Cg3 = Mult_AnXBn_SUMMA3D<PTDD, double, DCD>(A3, B3);
    SpParMat3D<int64_t, double, DCD> Cc3 = Mult_AnXBn_SUMMA3D<MyPlusTimes, double, DCD>(A3, B3);
    SpParMat<int64_t, double, DCD> Cg = Cg3.Convert2D(), Cc = Cc3.Convert2D();
    bool ok = (Cg == Cc) && Cg.getnnz() > 0;
    std::printf("%s Mult_AnXBn_SUMMA3D overlay vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)Cg.getnnz());
    fails += !ok;
    SpParMat3D<int64_t, double, DCD> Pg3 = MemEfficientSpGEMM3D<PTDD, double, DCD>(A3, B3, 2, 2.0, (int64_t)20, (int64_t)30, 0.9, 1, 1, (int64_t)0);
    SpParMat3D<int64_t, double, DCD> Pc3 = MemEfficientSpGEMM3D<MyPlusTimes, double, DCD>(A3, B3, 2, 2.0, (int64_t)20, (int64_t)30, 0.9, 1, 1, (int64_t)0);
    SpParMat<int64_t, double, DCD> Pg = Pg3.Convert2D(), Pc = Pc3.Convert2D();
    ok = same_matrix(Pg, Pc) && Pg.getnnz() > 0;
    std::printf("%s MemEfficientSpGEMM3D overlay vs reference: nnz %lld\n", ok ? "PASS" : "FAIL", (long long)Pg.getnnz());
    if (!ok) diff_report(Pg, Pc);
    fails += !ok;
  }
  {
    // indexing, unchanged: A(ri, ci) -> SubsRef_SR<BoolCopy1stSRing, BoolCopy2ndSRing> -> two Mult_AnXBn_DoubleBuff calls with
    // boolean selection matrices (SpParMat.cpp:2515-2566): the overlay's DoubleBuff overload + the BoolCopy device semirings
    std::shared_ptr<CommGrid> grid(new CommGrid(MPI_COMM_WORLD, 0, 0));
    SpParMat<int64_t, double, DCD> A(random_block(3000, 3000, 70000, 11), grid);
    FullyDistVec<int64_t, int64_t> ri(grid, 700, 0), ci(grid, 600, 0);
    std::mt19937_64 g(12);
    for (int64_t i = 0; i < 700; ++i) ri.SetElement(i, (int64_t)(g() % 3000)); // repeated indices allowed
    for (int64_t i = 0; i < 600; ++i) ci.SetElement(i, (int64_t)(g() % 3000));
    SpParMat<int64_t, double, DCD> Bg = A(ri, ci);                                           // -> libcbgpu.so
    SpParMat<int64_t, double, DCD> Bc = A.SubsRef_SR<MyCopy1st, MyCopy2nd>(ri, ci, false);   // -> reference
    bool ok = same_matrix(Bg, Bc) && Bg.getnnz() > 0 && Bg.getnrow() == 700 && Bg.getncol() == 600;
    std::printf("%s SpParMat::operator()(ri, ci) (SubsRef_SR, BoolCopy semirings through Mult_AnXBn_DoubleBuff) overlay vs reference: nnz %lld\n",
                ok ? "PASS" : "FAIL", (long long)Bg.getnnz());
    if (!ok) diff_report(Bg, Bc);
    fails += !ok;
  }
  MPI_Finalize();
  return fails;
}
