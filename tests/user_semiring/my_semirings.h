// Two semirings a driver defines itself, written the way the reference's drivers write them
// (ReleaseTests/KTipsTest.cpp:12-20 for OR-AND; Semirings.h:143-255 for the member list), with CBGPU_HD on the members so that
// the same struct is host code for the reference's templates and device code for the accumulation engine.
#pragma once
#include <mpi.h>
#include <stdint.h>
#include "combblas_b200/semiring_decl.h"

// KTips-style boolean semiring: add = OR, multiply = AND
struct KTipsOrAnd {
  static CBGPU_HD bool id() { return false; }
  static bool returnedSAID() { return false; }
  static MPI_Op mpi_op() { return MPI_LOR; }
  static CBGPU_HD bool add(const bool &a, const bool &b) { return a || b; }
  static CBGPU_HD bool multiply(const bool &a, const bool &b) { return a && b; }
  static void axpy(bool a, const bool &x, bool &y) { y = y || (a && x); }
};

// (max, x) on non-negative doubles: not in the library's list, 8-byte accumulators with a user-defined add
struct MaxTimesF64 {
  static CBGPU_HD double id() { return 0.0; }
  static bool returnedSAID() { return false; }
  static MPI_Op mpi_op() { return MPI_MAX; }
  static CBGPU_HD double add(const double &a, const double &b) { return a < b ? b : a; }
  static CBGPU_HD double multiply(const double &a, const double &b) { return a * b; }
  static void axpy(double a, const double &x, double &y) { y = add(y, multiply(a, x)); }
};

// (min, +) on int32 with a saturating "infinity": 4-byte accumulators
struct MinPlusI32 {
  static CBGPU_HD int32_t id() { return INT32_MAX; }
  static bool returnedSAID() { return false; }
  static MPI_Op mpi_op() { return MPI_MIN; }
  static CBGPU_HD int32_t add(const int32_t &a, const int32_t &b) { return a < b ? a : b; }
  static CBGPU_HD int32_t multiply(const int32_t &a, const int32_t &b) {
    return (a == INT32_MAX || b == INT32_MAX) ? INT32_MAX : a + b;
  }
  static void axpy(int32_t a, const int32_t &x, int32_t &y) { y = add(y, multiply(a, x)); }
};
