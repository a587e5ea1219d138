// User-defined semirings on the device path, host-compiler half (plain C++, no CUDA needed).
//
// The reference takes a semiring as a struct with static members (Semirings.h:143-255; ReleaseTests/KTipsTest.cpp:12-20):
//     struct OrAnd { static bool id(); static bool add(const bool&, const bool&); static bool multiply(const bool&, const bool&); ... };
// To run such a struct on the GPU its members must also be device functions. Mark them CBGPU_HD (expands to
// __host__ __device__ under nvcc and to nothing under a host compiler), keep the struct in a header of the application, and
//   1. in ONE .cu file of the application (compiled with nvcc -gencode arch=compute_100a,code=sm_100a, linked to libcbgpu.so):
//          #include "combblas_b200/device_semiring.cuh"
//          #include "my_semiring.h"
//          CBGPU_DEFINE_SEMIRING(my_or_and_id, OrAnd, bool, bool, bool)
//      which instantiates the accumulation engine for the struct and exports   extern "C" int my_or_and_id(void);
//   2. in the driver (host compiler, with the overlay of CombBLAS.h on the include path):
//          CBGPU_DECLARE_SEMIRING(my_or_and_id, OrAnd, bool, bool, bool)
//      after which PSpGEMM<OrAnd>(A, B), LocalHybridSpGEMM<OrAnd, bool>(...), MultiwayMerge<OrAnd>(...) take the device path.
// Value types: double, float, int64_t, int32_t, bool. SR::add must be associative and commutative and SR::id() its
// identity -- what the reference's own kernels assume when they pick heap or hash accumulation per column
// (mtSpGEMM.h:362-440 adds in B-column order, the heap kernel :74-202 in row order).
#ifndef CBGPU_SEMIRING_DECL_H
#define CBGPU_SEMIRING_DECL_H
#include <stdint.h>
#include "../cbgpu.h"

#if defined(__CUDACC__)
#define CBGPU_HD __host__ __device__
#else
#define CBGPU_HD
#endif

namespace cbgpu_overlay {

template <class T> struct dtype_of { static constexpr int value = -1; };
template <> struct dtype_of<double> { static constexpr int value = CBGPU_F64; };
template <> struct dtype_of<float> { static constexpr int value = CBGPU_F32; };
template <> struct dtype_of<int64_t> { static constexpr int value = CBGPU_I64; };
template <> struct dtype_of<int32_t> { static constexpr int value = CBGPU_I32; };
template <> struct dtype_of<bool> { static constexpr int value = CBGPU_BOOL; };

// a semiring struct that has a device instantiation: operand / result value types and the run-time id
template <class SR> struct user_semiring {
  static constexpr bool value = false;
  static constexpr int a_dtype = -1, b_dtype = -1, c_dtype = -1;
  static int id() { return -1; }
};

} // namespace cbgpu_overlay

#define CBGPU_DECLARE_SEMIRING(symbol, SR, T1, T2, TO)                                                                 \
  extern "C" int symbol(void);                                                                                         \
  namespace cbgpu_overlay {                                                                                            \
  template <> struct user_semiring<SR> {                                                                               \
    static constexpr bool value = true;                                                                                \
    static constexpr int a_dtype = dtype_of<T1>::value, b_dtype = dtype_of<T2>::value, c_dtype = dtype_of<TO>::value;  \
    static int id() { return symbol(); }                                                                               \
  };                                                                                                                   \
  }

#endif
