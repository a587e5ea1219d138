// Device semiring functors: the reference's static-struct semirings (Semirings.h:143-255;
// ReleaseTests/KTipsTest.cpp:12-20 for OR-AND) instantiated as device code.
//
// Interface used by the accumulation engine:
//   a_t, b_t           operand value types as stored in HBM (bool = uint8_t)
//   acc_t              accumulator type used in shared / global memory (>= 4 bytes so atomics exist)
//   out_t              value type of C as stored in HBM
//   mul(a, b)          SR::multiply
//   identity()         a true identity of SR::add (so "first product stored, later ones added",
//                      mtSpGEMM.h:398-416, equals "initialise with identity, add everything")
//   accumulate(p, v)   *p = SR::add(v, *p), atomically; p may point to shared or global memory
//   accumulate_out(p,v) same, but p points into C's value array (out_t) in global memory
//   add(a, b)          SR::add on two stored values (used by the streaming 2-way merge)
//   to_out / from_out  conversion between acc_t and out_t
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace cbgpu {

__device__ __forceinline__ void atomic_min_f64(double *p, double v) {
  unsigned long long *up = reinterpret_cast<unsigned long long *>(p);
  unsigned long long old = *up;
  while (true) {
    double cur = __longlong_as_double((long long)old);
    if (!(v < cur)) break; // std::min(v, cur) keeps cur unless v < cur
    unsigned long long prev = atomicCAS(up, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

template <int ID>
struct Semiring;

// 0: PlusTimesSRing<double,double>
template <>
struct Semiring<0> {
  typedef double a_t; typedef double b_t; typedef double acc_t; typedef double out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return a * b; }
  __device__ static __forceinline__ acc_t identity() { return 0.0; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomicAdd(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return a + b; } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v; }
};
// 1: PlusTimesSRing<float,float>
template <>
struct Semiring<1> {
  typedef float a_t; typedef float b_t; typedef float acc_t; typedef float out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return a * b; }
  __device__ static __forceinline__ acc_t identity() { return 0.0f; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomicAdd(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return a + b; } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v; }
};
// 2: PlusTimesSRing<int64_t,int64_t> (wrapping arithmetic, as two's complement hardware does on the host)
template <>
struct Semiring<2> {
  typedef long long a_t; typedef long long b_t; typedef unsigned long long acc_t; typedef long long out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return (acc_t)a * (acc_t)b; }
  __device__ static __forceinline__ acc_t identity() { return 0ull; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomicAdd(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return (out_t)((acc_t)a + (acc_t)b); } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return (out_t)v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return (acc_t)v; }
};
// 3: SelectMaxSRing<bool,int64_t>: multiply(bool, x) = x (Semirings.h:200-203), add = max
template <>
struct Semiring<3> {
  typedef uint8_t a_t; typedef long long b_t; typedef long long acc_t; typedef long long out_t;
  __device__ static __forceinline__ acc_t mul(a_t, b_t b) { return b; }
  __device__ static __forceinline__ acc_t identity() { return LLONG_MIN; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomicMax(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return a > b ? a : b; } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v; }
};
// 4: MinPlusSRing<double,double>: multiply = inf_plus (Semirings.h:41-47), add = min
template <>
struct Semiring<4> {
  typedef double a_t; typedef double b_t; typedef double acc_t; typedef double out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return (a == DBL_MAX || b == DBL_MAX) ? DBL_MAX : a + b; }
  __device__ static __forceinline__ acc_t identity() { return __longlong_as_double(0x7FF0000000000000LL); } // +inf
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomic_min_f64(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return b < a ? b : a; } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v; }
};
// 5: OR-AND on bool
template <>
struct Semiring<5> {
  typedef uint8_t a_t; typedef uint8_t b_t; typedef unsigned int acc_t; typedef uint8_t out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return (a && b) ? 1u : 0u; }
  __device__ static __forceinline__ acc_t identity() { return 0u; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { if (v) atomicOr(p, 1u); }
  // OR into a byte of C: every writer stores the same value, so a plain store is race-free in effect
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { if (v) *reinterpret_cast<volatile uint8_t *>(p) = 1; }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return (out_t)((a || b) ? 1 : 0); } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v ? 1 : 0; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v ? 1u : 0u; }
};
// 6: PlusTimesSRing<bool,double>: static_cast<double>(bool) * x
template <>
struct Semiring<6> {
  typedef uint8_t a_t; typedef double b_t; typedef double acc_t; typedef double out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return (a ? 1.0 : 0.0) * b; }
  __device__ static __forceinline__ acc_t identity() { return 0.0; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomicAdd(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return a + b; } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v; }
};
// 7: PlusTimesSRing<int32_t,int32_t>
template <>
struct Semiring<7> {
  typedef int a_t; typedef int b_t; typedef unsigned int acc_t; typedef int out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return (acc_t)a * (acc_t)b; }
  __device__ static __forceinline__ acc_t identity() { return 0u; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomicAdd(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return (out_t)((acc_t)a + (acc_t)b); } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return (out_t)v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return (acc_t)v; }
};
// 8: SelectMaxSRing<int64_t,int64_t>: multiply = a*b, add = max
template <>
struct Semiring<8> {
  typedef long long a_t; typedef long long b_t; typedef long long acc_t; typedef long long out_t;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return (long long)((unsigned long long)a * (unsigned long long)b); }
  __device__ static __forceinline__ acc_t identity() { return LLONG_MIN; }
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { atomicMax(p, v); }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { accumulate(reinterpret_cast<acc_t *>(p), v); }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return a > b ? a : b; } // SR::add on stored values
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v; }
};

} // namespace cbgpu
