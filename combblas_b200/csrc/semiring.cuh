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
//   acc_add(a, b)      SR::add on two accumulator values
//   accumulate_shared<FIRST>(p, v)  *p = SR::add(v, *p) for p in SHARED memory, safe against every other thread of the CTA
//                      (see exch_accumulate below); FIRST = most slots receive one product only
//   to_out / from_out  conversion between acc_t and out_t
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <string.h>
#include <type_traits>

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


// Lock-free accumulation into SHARED memory built on ATOMS.EXCH, which is native for 32 and 64 bits on sm_100a, where
// atomicAdd(double/float/u64) and atomicMax/Min(64 bit) compile to an ATOMS.CAST.SPIN loop. Measured on B200
// (tools/micro/prims.cu, profiles/r2_micro_prims.txt): EXCH.64 0.21 cycles per lane and SM, the CAS loop 0.79, a
// RED.ADD.F64 to L2 1.52. The identity of SR::add marks an empty slot -- it carries no information, so a slot that
// holds it and an empty slot are the same thing. A thread owns whatever it has swapped OUT of a slot; the sum of slot +
// all owned values never changes, and a thread finishes only when it has swapped its value INTO an empty slot.
template <class SR, bool FIRST>
__device__ __forceinline__ void exch_accumulate(typename SR::acc_t *p, typename SR::acc_t v) {
  typedef typename SR::acc_t T;
  typedef typename std::conditional<sizeof(T) == 8, unsigned long long, unsigned int>::type U;
  static_assert(sizeof(T) == sizeof(U), "accumulators are 4 or 8 bytes");
  U *up = reinterpret_cast<U *>(p);
  T idv = SR::identity();
  U ident, x;
  memcpy(&ident, &idv, sizeof(U));
  memcpy(&x, &v, sizeof(U));
  while (true) {
    U old;
    if (!FIRST) { // take what is there (usually a partial sum), add, put back
      old = atomicExch(up, ident);
      T a, b;
      memcpy(&a, &old, sizeof(U));
      memcpy(&b, &x, sizeof(U));
      a = SR::acc_add(a, b);
      memcpy(&x, &a, sizeof(U));
      old = atomicExch(up, x);
      if (old == ident) return;
      x = old; // somebody put a value in between: it is mine now
    } else { // put first (usually into an empty slot); otherwise merge what came out with what is there now
      old = atomicExch(up, x);
      if (old == ident) return;
      const U y = atomicExch(up, ident);
      T a, b;
      memcpy(&a, &old, sizeof(U));
      memcpy(&b, &y, sizeof(U));
      a = SR::acc_add(a, b);
      memcpy(&x, &a, sizeof(U));
    }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a + b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { exch_accumulate<Semiring, FIRST>(p, v); }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a + b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { exch_accumulate<Semiring, FIRST>(p, v); }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a + b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { exch_accumulate<Semiring, FIRST>(p, v); }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a > b ? a : b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { exch_accumulate<Semiring, FIRST>(p, v); }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return b < a ? b : a; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { exch_accumulate<Semiring, FIRST>(p, v); }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a | b; }
  // OR: every writer stores the same value, a plain store is enough
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { if (v) *reinterpret_cast<volatile acc_t *>(p) = 1u; }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a + b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { exch_accumulate<Semiring, FIRST>(p, v); }
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a + b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { atomicAdd(p, v); } // ATOMS.ADD is native for 32 bits
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
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return a > b ? a : b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { exch_accumulate<Semiring, FIRST>(p, v); }
  __device__ static __forceinline__ out_t to_out(acc_t v) { return v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return v; }
};

// 9..14: BoolCopy2ndSRing<OUT> / BoolCopy1stSRing<OUT> (Semirings.h:51-138), the pair SpParMat::SubsRef_SR multiplies with
// (S * A * T with boolean selection matrices, SpParMat.cpp:2515-2566): multiply copies the non-boolean operand, and add
// "should not happen" -- the reference throws from it, because a selection matrix has one entry per row / column and every
// output therefore receives exactly one product. The device path stores the product (bit-exact, no arithmetic); the engine
// and the merge compare products with outputs afterwards and fail with the reference's message if an add would have happened.
template <class A, class B, class ACC, class OUT, bool SECOND>
struct BoolCopySemiring {
  typedef A a_t; typedef B b_t; typedef ACC acc_t; typedef OUT out_t;
  static constexpr bool kAddForbidden = true;
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return SECOND ? (acc_t)b : (acc_t)a; }
  __device__ static __forceinline__ acc_t identity() { return acc_t(); } // OUT(): BoolCopy*SRing::id()
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) { *p = v; }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) { *p = (out_t)v; }
  __device__ static __forceinline__ out_t add(out_t, out_t b) { return b; } // unreachable in a valid product
  __device__ static __forceinline__ acc_t acc_add(acc_t, acc_t b) { return b; }
  template <bool FIRST> __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) { *p = v; }
  __device__ static __forceinline__ out_t to_out(acc_t v) { return (out_t)v; }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return (acc_t)v; }
};
template <> struct Semiring<9> : BoolCopySemiring<uint8_t, double, double, double, true> {};                    // BoolCopy2ndSRing<double>
template <> struct Semiring<10> : BoolCopySemiring<double, uint8_t, double, double, false> {};                  // BoolCopy1stSRing<double>
template <> struct Semiring<11> : BoolCopySemiring<uint8_t, long long, long long, long long, true> {};          // BoolCopy2ndSRing<int64_t>
template <> struct Semiring<12> : BoolCopySemiring<long long, uint8_t, long long, long long, false> {};         // BoolCopy1stSRing<int64_t>
template <> struct Semiring<13> : BoolCopySemiring<uint8_t, uint8_t, unsigned int, uint8_t, true> {};           // BoolCopy2ndSRing<bool>
template <> struct Semiring<14> : BoolCopySemiring<uint8_t, uint8_t, unsigned int, uint8_t, false> {};          // BoolCopy1stSRing<bool>

// semirings whose add must never run (kAddForbidden): checked by the host code after the symbolic pass / the merge count
template <class SR, class = void>
struct add_forbidden : std::false_type {};
template <class SR>
struct add_forbidden<SR, typename std::enable_if<SR::kAddForbidden>::type> : std::true_type {};

} // namespace cbgpu
