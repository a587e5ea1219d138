// User-defined semirings on the device path, nvcc half. See semiring_decl.h for the recipe.
//
// replaces: the SR template parameter of LocalHybridSpGEMM / LocalSpGEMMHash / MultiwayMerge (mtSpGEMM.h:213, :463,
// MultiwayMerge.h:428) for semirings that are not in the library's list (Semirings.h:143-255 are; a driver's own struct, e.g.
// ReleaseTests/KTipsTest.cpp:12-20, is not). The struct keeps the reference's interface
//     static TO id();  static TO add(const TO&, const TO&);  static TO multiply(const T1&, const T2&);
// with CBGPU_HD on the members; this header wraps it into the functor interface of the accumulation engine
// (csrc/semiring.cuh), instantiates the engine and the streaming merge for it in THIS translation unit, and registers the two
// entry points with libcbgpu.so under a fresh semiring id when the exported symbol is first called.
//
// Compile:  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
//                --cudart shared -I<repo>/include my_semiring.cu -L<repo>/combblas_b200 -lcbgpu -o libmy_semiring.so
// (--cudart shared: libcbgpu.so uses the shared CUDA runtime; one runtime instance per process keeps kernel handles, streams
// and events of both libraries in one place)
// The engine headers are internal to the library: rebuild this unit whenever libcbgpu.so is rebuilt (CBGPU_VERSION is checked).
#pragma once
#include "semiring_decl.h"
#include "../../combblas_b200/csrc/sr_instance.cuh"

namespace cbgpu {

template <class T> struct stored_as { typedef T type; };
template <> struct stored_as<bool> { typedef uint8_t type; }; // bool travels as one byte (cbgpu.h)

template <class T> struct dtype_code;
template <> struct dtype_code<double> { static constexpr int value = CBGPU_F64; };
template <> struct dtype_code<float> { static constexpr int value = CBGPU_F32; };
template <> struct dtype_code<int64_t> { static constexpr int value = CBGPU_I64; };
template <> struct dtype_code<int32_t> { static constexpr int value = CBGPU_I32; };
template <> struct dtype_code<bool> { static constexpr int value = CBGPU_BOOL; };

// The engine's view of a reference-style semiring struct. Accumulators are the raw bits of a TO in a 4- or 8-byte word, so
// that one compare-and-swap (global memory, open-addressing tables) or one exchange (shared-memory accumulators,
// exch_accumulate) moves a whole value whatever TO is.
template <class SR, class T1, class T2, class TO>
struct UserSemiringAdapter {
  typedef typename stored_as<T1>::type a_t;
  typedef typename stored_as<T2>::type b_t;
  typedef typename stored_as<TO>::type out_t;
  typedef typename std::conditional<sizeof(TO) == 8, unsigned long long, unsigned int>::type acc_t;
  static_assert(sizeof(TO) == 8 || sizeof(TO) == 4 || sizeof(TO) == 1, "value types: double, float, int64_t, int32_t, bool");

  __device__ static __forceinline__ acc_t pack(TO v) {
    acc_t r = 0;
    if constexpr (sizeof(TO) == 1) r = (acc_t)(v ? 1u : 0u);
    else memcpy(&r, &v, sizeof(TO));
    return r;
  }
  __device__ static __forceinline__ TO unpack(acc_t r) {
    if constexpr (sizeof(TO) == 1) return (TO)(r != 0);
    else {
      TO v;
      memcpy(&v, &r, sizeof(TO));
      return v;
    }
  }
  __device__ static __forceinline__ acc_t mul(a_t a, b_t b) { return pack(SR::multiply((T1)a, (T2)b)); }
  __device__ static __forceinline__ acc_t identity() { return pack(SR::id()); }
  __device__ static __forceinline__ acc_t acc_add(acc_t a, acc_t b) { return pack(SR::add(unpack(a), unpack(b))); }
  // *p = SR::add(v, *p): the argument order of the reference's hash kernel (mtSpGEMM.h:408)
  __device__ static __forceinline__ void accumulate(acc_t *p, acc_t v) {
    acc_t old = *reinterpret_cast<volatile acc_t *>(p);
    while (true) {
      const acc_t want = acc_add(v, old);
      if (want == old) return;
      const acc_t prev = atomicCAS(p, old, want);
      if (prev == old) return;
      old = prev;
    }
  }
  __device__ static __forceinline__ void accumulate_out(out_t *p, acc_t v) {
    if constexpr (sizeof(out_t) == sizeof(acc_t)) {
      accumulate(reinterpret_cast<acc_t *>(p), v);
    } else { // one byte of C: compare-and-swap on the aligned word that holds it
      const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
      unsigned int *word = reinterpret_cast<unsigned int *>(addr & ~(uintptr_t)3);
      const unsigned shift = (unsigned)(addr & 3) * 8u;
      unsigned int old = *reinterpret_cast<volatile unsigned int *>(word);
      while (true) {
        const acc_t cur = (old >> shift) & 0xFFu;
        const acc_t want = acc_add(v, cur);
        if (want == cur) return;
        const unsigned int next = (old & ~(0xFFu << shift)) | ((unsigned int)(want & 0xFFu) << shift);
        const unsigned int prev = atomicCAS(word, old, next);
        if (prev == old) return;
        old = prev;
      }
    }
  }
  __device__ static __forceinline__ out_t add(out_t a, out_t b) { return (out_t)SR::add((TO)a, (TO)b); }
  template <bool FIRST>
  __device__ static __forceinline__ void accumulate_shared(acc_t *p, acc_t v) {
    exch_accumulate<UserSemiringAdapter, FIRST>(p, v);
  }
  __device__ static __forceinline__ out_t to_out(acc_t v) { return (out_t)unpack(v); }
  __device__ static __forceinline__ acc_t from_out(out_t v) { return pack((TO)v); }
};

template <class SR, class T1, class T2, class TO>
int user_spgemm_entry(const SpgemmArgs &a) {
  return spgemm_impl<UserSemiringAdapter<SR, T1, T2, TO>>(a, dtype_code<TO>::value);
}
template <class SR, class T1, class T2, class TO>
int user_merge_entry(const MergeArgs &a) {
  return merge_impl<UserSemiringAdapter<SR, T1, T2, TO>>(a);
}
// id of the semiring, registering it on the first call; negative cbgpu_status when the library refuses it
template <class SR, class T1, class T2, class TO>
int user_semiring_id() {
  static const int id = (cbgpu_version() == CBGPU_VERSION)
                            ? register_user_semiring(&user_spgemm_entry<SR, T1, T2, TO>, &user_merge_entry<SR, T1, T2, TO>,
                                                     dtype_code<T1>::value, dtype_code<T2>::value, dtype_code<TO>::value)
                            : (int)CBGPU_ERR_UNSUPPORTED;
  return id;
}

} // namespace cbgpu

#define CBGPU_DEFINE_SEMIRING(symbol, SR, T1, T2, TO)                                                                  \
  extern "C" int symbol(void) { return ::cbgpu::user_semiring_id<SR, T1, T2, TO>(); }
