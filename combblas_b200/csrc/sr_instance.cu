// One translation unit per semiring (-DCB_SR=<id>): instantiates the accumulation engine for the local
// multiply (LocalHybridSpGEMM / LocalSpGEMMHash, mtSpGEMM.h:213,:463) and the k-way merge
// (MultiwayMerge / MultiwayMergeHash, MultiwayMerge.h:428,:553) of that semiring.
#include "sr_instance.cuh"

#ifndef CB_SR
#error "compile with -DCB_SR=<semiring id>"
#endif

namespace cbgpu {

typedef Semiring<CB_SR> SR;

static int library_out_dtype() {
  int ta, tb, tc;
  semiring_types(CB_SR, &ta, &tb, &tc);
  return tc;
}

#define CB_CAT2(a, b) a##b
#define CB_CAT(a, b) CB_CAT2(a, b)
int CB_CAT(spgemm_sr, CB_SR)(const SpgemmArgs &a) { return spgemm_impl<SR>(a, library_out_dtype()); }
int CB_CAT(merge_sr, CB_SR)(const MergeArgs &a) { return merge_impl<SR>(a); }

} // namespace cbgpu

#if defined(CBGPU_PHASE_TIMING) && CB_SR == 0
// tuning builds only: read (and clear) the per-phase cycle counters of the shared-accumulator kernels of semiring 0
extern "C" int cbgpu_debug_phase_cycles(unsigned long long *out /* 3 x 8 */) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, cbgpu::g_phase_cycles, sizeof(unsigned long long) * 24) != cudaSuccess) return -2;
  unsigned long long zero[24] = {0};
  return cudaMemcpyToSymbol(cbgpu::g_phase_cycles, zero, sizeof(zero)) == cudaSuccess ? 0 : -2;
}
#endif
