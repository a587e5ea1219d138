// Micro-benchmarks of the accumulation primitives the numeric kernels can choose from (B200, sm_100a).
// Reports cycles per lane-operation per SM at full occupancy (time * sm_clock * num_SMs / total lane-ops).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prims prims.cu ; run on one GPU.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

constexpr int ITER = 2048;
constexpr int SLOTS = 16384; // 128 KB of doubles per CTA region

// mode 0: RED.ADD.F64 to global (region per CTA, L2 resident)
// mode 1: atomicAdd(double) shared  (CAS loop)
// mode 2: non-atomic LDS.64 + DADD + STS.64 shared (racy; throughput only)
// mode 3: atomicOr 32-bit shared
// mode 4: atomicAdd int shared
// mode 5: match_any on the slot id + leader RMW
// mode 6: LDS.64 only (random)
// mode 7: atomicAdd u64 shared
// mode 8: atomicExch u64 shared
// mode 9: LDS.32 only (random)
// mode 10: STS.64 only (random)
template <int MODE>
__global__ void __launch_bounds__(512) prim_kernel(double *g, unsigned long long *sink, int slots) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *acc = reinterpret_cast<double *>(smem_raw);
  unsigned *acc32 = reinterpret_cast<unsigned *>(smem_raw);
  unsigned long long *acc64 = reinterpret_cast<unsigned long long *>(smem_raw);
  for (int i = threadIdx.x; i < slots; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  double *mine = g + (size_t)blockIdx.x * slots;
  unsigned s = mix(blockIdx.x * 1024u + threadIdx.x + 12345u);
  double local = 0;
  unsigned long long lsum = 0;
#pragma unroll 4
  for (int it = 0; it < ITER; ++it) {
    s = s * 1664525u + 1013904223u;
    const unsigned slot = (s >> 8) % (unsigned)slots;
    const double v = (double)(s & 255u);
    if (MODE == 0) atomicAdd(&mine[slot], v);
    else if (MODE == 1) atomicAdd(&acc[slot], v);
    else if (MODE == 2) acc[slot] += v;
    else if (MODE == 3) atomicOr(&acc32[slot], 1u << (s & 31));
    else if (MODE == 4) atomicAdd(&acc32[slot], 1u);
    else if (MODE == 5) {
      unsigned m = __match_any_sync(0xFFFFFFFFu, slot);
      if ((__ffs(m) - 1) == (int)(threadIdx.x & 31)) acc[slot] += v * __popc(m);
    } else if (MODE == 6) local += acc[slot];
    else if (MODE == 7) atomicAdd(&acc64[slot], (unsigned long long)(s & 255u));
    else if (MODE == 8) lsum += atomicExch(&acc64[slot], (unsigned long long)s);
    else if (MODE == 9) lsum += acc32[slot];
    else if (MODE == 10) acc[slot] = v;
  }
  __syncthreads();
  if (MODE != 0) {
    double t = 0;
    for (int i = threadIdx.x; i < slots; i += blockDim.x) t += acc[i];
    local += t;
  }
  if (local == 1.2345e-300 || lsum == 0x123456789ull) sink[0] = 1;
}

// token ring: 16 warps; every warp prepares K random (slot, v) per lane (distinct slots inside one warp-instruction are
// NOT guaranteed here: throughput only), waits for its turn, applies K non-atomic RMWs, passes the token.
template <int K>
__global__ void __launch_bounds__(512) ring_kernel(unsigned long long *sink, int slots, int rounds) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *acc = reinterpret_cast<double *>(smem_raw);
  for (int i = threadIdx.x; i < slots; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  unsigned s = mix(blockIdx.x * 1024u + threadIdx.x + 999u);
  // barrier ids 1..nwarp: barrier (w+1) is "warp w may go"; first turn of warp 0 needs no wait
  for (int r = 0; r < rounds; ++r) {
    unsigned slot[K];
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      s = s * 1664525u + 1013904223u;
      slot[k] = (s >> 8) % (unsigned)slots;
      v[k] = (double)(s & 255u);
    }
    if (!(r == 0 && warp == 0)) asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
    double o[K];
#pragma unroll
    for (int k = 0; k < K; ++k) o[k] = acc[slot[k]];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[slot[k]] = o[k] + v[k];
    const int next = (warp + 1) % nwarp;
    if (!(r == rounds - 1 && warp == nwarp - 1)) asm volatile("bar.arrive %0, 64;" ::"r"(next + 1) : "memory");
  }
  __syncthreads();
  double t = 0;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) t += acc[i];
  if (t == 1.2345e-300) sink[0] = 1;
}

template <int MODE>
int run(const char *name, int ctas_per_sm, int slots, double clock_ghz, int nsm, double *g, unsigned long long *sink) {
  const int grid = nsm * ctas_per_sm;
  size_t sm = (size_t)slots * 8;
  CHECK(cudaFuncSetAttribute(prim_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  prim_kernel<MODE><<<grid, 512, sm>>>(g, sink, slots);
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  prim_kernel<MODE><<<grid, 512, sm>>>(g, sink, slots);
  cudaEventRecord(b);
  CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)grid * 512 * ITER;
  printf("%-34s ctas/SM %d slots %6d : %8.3f ms  %.3f cyc/lane-op/SM\n", name, ctas_per_sm, slots, ms, ms * 1e-3 * clock_ghz * 1e9 * nsm / ops);
  return 0;
}

template <int K>
int run_ring(int ctas_per_sm, int slots, double clock_ghz, int nsm, unsigned long long *sink) {
  const int grid = nsm * ctas_per_sm, rounds = 4096 / K;
  size_t sm = (size_t)slots * 8;
  CHECK(cudaFuncSetAttribute(ring_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  ring_kernel<K><<<grid, 512, sm>>>(sink, slots, rounds);
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  ring_kernel<K><<<grid, 512, sm>>>(sink, slots, rounds);
  cudaEventRecord(b);
  CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)grid * 512 * rounds * K;
  printf("token ring K=%d                     ctas/SM %d slots %6d : %8.3f ms  %.3f cyc/lane-op/SM\n", K, ctas_per_sm, slots, ms, ms * 1e-3 * clock_ghz * 1e9 * nsm / ops);
  return 0;
}

int main() {
  cudaDeviceProp p;
  CHECK(cudaGetDeviceProperties(&p, 0));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  const int nsm = p.multiProcessorCount;
  printf("%s, %d SMs, %.3f GHz\n", p.name, nsm, ghz);
  double *g; unsigned long long *sink;
  CHECK(cudaMalloc(&g, (size_t)nsm * 4 * SLOTS * 8));
  CHECK(cudaMemset(g, 0, (size_t)nsm * 4 * SLOTS * 8));
  CHECK(cudaMalloc(&sink, 8));
  for (int c = 1; c <= 3; ++c) {
    const int slots = c == 1 ? SLOTS : (c == 2 ? 12288 : 8192);
    run<0>("RED.ADD.F64 global (L2)", c, slots, ghz, nsm, g, sink);
    run<1>("atomicAdd f64 shared (CAS)", c, slots, ghz, nsm, g, sink);
    run<2>("LDS+DADD+STS f64 (non-atomic)", c, slots, ghz, nsm, g, sink);
    run<3>("atomicOr u32 shared", c, slots, ghz, nsm, g, sink);
    run<4>("atomicAdd u32 shared", c, slots, ghz, nsm, g, sink);
    run<5>("match_any + leader RMW", c, slots, ghz, nsm, g, sink);
    run<6>("LDS.64 random", c, slots, ghz, nsm, g, sink);
    run<7>("atomicAdd u64 shared", c, slots, ghz, nsm, g, sink);
    run<8>("atomicExch u64 shared", c, slots, ghz, nsm, g, sink);
    run<9>("LDS.32 random", c, slots, ghz, nsm, g, sink);
    run<10>("STS.64 random", c, slots, ghz, nsm, g, sink);
    run_ring<4>(c, slots, ghz, nsm, sink);
    run_ring<8>(c, slots, ghz, nsm, sink);
    run_ring<16>(c, slots, ghz, nsm, sink);
  }
  return 0;
}
