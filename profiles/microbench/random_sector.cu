// random_sector.cu -- what the memory system of a B200 delivers for the probe kernel's access pattern:
// independent, uniformly random 4-byte (presence word) or 16-byte (filter block / half a table bucket) reads,
// each touching one 32-byte sector, over footprints from L2-resident to HBM-sized.  The numbers are the
// denominators of the "random-access floor" in profiles/README.md: k_probe2's DRAM traffic is 12 % streaming
// and 88 % such sectors, for which the copy-benchmark peak (MEASURED_PEAKS.json) is not reachable.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/random_sector random_sector.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint64_t policy(int keep) {
  uint64_t p;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// every thread: `iters` rounds of ILP independent loads; VEC = 1 (u32) or 4 (uint4)
template <int VEC, int ILP>
__global__ void __launch_bounds__(256, 4) k_gather(const uint32_t* __restrict__ base, uint64_t n_units, uint32_t iters, int keep,
                                                   uint32_t* sink) {
  const uint64_t pol = policy(keep);
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B1u + 12345u;
  uint32_t acc = 0;
  const uint64_t mask = n_units - 1;
  for (uint32_t it = 0; it < iters; it++) {
    uint32_t v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
      seed = mix(seed + 0x632BE5ABu);
      uint64_t u = ((uint64_t)seed | ((uint64_t)mix(seed ^ 0x5BD1E995u) << 32)) & mask;
      if (VEC == 4) {
        uint4 x;
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"((const uint4*)base + u), "l"(pol));
        v[j] = x.x ^ x.w;
      } else {
        asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v[j]) : "l"(base + u), "l"(pol));
      }
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) acc += v[j];
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <int VEC, int ILP>
static double run(const uint32_t* buf, uint64_t bytes, int keep, uint32_t* sink, int n_sm) {
  const uint64_t n_units = bytes / (4 * VEC);
  const uint32_t iters = 2048 / ILP;
  const int grid = n_sm * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_gather<VEC, ILP><<<grid, 256>>>(buf, n_units, iters / 4, keep, sink);  // warm-up
  cudaEventRecord(e0);
  k_gather<VEC, ILP><<<grid, 256>>>(buf, n_units, iters, keep, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double loads = (double)grid * 256 * iters * ILP;
  return loads / (ms * 1e-3) / 1e9;  // G loads/s
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
  const uint64_t max_bytes = 4ull << 30;
  uint32_t *buf, *sink;
  if (cudaMalloc(&buf, max_bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
  cudaMemset(buf, 1, max_bytes);
  printf("{\"device\": \"%s\", \"sms\": %d, \"unit\": \"G loads/s (one 32-byte sector each); GB/s = sectors x 32 B\", \"rows\": [\n", prop.name,
         prop.multiProcessorCount);
  const uint64_t sizes[] = {16ull << 20, 64ull << 20, 256ull << 20, 1ull << 30, 4ull << 30};
  bool first = true;
  for (uint64_t sz : sizes) {
    for (int keep = 1; keep >= 0; keep--) {
      double g4 = run<1, 4>(buf, sz, keep, sink, prop.multiProcessorCount);
      double g16 = run<4, 4>(buf, sz, keep, sink, prop.multiProcessorCount);
      double g16_8 = run<4, 8>(buf, sz, keep, sink, prop.multiProcessorCount);
      printf("%s {\"footprint_mib\": %llu, \"l2_policy\": \"%s\", \"u32_ilp4\": %.1f, \"uint4_ilp4\": %.1f, \"uint4_ilp8\": %.1f, "
             "\"uint4_ilp8_sector_gbs\": %.0f}",
             first ? "" : ",\n", (unsigned long long)(sz >> 20), keep ? "evict_last" : "evict_first", g4, g16, g16_8, g16_8 * 32);
      first = false;
    }
  }
  printf("\n]}\n");
  return 0;
}
