// fetch_granularity.cu -- how many random 16-byte reads per second a B200 serves out of HBM (1 GiB / 4 GiB footprint),
// by load flavour and by cudaLimitMaxL2FetchGranularity.  Background: the ncu sweep of k_probe2 over presence-filter
// sizes (profiles/README.md, round 2) shows ~124 bytes of DRAM traffic per filter-block fetch (a 16-byte read in a
// 1 GiB array), i.e. every sector miss fills a whole 128-byte line; if a flavour fetched less, the random-access floor
// of the probe would move.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fetch_granularity fetch_granularity.cu
//   build/fetch_granularity [granularity_bytes]     (the limit is set before the context exists)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}

// MODE 0: ld.global.nc + evict_first hint (the probe's block load)   1: plain ld.global   2: ld.global.cg
//      3: ld.global.cv   4: ld.global.nc.L1::no_allocate   5: ld.global.lu   6: ld.global.nc.L2::64B
//      7: ld.global.cs   8: ld.relaxed.gpu
template <int MODE>
__device__ __forceinline__ uint4 load16(const uint4* p, uint64_t pol) {
  uint4 x;
  if (MODE == 0) asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p), "l"(pol));
  if (MODE == 1) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  if (MODE == 5) asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  if (MODE == 6) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  if (MODE == 7) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  if (MODE == 8) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) k_gather(const uint4* __restrict__ base, uint64_t n_units, uint32_t iters, uint32_t* sink) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B1u + 12345u;
  uint32_t acc = 0;
  const uint64_t mask = n_units - 1;
  for (uint32_t it = 0; it < iters; it++) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      seed = mix(seed + 0x632BE5ABu);
      const uint64_t u = ((uint64_t)seed | ((uint64_t)mix(seed ^ 0x5BD1E995u) << 32)) & mask;
      const uint4 x = load16<MODE>(base + u, pol);
      v[j] = x.x ^ x.w;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) acc += v[j];
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <int MODE>
static double run(const uint4* buf, uint64_t bytes, uint32_t* sink, int n_sm) {
  const uint64_t n_units = bytes / 16;
  const uint32_t iters = 256;
  const int grid = n_sm * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_gather<MODE><<<grid, 256>>>(buf, n_units, iters / 4, sink);
  cudaEventRecord(e0);
  k_gather<MODE><<<grid, 256>>>(buf, n_units, iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double)grid * 256 * iters * 8 / (ms * 1e-3) / 1e9;
}

int main(int argc, char** argv) {
  int gran = argc > 1 ? atoi(argv[1]) : 0;
  if (gran) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
    if (e != cudaSuccess) fprintf(stderr, "cudaDeviceSetLimit: %s\n", cudaGetErrorString(e));
  }
  size_t got = 0;
  cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
  const uint64_t max_bytes = 4ull << 30;
  uint4* buf;
  uint32_t* sink;
  if (cudaMalloc(&buf, max_bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
  cudaMemset(buf, 1, max_bytes);
  const int n_sm = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"l2_fetch_granularity_limit\": %zu, \"unit\": \"G random 16-byte loads/s\", \"rows\": [\n", prop.name, got);
  const uint64_t sizes[] = {1ull << 30, 4ull << 30};
  for (int s = 0; s < 2; s++) {
    const uint64_t sz = sizes[s];
    printf(" {\"footprint_mib\": %llu, \"nc_evict_first\": %.1f, \"plain\": %.1f, \"cg\": %.1f, \"cv\": %.1f, \"nc_L1_no_allocate\": %.1f, \"lu\": %.1f, "
           "\"nc_L2_64B\": %.1f, \"cs\": %.1f, \"relaxed_gpu\": %.1f}%s\n",
           (unsigned long long)(sz >> 20), run<0>(buf, sz, sink, n_sm), run<1>(buf, sz, sink, n_sm), run<2>(buf, sz, sink, n_sm),
           run<3>(buf, sz, sink, n_sm), run<4>(buf, sz, sink, n_sm), run<5>(buf, sz, sink, n_sm), run<6>(buf, sz, sink, n_sm),
           run<7>(buf, sz, sink, n_sm), run<8>(buf, sz, sink, n_sm), s == 0 ? "," : "");
  }
  printf("]}\n");
  return 0;
}
