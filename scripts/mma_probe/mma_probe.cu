// Micro-benchmark (measurement infrastructure, not part of the library): issue rate of tcgen05.mma kind::f16,
// M = 128, K = 16 per instruction, as a function of N, of the A operand source (shared memory descriptor vs tensor
// memory) and of how many SMs run it at the same time (1 vs all: separates a per-SM pipe limit from chip-level
// power management).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tensorrec_b200/csrc
//   -o scripts/mma_probe/mma_probe scripts/mma_probe/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include "common.cuh"

using namespace trk;

template <int N, bool kTS>
__global__ void __launch_bounds__(128, 1) probe(int tiles, long long* out, int random_data) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  // A: 2 k-blocks of 128 x 64 fp16 (16 KB each) at 0; B: 2 k-blocks of N x 64 fp16 at 32 KB
  __half* h = reinterpret_cast<__half*>(smem);
  for (int i = threadIdx.x; i < (32768 + 2 * N * 128) / 2; i += blockDim.x)
  {
    uint32_t x = (i + 1) * 2654435761u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    // random_data: full-entropy sign + mantissa, exponent spread over 2^-3 .. 2^0 (what a real operand looks like)
    const uint16_t bits = static_cast<uint16_t>((x & 0x83ffu) | ((12u + ((x >> 10) & 3u)) << 10));
    h[i] = random_data ? __ushort_as_half(bits) : __float2half(static_cast<float>((x >> 20) % 17) - 8.0f);
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  {   // A operand into tensor memory columns [0, 64): lane = row, two fp16 per column
    uint32_t r[32];
    for (int kb = 0; kb < 2; ++kb) {
      for (int i = 0; i < 32; ++i) {
        uint32_t x = (threadIdx.x * 64 + kb * 32 + i + 7) * 2654435761u;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        r[i] = random_data ? ((x & 0x83ff83ffu) | 0x34003000u) : 0x3c003c00u + (threadIdx.x * 37 + i * 11) % 5;
      }
      tmem_st_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + kb * 32, r);
    }
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_f16_f32(128, N);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 32768);
    constexpr int kSlots = (512 - 64) / N >= 2 ? 2 : 1;
    t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const int slot = t % kSlots;
      if (t >= 2) mbar_wait(bars + (t & 1), ((t >> 1) - 1) & 1);
      tcgen05_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = umma_desc_k_major_sw128(a_base + kb * 16384);
          const uint64_t db = umma_desc_k_major_sw128(b_base + kb * N * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (kTS)
              umma_f16_ts(tmem_base + 64 + slot * N, tmem_base + kb * 32 + ks * 8, db + 2u * ks, idesc, kb | ks);
            else
              umma_f16_ss(tmem_base + 64 + slot * N, da + 2u * ks, db + 2u * ks, idesc, kb | ks);
          }
        }
        umma_commit(bars + (t & 1));
      }
      __syncwarp();
    }
    const int last = tiles - 1;
    mbar_wait(bars + (last & 1), (last >> 1) & 1);
    if (tiles >= 2) mbar_wait(bars + ((last - 1) & 1), ((last - 1) >> 1) & 1);
    t1 = clock64();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int N, bool kTS>
void run(int grid, long long* d_out, int tiles = 40000, int random_data = 0) {
  const int smem = 1024 + 32768 + 2 * N * 128 + 1024;
  cudaFuncSetAttribute(probe<N, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<N, kTS><<<grid, 128, 200 * 1024>>>(100, d_out, random_data);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  probe<N, kTS><<<grid, 128, 200 * 1024>>>(tiles, d_out, random_data);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long clk = 0;
  cudaMemcpy(&clk, d_out, sizeof(clk), cudaMemcpyDeviceToHost);
  const double per_mma = double(clk) / (8.0 * tiles);
  const double flops = 2.0 * 128 * N * 128 * double(tiles) * grid;
  printf("%s N=%3d A from %s, %3d SMs: %6.1f clk per MMA (math floor %3d)  %7.1f TFLOP/s  %.0f MHz  (%.2f ms, %s)\n", random_data ? "random" : "small-int", N,
         kTS ? "tmem" : "smem", grid, per_mma, N / 2, flops / (ms * 1e-3) / 1e12, clk / (ms * 1e3), ms,
         cudaGetErrorString(e));
  (void)smem;
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  for (int grid : {1, 148}) {
    run<64, false>(grid, d_out);
    run<64, true>(grid, d_out);
    run<128, false>(grid, d_out);
    run<128, true>(grid, d_out);
    run<256, false>(grid, d_out);
    run<256, true>(grid, d_out);
  }
  // sustained: ~0.4 s per configuration on all SMs (power management has time to act)
  run<128, false>(148, d_out, 1400000);
  run<128, true>(148, d_out, 1400000);
  run<256, false>(148, d_out, 700000);
  run<256, true>(148, d_out, 700000);
  run<128, false>(148, d_out, 1400000, 1);
  run<128, true>(148, d_out, 1400000, 1);
  run<256, false>(148, d_out, 700000, 1);
  run<256, true>(148, d_out, 700000, 1);
  run<128, true>(1, d_out, 1400000, 1);
  return 0;
}
