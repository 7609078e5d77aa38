"""Generates tmem_probe.cu: a micro-benchmark of tcgen05.ld throughput (TMEM -> registers) on sm_100a as a function of
the load width (x16 .. x128), the number of loads in flight per warp and the number of warps per SM.  Measurement
infrastructure only (not part of the library).  Build + run:  python scripts/tmem_probe/gen.py && ./scripts/tmem_probe/tmem_probe
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def ld(x):
    regs = ', '.join('%%%d' % i for i in range(x))
    outs = ', '.join('"=r"(r[%d])' % i for i in range(x))
    return ('__device__ __forceinline__ void ld_x%d(uint32_t taddr, uint32_t* r) {\n'
            '  asm volatile("tcgen05.ld.sync.aligned.32x32b.x%d.b32 {%s}, [%%%d];" : %s : "r"(taddr) : "memory");\n}\n'
            % (x, x, regs, x, outs))


SRC = r'''
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
%(lds)s
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int X, int DEPTH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(int iters, long long* clocks, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  extern __shared__ uint8_t pad[];
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + (static_cast<uint32_t>((warp %% 4) * 32) << 16);
  uint32_t r[DEPTH][X];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = ((warp / 4) * 128 + it * X * DEPTH) %% 512;
#pragma unroll
    for (int dd = 0; dd < DEPTH; ++dd) {
      uint32_t c = (col + dd * X) %% 512;
      if (c + X > 512) c = 0;
      if (X == 16) ld_x16(base + c, r[dd]);
      if (X == 32) ld_x32(base + c, r[dd]);
      if (X == 64) ld_x64(base + c, r[dd]);
      if (X == 128) ld_x128(base + c, r[dd]);
    }
    ld_wait();
#pragma unroll
    for (int dd = 0; dd < DEPTH; ++dd) acc ^= r[dd][0] ^ r[dd][X - 1];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %%0, 512;" ::"r"(tmem_base_s) : "memory");
}

template <int X, int DEPTH, int WARPS>
void run(long long* d_clk, uint32_t* d_sink) {
  const int warps = WARPS;
  const int iters = 20000;
  cudaFuncSetAttribute(probe<X, DEPTH, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<X, DEPTH, WARPS><<<148, warps * 32, 200 * 1024>>>(100, d_clk, d_sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  probe<X, DEPTH, WARPS><<<148, warps * 32, 200 * 1024>>>(iters, d_clk, d_sink);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long clk[148];
  cudaMemcpy(clk, d_clk, sizeof(clk), cudaMemcpyDeviceToHost);
  const double bytes = double(iters) * DEPTH * X * 4 * 32 * warps;
  printf("x%%-3d depth %%d warps %%2d: %%7.1f B/clk/SM  %%7.1f GB/s/SM  (%%.2f ms, %%lld clk, %%s)\n", X, DEPTH, warps,
         bytes / double(clk[0]), bytes / (ms * 1e-3) / 1e9, ms, clk[0], cudaGetErrorString(e));
}

int main() {
  long long* d_clk;
  uint32_t* d_sink;
  cudaMalloc(&d_clk, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 4);
#define SWEEP(W)                    \
  run<16, 1, W>(d_clk, d_sink);     \
  run<16, 2, W>(d_clk, d_sink);     \
  run<16, 4, W>(d_clk, d_sink);     \
  run<32, 1, W>(d_clk, d_sink);     \
  run<32, 2, W>(d_clk, d_sink);     \
  run<64, 1, W>(d_clk, d_sink);
  SWEEP(4)
  SWEEP(8)
  SWEEP(16)
  run<32, 4, 4>(d_clk, d_sink);
  run<64, 2, 4>(d_clk, d_sink);
  run<128, 1, 4>(d_clk, d_sink);
  run<32, 4, 8>(d_clk, d_sink);
  run<64, 2, 8>(d_clk, d_sink);
  run<128, 1, 8>(d_clk, d_sink);
  return 0;
}
'''

if __name__ == '__main__':
    src = SRC % {'lds': '\n'.join(ld(x) for x in (16, 32, 64, 128))}
    path = os.path.join(HERE, 'tmem_probe.cu')
    with open(path, 'w') as f:
        f.write(src)
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-o',
                           os.path.join(HERE, 'tmem_probe'), path])
