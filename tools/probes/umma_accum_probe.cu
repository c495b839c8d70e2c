// Hardware probe (not part of the product): is the ~95-cycle cost of a 128 x N x 16 fp16 MMA for N <= 192 (umma_rate_probe /
// umma_issue_probe) a DEPENDENCY latency on the accumulator?  One issuer warp, static operands in shared memory, no TMA, one
// commit at the end; consecutive MMAs go round-robin to `nacc` different TMEM accumulators (nacc = 1: every MMA accumulates into
// the accumulator the previous one wrote).  Also M = 64 (does a half-height MMA cost half?).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I ../../dad_3dheads_b200/csrc umma_accum_probe.cu \
//          ../../dad_3dheads_b200/csrc/tmap.cu -lcuda -o umma_accum_probe
#include <cstdio>
#include <vector>
#include "common.h"
#include "ptx.cuh"

using namespace dad3d;

__global__ void __launch_bounds__(128, 1) accum_kernel(int M, int N, int nacc, int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_static = smem;                    // 16 KB
  uint8_t* b_static = smem + 16384;            // 32 KB
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 4);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&done[0], 1);
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  if (warp == 0) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_f16(0, static_cast<uint32_t>(M), static_cast<uint32_t>(N));
    const uint64_t adesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(a_static));
    const uint64_t bdesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(b_static));
    const long long t0 = clock64();
    if (ptx::elect_one_sync()) {
      int a = 0;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ptx::umma_f16(tmem_base + static_cast<uint32_t>(a * N), adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
          if (++a == nacc) a = 0;
        }
      }
      ptx::umma_commit(&done[0]);
    }
    __syncwarp();
    ptx::mbar_wait(&done[0], 0);
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

int main() {
  unsigned long long* d_cyc;
  cudaMalloc(&d_cyc, 1024 * 8);
  cudaFuncSetAttribute(accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("grid | M | N | accumulators | clk per MMA | floor M/128*N/2 | tensor %%\n");
  const int iters = 4096;
  for (int grid : {1, 148})
    for (int M : {128, 64})
      for (int N : {64, 96, 128, 192, 256})
        for (int nacc : {1, 2, 3, 4}) {
          if (nacc * N > 512) continue;
          cudaMemset(d_cyc, 0, 1024 * 8);
          for (int rep = 0; rep < 2; ++rep) accum_kernel<<<grid, 128, 16384 + 32768 + 2048>>>(M, N, nacc, iters, d_cyc);
          cudaError_t le = cudaGetLastError(), se = cudaDeviceSynchronize();
          if (le != cudaSuccess || se != cudaSuccess) { printf("M %d N %d nacc %d failed: %s / %s\n", M, N, nacc, cudaGetErrorString(le), cudaGetErrorString(se)); return 2; }
          std::vector<unsigned long long> cyc(grid);
          cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
          double mean = 0;
          for (auto c : cyc) mean += static_cast<double>(c);
          mean /= grid;
          const double per = mean / (iters * 4.0), floor_clk = M / 128.0 * N / 2.0;
          printf("%4d | %3d | %3d | %d | %7.1f | %6.1f | %5.1f\n", grid, M, N, nacc, per, floor_clk, 100.0 * floor_clk / per);
        }
  return 0;
}
