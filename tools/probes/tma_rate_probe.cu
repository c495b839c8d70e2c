// Hardware probe (not part of the product): sustained TMA load throughput of one SM / of the whole chip for the box shapes
// the tile engine uses (rows of 128 B, SWIZZLE_128B), as a function of ring depth and box height -- with NO tensor-core work,
// so the number is the operand-supply ceiling the GEMM main loops run against.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I ../../dad_3dheads_b200/csrc tma_rate_probe.cu \
//          ../../dad_3dheads_b200/csrc/api.cu -o tma_rate_probe
#include <cstdio>
#include <vector>
#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

using namespace dad3d;

// one producer lane + one consumer lane per CTA; `iters` stage fills of `boxes` TMA boxes (rows x 128 B each)
__global__ void __launch_bounds__(64, 1)
rate_kernel(const __grid_constant__ CUtensorMap map, int rows, int boxes, int stages, int iters, int src_rows, int spread,
            unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = rows * 128 * boxes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + stages;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    int st = 0; uint32_t ph = 0;
    // spread = 1: every CTA walks its own region of the source; 0: all CTAs read the same rows (hot in L2, like weights)
    int row0 = spread ? static_cast<int>((static_cast<long long>(blockIdx.x) * 4099 * rows) % src_rows) : 0;
    for (int it = 0; it < iters; ++it) {
      ptx::mbar_wait(&empty[st], ph ^ 1u);
      if (ptx::elect_one_sync()) {
        ptx::mbar_expect_tx(&full[st], static_cast<uint32_t>(stage_bytes));
        for (int b = 0; b < boxes; ++b) {
          ptx::tma_load_2d(smem + st * stage_bytes + b * rows * 128, &map, &full[st], 0, row0);
          row0 += rows;
          if (row0 + rows > src_rows) row0 = 0;
        }
      }
      __syncwarp();
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
  } else {
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      ptx::mbar_wait(&full[st], ph);
      if (ptx::elect_one_sync()) ptx::mbar_arrive(&empty[st]);
      __syncwarp();
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int src_rows = 1 << 16;                       // 65536 rows x 128 B = 8 MiB (L2 resident after the first pass)
  uint16_t* d_src;
  cudaMalloc(&d_src, static_cast<size_t>(src_rows) * 128);
  cudaMemset(d_src, 0, static_cast<size_t>(src_rows) * 128);
  unsigned long long* d_cyc;
  cudaMalloc(&d_cyc, 1024 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("grid | box rows | boxes/stage | stage KB | stages | source | B/clk/SM | chip TB/s @1.965GHz\n");
  const int grids[2] = {1, sms};
  for (int gi = 0; gi < 2; ++gi)
    for (int rows : {64, 128, 256})
      for (int boxes : {1, 2, 4})
        for (int stages : {2, 3, 4, 6}) {
          const int stage_bytes = rows * 128 * boxes;
          if (stages * stage_bytes > 200 * 1024) continue;
          for (int spread = 0; spread < 2; ++spread) {
            CUtensorMap map;
            const uint64_t dims[2] = {64, static_cast<uint64_t>(src_rows)};
            const uint64_t str[1] = {128};
            const uint32_t box[2] = {64, static_cast<uint32_t>(rows)};
            if (!make_tmap_16bit(&map, d_src, 2, dims, str, box, nullptr)) { printf("tmap failed\n"); return 1; }
            const int iters = 2000;
            const int smem = stages * stage_bytes + 1024 + 256;
            for (int rep = 0; rep < 2; ++rep)      // first pass warms L2
              rate_kernel<<<grids[gi], 64, smem>>>(map, rows, boxes, stages, iters, src_rows, spread, d_cyc);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
            std::vector<unsigned long long> cyc(grids[gi]);
            cudaMemcpy(cyc.data(), d_cyc, grids[gi] * 8, cudaMemcpyDeviceToHost);
            double worst = 0;
            for (auto c : cyc) worst = c > worst ? c : worst;
            const double bpc = static_cast<double>(iters) * stage_bytes / worst;
            printf("%4d | %3d | %d | %5.1f | %d | %s | %6.1f | %5.2f\n", grids[gi], rows, boxes, stage_bytes / 1024.0, stages,
                   spread ? "spread" : "shared", bpc, bpc * grids[gi] * 1.965e9 / 1e12);
          }
        }
  return 0;
}
