// Hardware probe (not part of the product): what limits tcgen05.mma issue on one SM when shared memory is shared with
// TMA writes and epilogue staging traffic?  Measures cycles per 128 x N x 16 (fp16) MMA for
//   mode 0  operands static in shared memory (no TMA, no other traffic)            -> the tensor-pipe floor N/2
//   mode 1  A static, B streamed through a TMA ring with the real full/empty dependency ("A-stationary" main loop)
//   mode 2  A and B both streamed (what the tile engine's main loop does)
//   +8      eight extra warps hammer shared memory with STS.128 / LDS.128 (epilogue staging stand-in)
// with n_prod products per k-block (1 = one product; 3 = hi/lo split: the same A/B stage is read by 3 MMAs groups).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I ../../dad_3dheads_b200/csrc umma_rate_probe.cu \
//          ../../dad_3dheads_b200/csrc/api.cu -lcuda -o umma_rate_probe
#include <cstdio>
#include <vector>
#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

using namespace dad3d;

constexpr int kThreads = 320;

__global__ void __launch_bounds__(kThreads, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int N, int mode, int n_prod,
             int stages, int iters, int src_rows, unsigned long long* cycles, unsigned long long* epi_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool stream_b = (mode & 3) >= 1, stream_a = (mode & 3) >= 2, epi_traffic = (mode & 8) != 0;
  const int a_bytes = 128 * 128, b_bytes = N * 128;
  const int stage_bytes = (stream_a ? a_bytes : 0) + (stream_b ? b_bytes : 0);
  uint8_t* a_static = smem;                                  // 16 KB
  uint8_t* b_static = smem + a_bytes;                        // up to 32 KB
  uint8_t* ring = smem + a_bytes + 256 * 128;
  uint8_t* epi_buf = ring + stages * (stage_bytes > 0 ? stage_bytes : 0);     // 8 x 4 KB
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_buf + 8 * 4096);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;                                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);
  volatile uint32_t* stop = tmem_slot + 1;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(&done[0], 1);
    ptx::mbar_init(&done[1], 1);
    *stop = 0;
    ptx::fence_mbar_init();
  }
  // zero the static operands (values are irrelevant for timing; denormal/NaN patterns are avoided)
  for (int i = threadIdx.x; i < (a_bytes + 256 * 128) / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  if (warp == 0) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t0 = clock64();
  if (warp == 0) {
    if (stage_bytes > 0) {
      int st = 0; uint32_t ph = 0;
      int row0 = static_cast<int>((static_cast<long long>(blockIdx.x) * 4099 * 256) % src_rows);
      for (int it = 0; it < iters; ++it) {
        ptx::mbar_wait(&empty[st], ph ^ 1u);
        if (ptx::elect_one_sync()) {
          ptx::mbar_expect_tx(&full[st], static_cast<uint32_t>(stage_bytes));
          uint8_t* dst = ring + st * stage_bytes;
          if (stream_a) { ptx::tma_load_2d(dst, &map_a, &full[st], 0, row0); dst += a_bytes; }
          if (stream_b) ptx::tma_load_2d(dst, &map_b, &full[st], 0, row0);
          row0 += 256;
          if (row0 + 256 > src_rows) row0 = 0;
        }
        __syncwarp();
        if (++st == stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_f16(0, 128, static_cast<uint32_t>(N));
    int st = 0; uint32_t ph = 0;
    uint32_t dph[2] = {0, 0};
    for (int it = 0; it < iters; ++it) {
      uint32_t sa = ptx::smem_u32(a_static), sb = ptx::smem_u32(b_static);
      if (stage_bytes > 0) {
        ptx::mbar_wait(&full[st], ph);
        ptx::tc_fence_after();
        uint32_t p = ptx::smem_u32(ring + st * stage_bytes);
        if (stream_a) { sa = p; p += a_bytes; }
        if (stream_b) sb = p;
      }
      const uint32_t d = tmem_base + static_cast<uint32_t>((it & 1) * 256);
      const uint64_t adesc = ptx::make_kmajor_sw128_desc(sa), bdesc = ptx::make_kmajor_sw128_desc(sb);
      if (ptx::elect_one_sync()) {
        for (int p = 0; p < n_prod; ++p) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(d, adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
        }
        if (stage_bytes > 0) ptx::umma_commit(&empty[st]);
        if ((it & 15) == 15) ptx::umma_commit(&done[(it >> 4) & 1]);
      }
      __syncwarp();
      if ((it & 15) == 15 && it >= 31) {                    // keep at most ~32 k-blocks of MMAs in flight
        const int b = ((it >> 4) - 1) & 1;
        ptx::mbar_wait(&done[b], dph[b]);
        dph[b] ^= 1u;
      }
      if (stage_bytes > 0 && ++st == stages) { st = 0; ph ^= 1u; }
    }
    // drain
    if (ptx::elect_one_sync()) ptx::umma_commit(&full[7]);
    __syncwarp();
    ptx::mbar_wait(&full[7], 0);
    if (lane == 0) *stop = 1;
  } else if (epi_traffic) {
    // 8 warps: each writes its 4 KB staging tile with STS.128 (row pitch 33 x 16 B: conflict-free) and reads it back row-wise
    uint4* buf = reinterpret_cast<uint4*>(epi_buf + (warp - 2) * 4096);
    unsigned long long n = 0;
    uint4 v = make_uint4(lane, 1, 2, 3);
    while (*stop == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) buf[j * 32 + lane] = v;
      __syncwarp();
      uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const uint4 t = buf[j * 32 + ((lane + j) & 31)]; acc.x ^= t.x; acc.y += t.y; }
      v.x += acc.x; v.y ^= acc.y;
      __syncwarp();
      n += 2 * 8 * 512;
    }
    if (lane == 0) atomicAdd(epi_bytes + blockIdx.x, n);
    if (v.x == 0x12345678u) printf("%u", v.y);
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int src_rows = 1 << 16;                       // 8 MiB source: L2 resident
  uint16_t* d_src;
  cudaMalloc(&d_src, static_cast<size_t>(src_rows) * 128);
  cudaMemset(d_src, 0, static_cast<size_t>(src_rows) * 128);
  unsigned long long *d_cyc, *d_epi;
  cudaMalloc(&d_cyc, 1024 * 8);
  cudaMalloc(&d_epi, 1024 * 8);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("grid | N | mode | products | stages | clk/MMA | floor N/2 | tensor %% | TMA B/clk | epi smem B/clk | total smem B/clk (model)\n");
  const int iters = 4096;
  for (int grid : {1, sms})
    for (int N : {64, 96, 128, 192, 256})
      for (int n_prod : {1, 3})
        for (int mode : {0, 1, 2, 8, 9, 10}) {
          if (grid == 1 && (mode & 8)) continue;
          const bool sb = (mode & 3) >= 1, sa = (mode & 3) >= 2;
          const int stage_bytes = (sa ? 16384 : 0) + (sb ? N * 128 : 0);
          int stages = stage_bytes ? (227 * 1024 - 16384 - 32768 - 32768 - 2048) / stage_bytes : 1;
          if (stages > 6) stages = 6;
          CUtensorMap map_a, map_b;
          const uint64_t dims[2] = {64, static_cast<uint64_t>(src_rows)};
          const uint64_t str[1] = {128};
          const uint32_t box_a[2] = {64, 128}, box_b[2] = {64, static_cast<uint32_t>(N)};
          if (!make_tmap_16bit(&map_a, d_src, 2, dims, str, box_a, nullptr)) return 1;
          if (!make_tmap_16bit(&map_b, d_src, 2, dims, str, box_b, nullptr)) return 1;
          const int smem = 16384 + 32768 + stages * stage_bytes + 32768 + 1024 + 512;
          cudaMemset(d_epi, 0, 1024 * 8);
          for (int rep = 0; rep < 2; ++rep) {
            if (rep == 1) cudaMemset(d_epi, 0, 1024 * 8);
            probe_kernel<<<grid, kThreads, smem>>>(map_a, map_b, N, mode, n_prod, stages, iters, src_rows, d_cyc, d_epi);
          }
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
          std::vector<unsigned long long> cyc(grid), eb(grid);
          cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
          cudaMemcpy(eb.data(), d_epi, grid * 8, cudaMemcpyDeviceToHost);
          double worst = 0, epi = 0;
          for (int i = 0; i < grid; ++i) { if (cyc[i] > worst) worst = cyc[i]; epi += eb[i]; }
          epi /= grid;
          const double n_mma = static_cast<double>(iters) * 4 * n_prod;
          const double cpm = worst / n_mma;
          const double tma_bpc = static_cast<double>(iters) * stage_bytes / worst;
          const double mma_read = (4096.0 + N * 32.0) / cpm;
          printf("%4d | %3d | %2d | %d | %d | %6.1f | %5.1f | %5.1f | %6.1f | %6.1f | %6.1f\n", grid, N, mode, n_prod, stages, cpm,
                 N / 2.0, 100.0 * (N / 2.0) / cpm, tma_bpc, epi / worst, mma_read + tma_bpc + epi / worst);
        }
  return 0;
}
