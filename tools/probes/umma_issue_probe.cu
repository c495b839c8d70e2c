// Hardware probe (not part of the product), follow-up of umma_rate_probe: WHERE does the fixed ~370-cycle cost per k-block
// of the main loop (4 x UTCHMMA -> UTCBAR -> wait next full barrier) come from?  One MMA-issuer warp, N in {128,192,256},
// one product per k-block (4 MMAs), 128 x N x 16 fp16 MMAs, operands in shared memory.  Variants:
//   0  static operands, no TMA, ONE commit at the very end                      -> pure issue / execute floor
//   1  static operands, no TMA, tcgen05.commit to a dummy mbarrier after every k-block   (is the commit a pipeline bubble?)
//   2  static operands, no TMA, commit after every k-block AND the issuer waits for the commit of k-block i-2 before k-block i
//      (a 2-deep software window: is it the commit->mbarrier->try_wait round trip?)
//   3  B streamed by TMA (ring of S slots, one k-block per slot), issuer waits full[s], commit(empty[s]) per k-block (= the
//      engine's main loop with a resident A)
//   4  as 3 but slots of TWO k-blocks (two TMA boxes, one full barrier, 8 MMAs, one commit per slot)
//   5  as 3 but the issuer never waits on full[] (data races ignored): only the commit/TMA traffic remains
#include <cstdio>
#include <vector>
#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

using namespace dad3d;

__global__ void __launch_bounds__(128, 1)
issue_kernel(const __grid_constant__ CUtensorMap map_b, int N, int variant, int stages, int iters, int src_rows,
             unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int kb_per_slot = variant == 4 ? 2 : 1;
  const int slot_bytes = N * 128 * kb_per_slot;
  uint8_t* a_static = smem;                    // 16 KB
  uint8_t* b_static = smem + 16384;            // 32 KB
  uint8_t* ring = smem + 16384 + 32768;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + stages * slot_bytes);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;                  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 4);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 4; ++s) ptx::mbar_init(&done[s], 1);
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  if (warp == 0) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool use_tma = variant >= 3;
  if (warp == 0 && use_tma) {
    int st = 0; uint32_t ph = 0;
    int row0 = static_cast<int>((static_cast<long long>(blockIdx.x) * 4099 * 256) % src_rows);
    const int n_slots = iters / kb_per_slot;
    for (int it = 0; it < n_slots; ++it) {
      ptx::mbar_wait(&empty[st], ph ^ 1u);
      if (ptx::elect_one_sync()) {
        ptx::mbar_expect_tx(&full[st], static_cast<uint32_t>(slot_bytes));
        for (int j = 0; j < kb_per_slot; ++j) {
          ptx::tma_load_2d(ring + st * slot_bytes + j * N * 128, &map_b, &full[st], 0, row0);
          row0 += 256;
          if (row0 + 256 > src_rows) row0 = 0;
        }
      }
      __syncwarp();
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_f16(0, 128, static_cast<uint32_t>(N));
    int st = 0; uint32_t ph = 0;
    uint32_t dph[4] = {0, 0, 0, 0};
    const long long t0 = clock64();
    const int n_slots = iters / kb_per_slot;
    for (int it = 0; it < n_slots; ++it) {
      uint32_t sb = ptx::smem_u32(b_static);
      if (variant == 3 || variant == 4) {
        ptx::mbar_wait(&full[st], ph);
        ptx::tc_fence_after();
        sb = ptx::smem_u32(ring + st * slot_bytes);
      }
      if (variant == 2 && it >= 2) {                       // 2-deep window on the commits
        ptx::mbar_wait(&done[it & 1], dph[it & 1]);
        dph[it & 1] ^= 1u;
        ptx::tc_fence_after();
      }
      const uint32_t d = tmem_base + static_cast<uint32_t>((it & 1) * 256);
      const uint64_t adesc = ptx::make_kmajor_sw128_desc(ptx::smem_u32(a_static));
      if (ptx::elect_one_sync()) {
        for (int j = 0; j < kb_per_slot; ++j) {
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sb + j * N * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(d, adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
        }
        if (variant == 1) ptx::umma_commit(&done[2]);      // nobody waits on it (phase wraps freely)
        if (variant == 2) ptx::umma_commit(&done[it & 1]);
        if (variant >= 3) ptx::umma_commit(&empty[st]);
      }
      __syncwarp();
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
    if (ptx::elect_one_sync()) ptx::umma_commit(&done[3]);
    __syncwarp();
    ptx::mbar_wait(&done[3], 0);
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int src_rows = 1 << 16;
  uint16_t* d_src;
  cudaMalloc(&d_src, static_cast<size_t>(src_rows) * 128);
  cudaMemset(d_src, 0, static_cast<size_t>(src_rows) * 128);
  unsigned long long* d_cyc;
  cudaMalloc(&d_cyc, 1024 * 8);
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("grid | N | variant | slots | clk per k-block (4 MMAs) | floor 2N | tensor %%\n");
  const int iters = 4096;
  for (int grid : {1, sms})
    for (int N : {128, 192, 256})
      for (int variant : {0, 1, 2, 3, 4, 5}) {
        const int kbs = variant == 4 ? 2 : 1;
        const int slot_bytes = N * 128 * kbs;
        int stages = (227 * 1024 - 16384 - 32768 - 2048) / slot_bytes;
        if (stages > 6) stages = 6;
        CUtensorMap map_b;
        const uint64_t dims[2] = {64, static_cast<uint64_t>(src_rows)};
        const uint64_t str[1] = {128};
        const uint32_t box_b[2] = {64, static_cast<uint32_t>(N)};
        if (!make_tmap_16bit(&map_b, d_src, 2, dims, str, box_b, nullptr)) return 1;
        const int smem = 16384 + 32768 + stages * slot_bytes + 1024 + 512;
        cudaMemset(d_cyc, 0, 1024 * 8);
        for (int rep = 0; rep < 2; ++rep)
          issue_kernel<<<grid, 128, smem>>>(map_b, N, variant, stages, iters, src_rows, d_cyc);
        cudaError_t le = cudaGetLastError();
        cudaError_t se = cudaDeviceSynchronize();
        if (le != cudaSuccess || se != cudaSuccess) { printf("variant %d N %d failed: %s / %s\n", variant, N, cudaGetErrorString(le), cudaGetErrorString(se)); return 2; }
        std::vector<unsigned long long> cyc(grid);
        cudaMemcpy(cyc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int i = 0; i < grid; ++i) if (cyc[i] > worst) worst = cyc[i];
        const double per_kb = worst / iters;
        printf("%4d | %3d | %d | %d | %7.1f | %5.1f | %5.1f\n", grid, N, variant, stages, per_kb, 2.0 * N, 100.0 * 2.0 * N / per_kb);
      }
  return 0;
}
