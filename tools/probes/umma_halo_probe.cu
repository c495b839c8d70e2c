// Hardware probe (not part of the product): can one shared-memory HALO tile written by a single TMA box
// (SWIZZLE_128B, rows = pixels of a (tw+2) x (th+2) patch, 128 B each) feed the nine taps of a 3x3 convolution through
// K-major UMMA descriptors that only differ in their start address (+ (r*PW + s) rows) with a stride-byte-offset of
// PW rows (= 1280 B for 8-pixel-wide tiles), i.e. a start that is NOT 1024-byte aligned and an SBO that is not a
// multiple of 1024?  Prints, per tap and per base_offset convention, whether D = A * I reproduces the expected rows.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I ../../dad_3dheads_b200/csrc umma_halo_probe.cu \
//          ../../dad_3dheads_b200/csrc/api.cu -o umma_halo_probe
#include <cstdio>
#include <vector>
#include <cuda_fp16.h>
#include "common.h"
#include "ptx.cuh"
#include "tmap.h"

using namespace dad3d;

constexpr int TW = 8, TH = 16, PW = TW + 2, PH = TH + 2, ROWS = PW * PH;   // 180 halo rows of 128 B

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7u) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// mode: 0 base_offset = 0, 1 base_offset = (start >> 7) & 7
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int r, int s, int mode,
             float* out /*[128][64]*/) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;                       // halo tile: 180 rows x 128 B = 23040 B (room: 24 KiB)
  uint8_t* sb = smem + 24 * 1024;           // B: 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32 * 1024);
  uint64_t* mma_bar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar, ROWS * 128 + 64 * 128);
    ptx::tma_load_4d(sa, &map_a, bar, 0, 0, 0, 0);
    ptx::tma_load_2d(sb, &map_b, bar, 0, 0);
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();
    const uint32_t a_addr = ptx::smem_u32(sa) + static_cast<uint32_t>((r * PW + s) * 128);
    const uint64_t adesc = make_desc(a_addr, PW * 128, mode ? ((a_addr >> 7) & 7u) : 0u);
    const uint64_t bdesc = make_desc(ptx::smem_u32(sb), 1024, 0);
    const uint32_t idesc = ptx::make_idesc_f16(0, 128, 64);
    for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, adesc + 2u * k, bdesc + 2u * k, idesc, k > 0);
    ptx::umma_commit(mma_bar);
  }
  ptx::mbar_wait(mma_bar, 0);
  ptx::tc_fence_after();
  float v[32];
  for (int c = 0; c < 64; c += 32) {
    ptx::tmem_ld_32x32b_x32_f(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = v[j];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

int main() {
  // halo source [PH][PW][64] fp16; two value patterns: row id (<= 179, exact in fp16) and channel id
  std::vector<__half> hrow(ROWS * 64), hch(ROWS * 64), hb(64 * 64);
  for (int p = 0; p < ROWS; ++p)
    for (int c = 0; c < 64; ++c) {
      hrow[p * 64 + c] = __float2half(static_cast<float>(p));
      hch[p * 64 + c] = __float2half(static_cast<float>(c));
    }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *d_row, *d_ch, *d_b;
  float* d_out;
  cudaMalloc(&d_row, hrow.size() * 2); cudaMalloc(&d_ch, hch.size() * 2); cudaMalloc(&d_b, hb.size() * 2);
  cudaMalloc(&d_out, 128 * 64 * 4);
  cudaMemcpy(d_row, hrow.data(), hrow.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_ch, hch.data(), hch.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma_row, ma_ch, mb;
  const uint64_t dims[4] = {64, PW, PH, 1};
  const uint64_t str[3] = {128, PW * 128, static_cast<uint64_t>(PW) * PH * 128};
  const uint32_t box[4] = {64, PW, PH, 1};
  const uint64_t bd[2] = {64, 64};
  const uint64_t bs[1] = {128};
  const uint32_t bb[2] = {64, 64};
  if (!make_tmap_16bit(&ma_row, d_row, 4, dims, str, box, nullptr) || !make_tmap_16bit(&ma_ch, d_ch, 4, dims, str, box, nullptr) ||
      !make_tmap_16bit(&mb, d_b, 2, bd, bs, bb, nullptr)) {
    printf("tensor map creation failed\n");
    return 1;
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  std::vector<float> out(128 * 64);
  int ok_total[2] = {0, 0};
  for (int mode = 0; mode < 2; ++mode)
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        int bad_row = 0, bad_ch = 0;
        for (int pat = 0; pat < 2; ++pat) {
          probe_kernel<<<1, 128, 40 * 1024>>>(pat ? ma_ch : ma_row, mb, r, s, mode, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
          cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
              const int y = m / TW, x = m % TW;
              const float want = pat ? static_cast<float>(n) : static_cast<float>((y + r) * PW + x + s);
              if (out[m * 64 + n] != want) (pat ? bad_ch : bad_row)++;
            }
        }
        printf("mode %d (base_offset %s) tap (%d,%d): wrong rows %d, wrong channels %d  -> %s\n", mode,
               mode ? "= start>>7 & 7" : "= 0", r, s, bad_row, bad_ch, (bad_row == 0 && bad_ch == 0) ? "OK" : "MISMATCH");
        if (bad_row == 0 && bad_ch == 0) ok_total[mode]++;
      }
  printf("SUMMARY: base_offset=0: %d/9 taps exact; base_offset=start>>7&7: %d/9 taps exact\n", ok_total[0], ok_total[1]);
  return 0;
}
