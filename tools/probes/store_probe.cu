// Store-path probe for the decode epilogue: how fast can 148 persistent CTAs (8 storing warps each, like the decode kernel's
// epilogue) write a [rows][15069] fp32 matrix (row pitch 60 276 B, rows only 4-byte aligned) for different request shapes?
//   mode 0: reference -- every warp instruction writes 1 KiB contiguous (STG.256 x 32 lanes), rows treated as one flat array
//   mode 1: one lane = one row, STG.256: 32 single-sector requests in 32 different rows per instruction (the decode kernel's shape)
//   mode 2: quads: 4 lanes write the 4 sectors of one 128-byte line, 8 rows per instruction
//   mode 3: 8 lanes write 2 lines (256 B) of a row, 4 rows per instruction
//   mode 4: like 1 but STG.128 pairs (two half-sector stores)
//   mode 5: one lane = one row, scalar STG.32 x 8 per sector (what partial-sector writes cost)
// Rows are 32-byte aligned here by rounding addresses down (data are garbage; only the request shapes matter).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_probe store_probe.cu && ./store_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void st8(float* p, float v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st4(float* p, float v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(256, 1) probe(float* out, int rows, int pitch, int mode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  const float v = static_cast<float>(gw);
  const int sectors_per_row = pitch / 8;          // whole sectors only
  if (mode == 0) {
    const size_t total = static_cast<size_t>(rows) * pitch / 256;     // 1 KiB chunks
    for (size_t c = gw; c < total; c += nw) st8(out + c * 256 + lane * 8, v);
    return;
  }
  // a warp owns 32 rows at a time (row blocks dealt round-robin), walks the row in steps of 3 sectors like the decode passes
  for (int rb = gw; rb * 32 < rows; rb += nw) {
    float* base = out + static_cast<size_t>(rb) * 32 * pitch;
    for (int s0 = 0; s0 + 3 <= sectors_per_row; s0 += 12) {
      // 12 sectors per row per outer step = 4 "passes" of 3 sectors
      if (mode == 1 || mode == 4 || mode == 5) {
        for (int q = 0; q < 4; ++q)
          for (int k = 0; k < 3; ++k) {
            const int s = s0 + 3 * q + k;
            if (s >= sectors_per_row) continue;
            float* p = reinterpret_cast<float*>(reinterpret_cast<unsigned long long>(base + static_cast<size_t>(lane) * pitch + s * 8) & ~31ull);
            if (mode == 1) st8(p, v);
            else if (mode == 4) { st4(p, v); st4(p + 4, v); }
            else { for (int j = 0; j < 8; ++j) p[j] = v; }
          }
      } else if (mode == 2) {
        // quad i writes lines: 12 sectors x 4 rows = 12 lines per quad per outer step
        for (int L = 0; L < 12; ++L) {
          const int r = (lane & ~3) + (L & 3);
          const int s = s0 + (L >> 2) * 4 + (lane & 3);
          if (s >= sectors_per_row) continue;
          float* p = reinterpret_cast<float*>(reinterpret_cast<unsigned long long>(base + static_cast<size_t>(r) * pitch + s * 8) & ~31ull);
          st8(p, v);
        }
      } else if (mode == 3) {
        for (int L = 0; L < 12; ++L) {           // 8 lanes x 8 sectors...: rows (lane & ~7) + (L % 8), 8 sectors at s0 + 8 (L / 8) -- 12 sectors = 1.5 chunks
          const int r = (lane & ~7) + (L & 7);
          const int s = s0 + (L >> 3) * 8 + (lane & 7);
          if (s >= sectors_per_row || (L >> 3) * 8 + (lane & 7) >= 12) continue;
          float* p = reinterpret_cast<float*>(reinterpret_cast<unsigned long long>(base + static_cast<size_t>(r) * pitch + s * 8) & ~31ull);
          st8(p, v);
        }
      }
    }
  }
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? std::atoi(argv[1]) : 75776;
  const int pitch = 15069;
  float* d;
  cudaMalloc(&d, static_cast<size_t>(rows) * pitch * 4 + 4096);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 6; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      probe<<<148, 256>>>(d, rows, pitch, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      if (cudaGetLastError() != cudaSuccess) { std::printf("mode %d: launch failed\n", mode); return 1; }
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gb = static_cast<double>(rows) * pitch * 4 / 1e9;
    std::printf("mode %d: %.3f ms  %.2f TB/s  (= %.1f M heads/s of vertex writes)\n", mode, ms, gb / ms, rows / ms / 1e3);
  }
  // occupancy variant: more CTAs / warps in flight for mode 1 and 2
  for (int mode = 1; mode <= 2; ++mode) {
    cudaEventRecord(e0);
    probe<<<148 * 4, 256>>>(d, rows, pitch, mode);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::printf("mode %d, 4 CTAs/SM: %.3f ms  %.2f TB/s\n", mode, ms, static_cast<double>(rows) * pitch * 4 / 1e9 / ms);
  }
  return 0;
}
