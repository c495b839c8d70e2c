// FFMA vs FFMA2 (fma.rn.f32x2, sm_100) throughput per SM: 8 warps per SM (2 per scheduler, like the decode epilogue), 16
// independent accumulator chains per thread.  Reports FMAs per clock per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, int iters, float a, float b) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, A.x, B.x); acc[i].y = fmaf(acc[i].y, A.y, B.y); }
      else acc[i] = __ffma2_rn(acc[i], A, B);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + (1 << 20))[0] = t1 - t0;
}

int main() {
  float* d;
  cudaMalloc(&d, (1 << 20) * 4 + 1024);
  long long* dt = reinterpret_cast<long long*>(d + (1 << 20));
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 256>>>(d, iters, 1.0001f, 0.5f);
      else k<1><<<148, 256>>>(d, iters, 1.0001f, 0.5f);
      cudaDeviceSynchronize();
    }
    long long cyc;
    cudaMemcpy(&cyc, dt, 8, cudaMemcpyDeviceToHost);
    const double fmas = 256.0 * 32 * iters;      // per SM
    std::printf("%s: %lld cycles, %.1f FMA/clk/SM, %.2f instr/clk/SM (warp instr)\n", mode ? "FFMA2" : "FFMA ", cyc, fmas / cyc,
                (mode ? 16.0 : 32.0) * 8 * iters / cyc);
  }
  return 0;
}
