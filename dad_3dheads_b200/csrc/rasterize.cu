// Z-buffer rasteriser and vertex normals on the GPU: the reference's Sim3DR native component
// (Sim3DR/lib/rasterize_kernel.cpp:158-292 `_rasterize`, :158-236 `_get_normal`; callers inference/pncc_estimator.py:16-43,
// Sim3DR/Sim3DR.py:8-29).  BIT-EXACT with the reference's sequential C++ loop:
//   * every floating-point operation is issued in the reference's order with round-to-nearest intrinsics (no FMA contraction:
//     the reference's build has none), so barycentric weights, depths and colours are the same bits;
//   * the sequential z-test "if (p_depth > depth_buffer[pix])" over triangles in index order = per pixel the maximum depth, ties
//     won by the LOWEST triangle index: one 64-bit atomicMax on (orderable(depth) << 32 | ~index) per covered pixel, then a
//     per-pixel resolve pass that shades the winner.  alpha must be 1 (what Sim3DR.rasterize passes; blending with alpha < 1 is
//     order-dependent).
//   * vertex normals accumulate the un-normalised face normals of the incident triangles in ascending triangle order (caller-built
//     CSR adjacency), the order of the reference's scatter loop.
// Integer / float HBM-bound work; no tensor cores.
#include <cstdint>

#include "../../include/dad3d.h"
#include "common.h"

namespace dad3d {

__device__ __forceinline__ float rmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float radd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float rsub(float a, float b) { return __fsub_rn(a, b); }

// get_point_weight (rasterize_kernel.cpp:52-82), operation for operation
__device__ __forceinline__ void point_weight(float* w, float px, float py, float p0x, float p0y, float p1x, float p1y, float p2x,
                                             float p2y) {
  const float v0x = rsub(p2x, p0x), v0y = rsub(p2y, p0y);
  const float v1x = rsub(p1x, p0x), v1y = rsub(p1y, p0y);
  const float v2x = rsub(px, p0x), v2y = rsub(py, p0y);
  const float dot00 = radd(rmul(v0x, v0x), rmul(v0y, v0y));
  const float dot01 = radd(rmul(v0x, v1x), rmul(v0y, v1y));
  const float dot02 = radd(rmul(v0x, v2x), rmul(v0y, v2y));
  const float dot11 = radd(rmul(v1x, v1x), rmul(v1y, v1y));
  const float dot12 = radd(rmul(v1x, v2x), rmul(v1y, v2y));
  const float den = rsub(rmul(dot00, dot11), rmul(dot01, dot01));
  const float inv = (den == 0.f) ? 0.f : __fdiv_rn(1.0f, den);
  const float u = rmul(rsub(rmul(dot11, dot02), rmul(dot01, dot12)), inv);
  const float v = rmul(rsub(rmul(dot00, dot12), rmul(dot01, dot02)), inv);
  w[0] = rsub(rsub(1.0f, u), v);
  w[1] = v;
  w[2] = u;
}

__device__ __forceinline__ unsigned int orderable(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct TriSetup {
  float p0x, p0y, d0, p1x, p1y, d1, p2x, p2y, d2;
  int i0, i1, i2;
};
__device__ __forceinline__ TriSetup load_tri(const float* __restrict__ v, const int* __restrict__ t, int i) {
  TriSetup s;
  s.i0 = t[3 * i]; s.i1 = t[3 * i + 1]; s.i2 = t[3 * i + 2];
  s.p0x = v[3 * s.i0]; s.p0y = v[3 * s.i0 + 1]; s.d0 = v[3 * s.i0 + 2];
  s.p1x = v[3 * s.i1]; s.p1y = v[3 * s.i1 + 1]; s.d1 = v[3 * s.i1 + 2];
  s.p2x = v[3 * s.i2]; s.p2y = v[3 * s.i2 + 1]; s.d2 = v[3 * s.i2 + 2];
  return s;
}

// pass 1: one thread per triangle walks its bounding box
__global__ void raster_tri_kernel(const float* __restrict__ verts, const int* __restrict__ tris, int ntri, int h, int w,
                                  const float* __restrict__ depth_in, unsigned long long* __restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntri) return;
  const TriSetup s = load_tri(verts, tris, i);
  const int x_min = max(static_cast<int>(ceilf(fminf(s.p0x, fminf(s.p1x, s.p2x)))), 0);
  const int x_max = min(static_cast<int>(floorf(fmaxf(s.p0x, fmaxf(s.p1x, s.p2x)))), w - 1);
  const int y_min = max(static_cast<int>(ceilf(fminf(s.p0y, fminf(s.p1y, s.p2y)))), 0);
  const int y_max = min(static_cast<int>(floorf(fmaxf(s.p0y, fmaxf(s.p1y, s.p2y)))), h - 1);
  if (x_max < x_min || y_max < y_min) return;
  for (int y = y_min; y <= y_max; ++y)
    for (int x = x_min; x <= x_max; ++x) {
      float wt[3];
      point_weight(wt, static_cast<float>(x), static_cast<float>(y), s.p0x, s.p0y, s.p1x, s.p1y, s.p2x, s.p2y);
      if (wt[2] > 0 && wt[1] > 0 && wt[0] > 0) {
        const float d = radd(radd(rmul(wt[0], s.d0), rmul(wt[1], s.d1)), rmul(wt[2], s.d2));
        if (d > depth_in[y * w + x]) {
          const unsigned long long k = (static_cast<unsigned long long>(orderable(d)) << 32) |
                                       static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned int>(i));
          atomicMax(&key[y * w + x], k);
        }
      }
    }
}

// pass 2: one thread per pixel shades the winning triangle
__global__ void raster_resolve_kernel(const float* __restrict__ verts, const int* __restrict__ tris, const float* __restrict__ colors,
                                      int h, int w, int c, int reverse, const unsigned long long* __restrict__ key,
                                      unsigned char* __restrict__ image, float* __restrict__ depth) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= h * w) return;
  const unsigned long long k = key[pix];
  if (k == 0ull) return;
  const int i = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned int>(k & 0xFFFFFFFFull));
  const int y = pix / w, x = pix - y * w;
  const TriSetup s = load_tri(verts, tris, i);
  float wt[3];
  point_weight(wt, static_cast<float>(x), static_cast<float>(y), s.p0x, s.p0y, s.p1x, s.p1y, s.p2x, s.p2y);
  depth[pix] = radd(radd(rmul(wt[0], s.d0), rmul(wt[1], s.d1)), rmul(wt[2], s.d2));
  const int yo = reverse ? (h - 1 - y) : y;
  for (int ch = 0; ch < c; ++ch) {
    const float pc = radd(radd(rmul(wt[0], colors[c * s.i0 + ch]), rmul(wt[1], colors[c * s.i1 + ch])), rmul(wt[2], colors[c * s.i2 + ch]));
    // (unsigned char)((1 - alpha) * image + alpha * 255 * p_color) with alpha = 1
    const float old = static_cast<float>(image[(yo * w + x) * c + ch]);
    const float val = radd(rmul(rsub(1.0f, 1.0f), old), rmul(rmul(1.0f, 255.0f), pc));
    image[(yo * w + x) * c + ch] = static_cast<unsigned char>(static_cast<int>(val));
  }
}

// one thread per vertex: un-normalised face normals of its incident triangles, added in ascending triangle order
__global__ void vertex_normal_kernel(const float* __restrict__ v, const int* __restrict__ t, const int* __restrict__ adj_off,
                                     const int* __restrict__ adj_tri, int nver, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nver) return;
  float nx = 0.f, ny = 0.f, nz = 0.f;
  for (int e = adj_off[i]; e < adj_off[i + 1]; ++e) {
    const int tr = adj_tri[e];
    const int a = t[3 * tr], b = t[3 * tr + 1], c = t[3 * tr + 2];
    const float v1x = rsub(v[3 * b], v[3 * a]), v1y = rsub(v[3 * b + 1], v[3 * a + 1]), v1z = rsub(v[3 * b + 2], v[3 * a + 2]);
    const float v2x = rsub(v[3 * c], v[3 * a]), v2y = rsub(v[3 * c + 1], v[3 * a + 1]), v2z = rsub(v[3 * c + 2], v[3 * a + 2]);
    nx = radd(nx, rsub(rmul(v1y, v2z), rmul(v1z, v2y)));
    ny = radd(ny, rsub(rmul(v1z, v2x), rmul(v1x, v2z)));
    nz = radd(nz, rsub(rmul(v1x, v2y), rmul(v1y, v2x)));
  }
  float det = __fsqrt_rn(radd(radd(rmul(nx, nx), rmul(ny, ny)), rmul(nz, nz)));
  if (det <= 0) det = 1e-6f;
  out[3 * i] = __fdiv_rn(nx, det);
  out[3 * i + 1] = __fdiv_rn(ny, det);
  out[3 * i + 2] = __fdiv_rn(nz, det);
}

}  // namespace dad3d

using namespace dad3d;

extern "C" {

int dad3d_rasterize(const float* vertices_d, const int32_t* triangles_d, const float* colors_d, int32_t ntri, uint8_t* image_d,
                    float* depth_d, unsigned long long* key_ws_d, int32_t h, int32_t w, int32_t c, int32_t reverse,
                    dad3d_stream stream_) {
  DAD3D_REQUIRE(vertices_d && triangles_d && colors_d && image_d && depth_d && key_ws_d, "null pointer");
  DAD3D_REQUIRE(ntri >= 0 && h > 0 && w > 0 && c > 0 && static_cast<long long>(h) * w < (1ll << 31), "shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DAD3D_CUDA_OK(cudaMemsetAsync(key_ws_d, 0, sizeof(unsigned long long) * static_cast<size_t>(h) * w, stream));
  if (ntri == 0) return DAD3D_OK;
  raster_tri_kernel<<<(ntri + 127) / 128, 128, 0, stream>>>(vertices_d, triangles_d, ntri, h, w, depth_d, key_ws_d);
  count_launch();
  DAD3D_CUDA_OK(cudaGetLastError());
  raster_resolve_kernel<<<(h * w + 255) / 256, 256, 0, stream>>>(vertices_d, triangles_d, colors_d, h, w, c, reverse, key_ws_d,
                                                                 image_d, depth_d);
  count_launch();
  DAD3D_CUDA_OK(cudaGetLastError());
  return DAD3D_OK;
}

int dad3d_vertex_normals(const float* vertices_d, const int32_t* triangles_d, const int32_t* adj_offsets_d,
                         const int32_t* adj_triangles_d, int32_t nver, float* normals_d, dad3d_stream stream_) {
  DAD3D_REQUIRE(vertices_d && triangles_d && adj_offsets_d && adj_triangles_d && normals_d, "null pointer");
  DAD3D_REQUIRE(nver >= 0, "shape");
  if (nver == 0) return DAD3D_OK;
  vertex_normal_kernel<<<(nver + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      vertices_d, triangles_d, adj_offsets_d, adj_triangles_d, nver, normals_d);
  count_launch();
  DAD3D_CUDA_OK(cudaGetLastError());
  return DAD3D_OK;
}

}  // extern "C"
