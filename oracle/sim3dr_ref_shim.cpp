// extern "C" entry points over the reference's UNMODIFIED rasteriser (Sim3DR/lib/rasterize_kernel.cpp), so that the tests can
// call it through ctypes without building the reference's Cython module.  TEST INFRASTRUCTURE: compiled by oracle/build_ref.py
// together with the reference source file, where it lies, into oracle/_ref/libsim3dr_ref.so.
#include "rasterize.h"

extern "C" {

void sim3dr_ref_rasterize(unsigned char* image, float* vertices, int* triangles, float* colors, float* depth_buffer, int ntri,
                          int h, int w, int c, float alpha, int reverse) {
  _rasterize(image, vertices, triangles, colors, depth_buffer, ntri, h, w, c, alpha, reverse != 0);
}

void sim3dr_ref_get_normal(float* ver_normal, float* vertices, int* triangles, int nver, int ntri) {
  _get_normal(ver_normal, vertices, triangles, nver, ntri);
}

}
