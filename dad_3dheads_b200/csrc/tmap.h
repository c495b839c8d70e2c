// Host-side TMA tensor-map construction.  The driver entry point is fetched through the runtime
// (cudaGetDriverEntryPoint) so libdad3d.so has no link-time dependency on libcuda and loads on a GPU-less build box.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <string>

namespace dad3d {

void set_error(const std::string& msg);   // api.cu

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error(std::string("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: ") + cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 16-bit element tensor, rank r (<= 4).  dims[0] is the innermost (contiguous) extent; strides_bytes[i] is the byte
// stride of dim i+1.  box[] is in tensor elements BEFORE the traversal stride (elem_strides); swizzle_bytes (128 / 64 / 0)
// must equal box[0] * 2 bytes for the swizzled modes.  Out-of-bounds elements read as zero / are not written.
inline bool make_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                            int swizzle_bytes = 128) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) + " (rank " +
              std::to_string(rank) + ", dims " + std::to_string(dims[0]) + "," + std::to_string(rank > 1 ? dims[1] : 0) +
              "," + std::to_string(rank > 2 ? dims[2] : 0) + "," + std::to_string(rank > 3 ? dims[3] : 0) + ")");
    return false;
  }
  return true;
}

}  // namespace dad3d
