/* libdad3d.so -- C ABI of the B200-native DAD-3DNet image->3D-head hot path.
 *
 * The reference (PinataFarms/DAD-3DHeads) has no FFI seam: its "operator interface" for this path is three Python
 * objects.  Each entry point below names the reference call it replaces (file:line under
 * /root/reference); INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions: plain C, no torch types.  Pointers suffixed _h are HOST pointers (read during the call only), _d are
 * DEVICE pointers owned by the caller.  Every compute call is asynchronous on the given CUDA stream.  Return value
 * 0 = OK, negative = error (text via dad3d_last_error(), thread-local).  Handles are opaque, created/destroyed
 * explicitly; one handle may be used by one host thread at a time.  Nothing here ever falls back to the CPU.
 */
#ifndef DAD3D_H_
#define DAD3D_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DAD3D_API __attribute__((visibility("default")))
#else
#define DAD3D_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dad3d_stream;      /* == cudaStream_t / CUstream */
typedef struct dad3d_flame dad3d_flame;
typedef struct dad3d_encoder dad3d_encoder;

#define DAD3D_OK 0
#define DAD3D_ERR_INVALID (-1)
#define DAD3D_ERR_CUDA (-2)
#define DAD3D_ERR_UNSUPPORTED (-3)

/* Widths of the fields of the 3DMM parameter vector, sliced in the reference's hard-coded order
 * shape, expression, jaw, rotation, eyeballs, neck, translation, scale
 * (model_training/model/flame.py:41-84 FlameParams.from_3dmm; dad_3dnet.yaml:5-12 gives 300/100/3/6/0/0/3/1 = 413). */
typedef struct dad3d_flame_layout {
  int32_t shape, expression, jaw, rotation, eyeballs, neck, translation, scale;
} dad3d_flame_layout;

/* decode flags */
#define DAD3D_ZERO_ROT 1        /* FLAMELayer.forward(zero_rot=True)   flame.py:225 */
#define DAD3D_ZERO_JAW 2        /* FLAMELayer.forward(zero_jaw=True)   flame.py:206 */
#define DAD3D_BLEND_FAST 4      /* blend-shape product in ONE fp16 tensor-core pass (11-bit operands, like TF32);
                                   default is the 3-product hi/lo split (fp32-class accuracy) */
#define DAD3D_BLEND_SIMT 8      /* verification aid: blend-shape product on CUDA cores in fp32 (slow) */

DAD3D_API const char* dad3d_last_error(void);
DAD3D_API int dad3d_version(void);

/* ---- FLAME head decoder -------------------------------------------------------------------------------------------
 * dad3d_flame_create  replaces FLAMELayer.__init__ (model_training/model/flame.py:124-180): takes the fp32 constants
 *   the reference registers as buffers and packs them for the GPU (fp16 hi/lo basis planes, folded joint regressor).
 *   shapedirs_h [n_vertices*3, n_betas] (i.e. [V,3,400] row-major), posedirs_h [(n_joints-1)*9, n_vertices*3],
 *   v_template_h [n_vertices*3], j_regressor_h [n_joints, n_vertices], parents_h [n_joints] (parents[0] = -1),
 *   lbs_weights_h [n_vertices, n_joints].  n_joints must be 5 (FLAME: global, neck, jaw, eye, eye), n_betas 400. */
DAD3D_API int dad3d_flame_create(dad3d_flame** out, const float* shapedirs_h, const float* posedirs_h, const float* v_template_h,
                       const float* j_regressor_h, const int32_t* parents_h, const float* lbs_weights_h,
                       int32_t n_vertices, int32_t n_betas, int32_t n_joints, const dad3d_flame_layout* layout,
                       int32_t device);
DAD3D_API void dad3d_flame_destroy(dad3d_flame* h);
DAD3D_API int32_t dad3d_flame_num_params(const dad3d_flame* h);      /* 413 for the released layout */
DAD3D_API int32_t dad3d_flame_num_vertices(const dad3d_flame* h);
/* scratch the decode needs for a batch of B heads (decode streams internally in chunks, so this saturates) */
DAD3D_API size_t dad3d_flame_workspace_bytes(const dad3d_flame* h, int32_t B);

/* dad3d_flame_decode  replaces HeadMesh.vertices_3d + HeadMesh.reprojected_vertices
 *   (model_training/head_mesh.py:28-46 -> FLAMELayer.forward flame.py:182-229 -> smplx.lbs.lbs) in ONE pass:
 *   params_d   [B, num_params] fp32 row-major (not modified -- the reference's in-place zeroing of translation z,
 *              head_mesh.py:41, is reproduced by the Python wrapper, not here)
 *   vertices3d_d  [B, V, 3] fp32 or NULL : rotated model-space mesh (vertices_3d, zero_rot per flags)
 *   projected_d   [B, V, 2] (to_2d != 0) or [B, V, 3] fp32 or NULL : ((v*max(s+1,1e-8) + [tx,ty,0]) + 1)/2*image_size
 *                 (always uses the 6-DoF rotation, as reprojected_vertices does, unless DAD3D_ZERO_ROT is set) */
DAD3D_API int dad3d_flame_decode(dad3d_flame* h, const float* params_d, int32_t B, int32_t flags, float* vertices3d_d,
                       float* projected_d, float image_size, int32_t to_2d, void* workspace_d, size_t workspace_bytes,
                       dad3d_stream stream);

/* dad3d_gather_landmarks  replaces np.take(projected_vertices, indices, axis=0) (demo_utils.py:37-47) and
 *   FLAMELayer.indices_2d style subset selection: out[b,l,:] = src[b, idx[l], :].  ncomp = 2 or 3. */
DAD3D_API int dad3d_gather_landmarks(const float* src_d, int32_t B, int32_t n_vertices, int32_t ncomp, const int32_t* idx_d,
                           int32_t L, float* out_d, dad3d_stream stream);
/* barycentric variant (model_training/data/utils.py:120-206 get_68_landmarks): out[b,l,:] = sum_k bary[l,k]*src[b,tri[l,k],:] */
DAD3D_API int dad3d_gather_landmarks_bary(const float* src_d, int32_t B, int32_t n_vertices, int32_t ncomp,
                                const int32_t* tri_idx_d, const float* bary_d, int32_t L, float* out_d,
                                dad3d_stream stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
DAD3D_API unsigned long long dad3d_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DAD3D_H_ */
