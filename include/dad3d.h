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

#define DAD3D_DECODE_CLUSTER 32 /* A/B aid: fused decode as 2x2 thread-block clusters with TMA multicast of both operands
                                  (measured slower on B200, so off by default) */

/* Widths of the fields of the 3DMM parameter vector, sliced in the reference's hard-coded order
 * shape, expression, jaw, rotation, eyeballs, neck, translation, scale
 * (model_training/model/flame.py:41-84 FlameParams.from_3dmm; dad_3dnet.yaml:5-12 gives 300/100/3/6/0/0/3/1 = 413). */
typedef struct dad3d_flame_layout {
  int32_t shape, expression, jaw, rotation, eyeballs, neck, translation, scale;
} dad3d_flame_layout;

/* decode flags */
#define DAD3D_ZERO_ROT 1        /* FLAMELayer.forward(zero_rot=True)   flame.py:225 */
#define DAD3D_ZERO_JAW 2        /* FLAMELayer.forward(zero_jaw=True)   flame.py:206 */
#define DAD3D_BLEND_FAST 4      /* (old name of today's default; accepted, no effect) */
#define DAD3D_BLEND_HILO 64     /* blend-shape product with fp16 hi/lo split operands (3 tensor-core products, 22-bit, vertices
                                   relL2 2e-7 vs fp64) through the generic tile engine.  DEFAULT (flag absent): ONE fp16 product
                                   (11-bit operand mantissa like TF32, template exact to 22 bits; vertices relL2 1.5e-5, inside
                                   the 1e-4 contract) in the dedicated A-stationary decode kernel (csrc/flame_decode.cuh) */
#define DAD3D_DECODE_PAIR 128   /* run the dedicated decode kernel as CTA pairs (cta_group::2) when the batch has at least two row
                                   tiles per SM.  Pairs are the default there since round 2 (environment DAD3D_DECODE_PAIR=0
                                   switches them off for A/B runs); the flag forces them regardless of the environment */
#define DAD3D_BLEND_SIMT 8      /* verification aid: blend-shape product on CUDA cores in fp32 (slow) */
#define DAD3D_DECODE_UNFUSED 16 /* A/B aid: tensor-core blend product to a v_posed scratch + separate skinning kernel
                                   (default: skinning / rotation / projection fused into the GEMM epilogue) */

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
/* Backward of dad3d_flame_decode: grad_params_d [B, num_params] = d L / d params given
 * grad_vertices_d [B,V,3] = dL/d(vertices3d) and / or grad_projected_d [B,V,2|3] = dL/d(projected) (either may be NULL).
 * What autograd gives the reference when its losses call HeadMesh.vertices_3d / reprojected_vertices
 * (losses/vertices_3d_loss.py:30-47, losses/reprojection_loss.py:22-46, train/flame_lightning_model.py:329-351).  The dense
 * part (d beta, d pose features) is a tcgen05 GEMM over the transposed basis; layouts without neck / eyeball pose only. */
DAD3D_API size_t dad3d_flame_backward_workspace_bytes(const dad3d_flame* h, int32_t B);
DAD3D_API int dad3d_flame_backward(dad3d_flame* h, const float* params_d, int32_t B, int32_t flags, const float* grad_vertices_d,
                                   const float* grad_projected_d, float image_size, int32_t to_2d, float* grad_params_d,
                                   void* workspace_d, size_t workspace_bytes, dad3d_stream stream);

DAD3D_API int dad3d_gather_landmarks(const float* src_d, int32_t B, int32_t n_vertices, int32_t ncomp, const int32_t* idx_d,
                           int32_t L, float* out_d, dad3d_stream stream);
/* barycentric variant (model_training/data/utils.py:120-206 get_68_landmarks): out[b,l,:] = sum_k bary[l,k]*src[b,tri[l,k],:] */
DAD3D_API int dad3d_gather_landmarks_bary(const float* src_d, int32_t B, int32_t n_vertices, int32_t ncomp,
                                const int32_t* tri_idx_d, const float* bary_d, int32_t L, float* out_d,
                                dad3d_stream stream);

/* ---- DAD-3DNet encoder ---------------------------------------------------------------------------------------------
 * Replaces the TorchScript module the reference's predictor runs (predictor.py:72,97-100), i.e.
 * FlameRegression.forward (model_training/model/flame_regression.py:87-106): pytorchcv ResNet-50 stages -> BiFPN(256) ->
 * heat-map head -> FusionLayer -> stage 4 -> three MLP heads.
 *
 * Weights arrive BN-folded, one record per GEMM-able layer, fp32, laid out [cout][R][S][cin] (channels-last taps).
 * Layer names (the folding itself is host-side Python, dad_3dheads_b200/encoder.py, from the reference's state_dict):
 *   "stem" (7x7 3->64)                       encoder.model.init_block.conv
 *   "s{1..4}u{k}c{1,2,3}"                    encoder.model.stage{i}.unit{k}.body.conv{1,2,3}; the first unit's c3 carries
 *                                            the projection shortcut K-concatenated: [W3 | W_identity_conv], bias b3 + bid
 *   "lat4".."lat7" (+ optional "lat3")      bifpn.p4 .. bifpn.p7; bifpn.p3 is normally composed into "b0_p3td" (its only
 *                                            consumer: (W_node W_p3) c2 + W_node b_p3), a "lat3" record keeps it separate
 *   "b{0,1}_{p4out,p5out,p6out,p7out}"       bifpn.bifpn.{0,1}.<node> (depthwise scale, pointwise, BN folded)
 *   "b{0,1}_{p6td,p5td,p4td,p3td}" and "..._u"   top-down nodes split in two: W*(w0 a) at the node's resolution and
 *                                            W*(w1 b) at the lower one (fusion scalars folded in; the second has no bias)
 *   "heat"                                   head.heatmap (3x3 256->68 + bias)
 *   "fusion"                                 fusion_layer.conv1x1 with K laid out [x 1024 | heat 68 + 60 zero | p5 256]
 *   "mlp1" (2048 -> 3x512)  "mlp2" (block-diagonal 1536 -> 403|10|136)   {shape,pose,landmarks}.logit_image.{0,3}
 * bifpn_fusion_w_h: [2][20] = per BiFPN block the normalised fusion weights relu(w)/sum + 1e-4, w1 [2][4] then w2 [3][4]
 *   (bifpn.py:105-108).
 * pieces / operand_format select the arithmetic (accumulation is always fp32):
 *   DAD3D_OPERAND_BF16: 1 = plain bf16 operands (1 tensor-core product, throughput mode), 2 = bf16 hi/lo (3 products,
 *     16-bit operand mantissa), 3 = bf16 three-way split (6 products, 24-bit operand mantissa: strict fp32 operands);
 *   DAD3D_OPERAND_FP16: 2 = fp16 hi/lo (3 products, 22-bit operand mantissa; weights scaled per output channel by a power
 *     of two that the epilogue undoes; activations saturate at +-65504 and carry an absolute representation error
 *     <= 2^-25 below |x| = 2^-3), 1 = plain fp16 (11-bit, TF32-class). */
#define DAD3D_OPERAND_BF16 0
#define DAD3D_OPERAND_FP16 1
typedef struct dad3d_conv_weights {
  const char* name;
  const float* weight_h;     /* [cout][R][S][cin] */
  const float* bias_h;       /* [cout] */
  int32_t cout, cin, R, S;
} dad3d_conv_weights;

DAD3D_API int dad3d_encoder_create(dad3d_encoder** out, const dad3d_conv_weights* layers, int32_t n_layers,
                                   const float* bifpn_fusion_w_h, int32_t pieces, int32_t operand_format, int32_t device);
DAD3D_API void dad3d_encoder_destroy(dad3d_encoder* enc);
DAD3D_API int dad3d_encoder_num_layers(const dad3d_encoder* enc);
DAD3D_API size_t dad3d_encoder_workspace_bytes(dad3d_encoder* enc, int32_t B);
/* images_d [B,3,256,256] NCHW fp32 (already normalised, predictor.py:195-203) ->
 *   params_d [B,413] (OUTPUT_3DMM_PARAMS), landmarks_d [B,68,2] (OUTPUT_2D_LANDMARKS, in [0,1] image units),
 *   heatmap_d [B,68,64,64] NCHW fp32 or NULL (OUTPUT_LANDMARKS_HEATMAP). */
DAD3D_API int dad3d_encoder_forward(dad3d_encoder* enc, const float* images_d, int32_t B, float* params_d,
                                    float* landmarks_d, float* heatmap_d, void* workspace_d, size_t workspace_bytes,
                                    dad3d_stream stream);

/* ---- device-side pre-processing (SURVEY §8f "next" row 2) --------------------------------------------------------------
 * dad3d_preprocess replaces FaceMeshPredictor._transform + _array_to_batch (predictor.py:85-89,195-203: albumentations
 *   LongestMaxSize -> PadIfNeeded -> Normalize -> HWC->CHW) for one image: image_d [H,W,3] uint8 RGB (device) ->
 *   out_d [3,img_size,img_size] fp32.  new_h/new_w are the letter-boxed sizes (py3round(dim * img_size / max(H,W)),
 *   computed by the caller exactly as predictor.py:117-123 does); the 8-bit bilinear resize is bit-exact with
 *   cv2.resize(INTER_LINEAR).  mean255_h / inv_std255_h: the three fp32 constants mean*255 and 1/(std*255). */
DAD3D_API int dad3d_preprocess(const uint8_t* image_d, int32_t H, int32_t W, int32_t new_h, int32_t new_w,
                               int32_t img_size, const float* mean255_h, const float* inv_std255_h, float* out_d,
                               dad3d_stream stream);
/* the same for B images of one size: images_d [B,H,W,3] uint8 -> out_d [B,3,img_size,img_size] fp32, one launch */
DAD3D_API int dad3d_preprocess_batch(const uint8_t* images_d, int32_t B, int32_t H, int32_t W, int32_t new_h, int32_t new_w,
                                     int32_t img_size, const float* mean255_h, const float* inv_std255_h, float* out_d,
                                     dad3d_stream stream);

/* Live timing of the dominant kernel (the tcgen05 tile engine) for bench.py's roofline: while on, every conv / linear
 * launch is bracketed by CUDA events on the launching stream.  profile_read synchronises those events and returns their
 * summed duration, the launch count and the ALGORITHMIC FLOPs (2 * true MACs, one product per MAC, no padding) of the
 * recorded launches, then resets the counters. */
DAD3D_API int dad3d_encoder_set_profile(dad3d_encoder* enc, int32_t on);
DAD3D_API int dad3d_encoder_profile_read(dad3d_encoder* enc, double* gemm_ms, long long* gemm_launches,
                                         double* useful_flops);

/* One recorded launch of the current profiling window (call BEFORE dad3d_encoder_profile_read, which clears the window):
 * layer name, device time, useful FLOPs, algorithmic HBM bytes (every operand once), and info8 = {M output pixels, K, N,
 * tensor-core products per MAC, tiles, tile N, pipeline stages, k-blocks per tile}.  Returns DAD3D_ERR_INVALID past the
 * last recorded launch. */
DAD3D_API int dad3d_encoder_profile_layer(dad3d_encoder* enc, int32_t index, char* name, int32_t name_cap, double* ms,
                                          double* useful_flops, double* algo_bytes, int32_t* info8);

/* test hooks: keep_all != 0 disables workspace reuse so that, after a forward, any activation can be read back by the
 * name of the layer that produced it ("stem", "s2u1c3", "b1_p4out", "cat", "fusion", "gap", "heat", "mlp2" ...) as fp32
 * NHWC with channels padded as stored; dims4 receives [N,H,W,C] (pass out_d = NULL to query the shape only). */
DAD3D_API int dad3d_encoder_set_debug(dad3d_encoder* enc, int32_t keep_all);
DAD3D_API int dad3d_encoder_read_activation(dad3d_encoder* enc, const char* name, float* out_d, size_t capacity_floats,
                                            int32_t* dims4, dad3d_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Benchmark evaluator hot spots (dad_3dheads_benchmark/benchmark.py, dad_3dheads_benchmark/utils.py),
 * batched over B heads, device pointers, asynchronous on `stream`.
 *   dad3d_eval_chamfer : out[b] = mean_i min_j |a[b,i] - b[b,j]|^2 -- the term the evaluator asks kaolin for
 *                        (utils.py:139: chamfer_distance(gt_face, aligned_pred, 1.0, 0.0)); a [B,na,3], b [B,nb,3].
 *   dad3d_eval_zn      : Z_n ordinal depth accuracy exactly as DADEvaluator.calc_zn computes it (benchmark.py:110-138):
 *                        distances gt->gt (torch.cdist formula), COLUMN-wise argsort, columns 1..top_k of the index matrix,
 *                        mean agreement of the z-order of (i, index[i][j]) between gt and pred.  pred, gt [B,K,3], K <= 4096.
 *   dad3d_eval_align   : out = scale[b] * (verts[b] @ rot[b]) + trans[b] for every vertex (utils.py:178-197; rot [B,3,3]
 *                        row-major, the procrustes tform of the 7 landmark pairs). */
DAD3D_API int dad3d_eval_chamfer(const float* a_d, int32_t na, const float* b_d, int32_t nb, int32_t B, float* out_d,
                                 dad3d_stream stream);
DAD3D_API int dad3d_eval_zn(const float* pred_d, const float* gt_d, int32_t K, int32_t B, int32_t top_k, float* out_d,
                            dad3d_stream stream);
DAD3D_API int dad3d_eval_align(const float* verts_d, int32_t nv, int32_t B, const float* scale_d, const float* rot_d,
                               const float* trans_d, float* out_d, dad3d_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU helpers: one process per GPU; images are independent, so the only exchanges are the start-up
 * broadcast of the constants and the per-batch all-gather of the outputs, both over NCCL (NVLink 5 / NVSwitch).  libnccl is
 * bound at run time (dlopen: the copy already loaded in the process, e.g. PyTorch's, else the system's).
 *   dad3d_comm_unique_id : rank 0 creates the 128-byte NCCL id; the host distributes it (any side channel, e.g. a
 *                          torch.distributed / MPI broadcast of 128 bytes)
 *   dad3d_comm_init      : every rank joins (collective)
 *   dad3d_bcast_constants: in-place broadcast of `bytes` bytes of device memory from `root` (FLAME bases, packed weights)
 *   dad3d_allgather_outputs: recv_d [world * bytes_per_rank] <- every rank's send_d [bytes_per_rank], rank-major
 *                          (params [B,413], vertices [B,5023,3], landmarks ...) */
typedef struct dad3d_comm dad3d_comm;
DAD3D_API int dad3d_comm_unique_id(uint8_t* id128_h);
DAD3D_API int dad3d_comm_init(dad3d_comm** out, const uint8_t* id128_h, int32_t rank, int32_t world, int32_t device);
DAD3D_API void dad3d_comm_destroy(dad3d_comm* c);
DAD3D_API int dad3d_bcast_constants(dad3d_comm* c, void* buf_d, size_t bytes, int32_t root, dad3d_stream stream);
DAD3D_API int dad3d_allgather_outputs(dad3d_comm* c, const void* send_d, void* recv_d, size_t bytes_per_rank, dad3d_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Rasteriser: the reference's native Sim3DR component (Sim3DR/lib/rasterize_kernel.cpp `_rasterize`
 * :238-292 and `_get_normal` :158-236; callers Sim3DR/Sim3DR.py:8-29, inference/pncc_estimator.py:16-43), bit-exact.
 *   dad3d_rasterize: z-buffer rendering of per-vertex colours with alpha = 1 (what Sim3DR.rasterize uses): vertices_d [nv,3]
 *     (x, y in pixels, z = depth, larger wins), triangles_d [ntri,3], colors_d [nv,c] in [0,1]; image_d [h,w,c] uint8 and
 *     depth_d [h,w] are read-modify-write exactly like the reference's buffers (depth is usually initialised to -1e8);
 *     key_ws_d = h*w 64-bit words of scratch; reverse != 0 flips the image rows.
 *   dad3d_vertex_normals: normalised sum of the incident (un-normalised) face normals per vertex; adj_offsets_d [nv+1] /
 *     adj_triangles_d = CSR list of each vertex's triangles in ascending order (fixed per topology). */
DAD3D_API int dad3d_rasterize(const float* vertices_d, const int32_t* triangles_d, const float* colors_d, int32_t ntri,
                              uint8_t* image_d, float* depth_d, unsigned long long* key_ws_d, int32_t h, int32_t w, int32_t c,
                              int32_t reverse, dad3d_stream stream);
DAD3D_API int dad3d_vertex_normals(const float* vertices_d, const int32_t* triangles_d, const int32_t* adj_offsets_d,
                                   const int32_t* adj_triangles_d, int32_t nver, float* normals_d, dad3d_stream stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
DAD3D_API unsigned long long dad3d_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DAD3D_H_ */
