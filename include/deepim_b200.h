/*
 * deepim_b200.h -- C ABI of libdeepim_b200.so (hand-written sm_100a CUDA, no CPU fallback).
 *
 * Drop-in boundary for the mx-DeepIM render-and-compare hot path.  Each entry point names the
 * reference interface it replaces (paths relative to the mx-DeepIM repo).  The reference binds its
 * native/GPU pieces from Python (mx.operator.CustomOp classes in deepim/operator_py/, the Cython
 * wrapper lib/flow_c/gpu_flow.pyx, the glumpy renderer class); the matching binding here is the
 * ctypes stub shown in INTEGRATION.md and shipped as mx-deepim_b200/deepim_b200/_capi.py.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers owned by the caller, contiguous, NCHW float32 unless
 *     noted; "host" in a parameter comment means a host pointer (small attribute arrays);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and allocates
 *     nothing (all scratch is sized by dim_ctx_create); one context per device, not thread-safe;
 *   - return 0 on success, non-zero on error; dim_last_error() gives the message
 *     (the reference ops raise Python exceptions instead; the Python shims re-raise);
 *   - H, W are fixed per context (480 x 640 in every shipped config).
 */
#ifndef DEEPIM_B200_H_
#define DEEPIM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DIM_API
#else
#define DIM_API __attribute__((visibility("default")))
#endif

typedef struct dim_ctx dim_ctx;

#define DIM_ABI_VERSION 2
DIM_API int32_t dim_abi_version(void);
DIM_API const char *dim_last_error(void);

/* Context: owns meshes, weights and all scratch.  max_verts/max_faces bound the largest mesh.
 * Replaces the per-process state of Render_Py.__init__ (lib/render_glumpy/render_py_multi.py:54-99)
 * and of the MXNet executor (deepim/core/tester.py:41-43). */
DIM_API int32_t dim_ctx_create(int32_t device, int32_t max_batch, int32_t height, int32_t width,
                               int32_t max_classes, int32_t max_verts, int32_t max_faces,
                               dim_ctx **out);
DIM_API void dim_ctx_destroy(dim_ctx *ctx);

/* Upload one class mesh (host pointers).  verts f32[V,3] metres, uvs f32[V,2], faces i32[F,3],
 * tex u8[Th,Tw,3] RGB with row 0 = v 0 (i.e. already flipped as render_py_multi.py:76 does).
 * Replaces data.objload + gloo.Program.bind + u_texture upload (render_py_multi.py:72-76). */
DIM_API int32_t dim_mesh_upload(dim_ctx *ctx, int32_t cls_idx, const float *verts_host,
                                const float *uvs_host, int32_t V, const int32_t *faces_host,
                                int32_t F, const uint8_t *tex_host, int32_t Th, int32_t Tw);

/* Rasterise B instances.  Replaces Render_Py.render (render_py_multi.py:101-129) plus the
 * post-render glue (deepim/core/tester.py:185-188,433-442; lib/utils/image.py:583-594).
 *   cls_idx i32[B] (device), pose f32[B,3,4] (device), K9 host f32[9], pixel_means_rgb host f64[3]
 *   trunc_u8: 1 = test path (uint8 truncation, tester.py:188), 0 = train path
 *   outputs (each may be NULL):
 *     out_image f32[B,3,H,W] RGB - means;  out_depth f32[B,1,H,W] metres;  out_mask f32[B,1,H,W]
 *     out_bgr f32[B,H,W,3] BGR in [0,255] (the Render_Py return layout);
 *     out_bbox i32[B,4] x0,x1,y0,y1 of out_mask (min/max nonzero col/row; -1 when empty). */
DIM_API int32_t dim_render(dim_ctx *ctx, const int32_t *cls_idx, const float *pose, int32_t B,
                           const float *K9_host, float znear, float zfar,
                           const double *pixel_means_rgb_host, int32_t trunc_u8, float *out_image,
                           float *out_depth, float *out_mask, float *out_bgr, int32_t *out_bbox,
                           void *stream);

/* ZoomMask forward (deepim/operator_py/zoom_mask.py:29-112).
 * in : mask_observed, mask_gt_observed, mask_rendered f32[B,1,H,W]; src_pose f32[B,3,4]; K9 host
 * out: 3 zoomed masks f32[B,1,H,W] (any may be NULL), zoom_factor f32[B,4],
 *      bbox i32[B,8] = observed x0,x1,y0,y1, rendered x0,x1,y0,y1 (may be NULL),
 *      status i32[B] (may be NULL): 1 where the observed mask is empty (the reference raises). */
DIM_API int32_t dim_zoom_mask_fwd(dim_ctx *ctx, const float *mask_observed,
                                  const float *mask_gt_observed, const float *mask_rendered,
                                  const float *src_pose, int32_t B, const float *K9_host,
                                  float *zoom_mask_observed, float *zoom_mask_gt_observed,
                                  float *zoom_mask_rendered, float *zoom_factor, int32_t *bbox,
                                  int32_t *status, void *stream);

/* ZoomImageWithFactor forward (zoom_image_with_factor.py:31-65). pixel_means_rgb host f32[3] is the
 * already-reversed attr (l.79-81).  images f32[B,3,H,W]. */
DIM_API int32_t dim_zoom_image_with_factor_fwd(dim_ctx *ctx, const float *zoom_factor,
                                               const float *image_observed,
                                               const float *image_rendered, int32_t B,
                                               const float *pixel_means_rgb_host,
                                               float *zoom_image_observed,
                                               float *zoom_image_rendered, void *stream);

/* Lit renderer (lib/render_glumpy/render_py_light_modelnet_multi.py:36-79 shader, 131-175 render): Lambert shading
 * colour = texel * ((1 - brightness_ratio) + brightness_ratio * clamp(cos(normal, light - position), 0, 1)) *
 * light_intensity, quantised to 8 bits like the framebuffer the reference reads back.  normals f32[V,3] (host) per
 * class; light_position / light_intensity f32[B,3] (device), position in the GL camera frame (x, -y, -z of the
 * OpenCV frame).  Outputs as dim_render (out_bgr holds the quantised colours as floats). */
DIM_API int32_t dim_mesh_upload_normals(dim_ctx *ctx, int32_t cls_idx, const float *normals_host, int32_t V);
DIM_API int32_t dim_render_lit(dim_ctx *ctx, const int32_t *cls_idx, const float *pose, int32_t B,
                               const float *K9_host, float znear, float zfar,
                               const double *pixel_means_rgb_host, const float *light_position,
                               const float *light_intensity, float brightness_ratio, float *out_image,
                               float *out_depth, float *out_mask, float *out_bgr, int32_t *out_bbox,
                               void *stream);

/* ZoomImage forward (zoom_image.py:26-107, the INPUT_MASK: False front end): the two boxes come from the
 * images themselves, valid = sum_c(image + pixel_mean_c) > 0.01; centre / crop / sampling as ZoomMask +
 * ZoomImageWithFactor.  pixel_means_rgb_host = the op's (already reversed) pixel_means attr.
 * bbox i32[B,8] / status i32[B] as dim_zoom_mask_fwd (may be NULL). */
DIM_API int32_t dim_zoom_image_fwd(dim_ctx *ctx, const float *image_observed, const float *image_rendered,
                                   const float *src_pose, int32_t B, const float *K9_host,
                                   const float *pixel_means_rgb_host, float *zoom_image_observed,
                                   float *zoom_image_rendered, float *zoom_factor, int32_t *bbox,
                                   int32_t *status, void *stream);

/* GroupPicker (group_picker.py:22-60): forward out[b] = in[b, g*cg:(g+1)*cg] (g = group_idx[b], a float like the
 * NDArray the reference reads); backward = 1 scatters out_grad (in) into a zero [B,channels,...] gradient. */
DIM_API int32_t dim_group_picker(dim_ctx *ctx, const float *in, const float *group_idx, int32_t B,
                                 int32_t channels, int32_t group_num, int64_t elems_per_channel,
                                 int32_t backward, float *out, void *stream);

/* ZoomMaskWithFactor forward (zoom_mask_with_factor.py:29-64): mask f32[B,1,H,W]. */
DIM_API int32_t dim_zoom_mask_with_factor_fwd(dim_ctx *ctx, const float *zoom_factor,
                                              const float *mask, int32_t B, int32_t b_inv_zoom,
                                              float *zoom_mask, void *stream);

/* ZoomFlow forward (zoom_flow.py:28-71): flow f32[B,2,H,W]; flow_weights/zoom_flow_weights
 * f32[B,fw_channels,H,W] (1, or 2 as tiled by batch_updater_py_multi.py:293-296) used only when
 * b_inv_zoom == 0 (may be NULL). */
DIM_API int32_t dim_zoom_flow_fwd(dim_ctx *ctx, const float *zoom_factor, const float *flow,
                                  const float *flow_weights, int32_t fw_channels, int32_t B,
                                  int32_t b_inv_zoom, float *zoom_flow, float *zoom_flow_weights,
                                  void *stream);

/* ZoomDepth forward (zoom_depth.py:24-44): depth f32[B,1,H,W] x2. */
DIM_API int32_t dim_zoom_depth_fwd(dim_ctx *ctx, const float *zoom_factor,
                                   const float *depth_observed, const float *depth_rendered,
                                   int32_t B, float *zoom_depth_observed,
                                   float *zoom_depth_rendered, void *stream);

/* ZoomTrans forward / backward (zoom_trans.py:22-74): trans f32[B,3]. */
DIM_API int32_t dim_zoom_trans_fwd(dim_ctx *ctx, const float *zoom_factor, const float *trans_delta,
                                   int32_t B, int32_t b_inv_zoom, float *zoom_trans_delta,
                                   void *stream);
DIM_API int32_t dim_zoom_trans_bwd(dim_ctx *ctx, const float *zoom_factor, const float *out_grad,
                                   int32_t B, int32_t b_inv_zoom, int32_t b_zoom_grad,
                                   float *trans_grad, void *stream);

/* mask_observed := end-exclusive bbox rectangle of mask_rendered
 * (lib/pair_matching/data_pair.py:93-105).  bbox i32[B,4] as written by dim_render. */
DIM_API int32_t dim_update_mask_box(dim_ctx *ctx, const int32_t *bbox, int32_t B,
                                    float *mask_observed, void *stream);

/* SE(3) compose in float64 (lib/pair_matching/RT_transform.py:127-151).
 * pose_src f64[B,3,4], se3 f32[B,7] = (quat w,x,y,z un-normalised, trans), T_means/T_stds host
 * f64[3], rot_coord 0 MODEL / 1 CAMERA / 2 CAMERA_NEW; pose_out f64[B,3,4]. */
DIM_API int32_t dim_se3_compose(dim_ctx *ctx, const double *pose_src, const float *se3, int32_t B,
                                const double *T_means_host, const double *T_stds_host,
                                int32_t rot_coord, double *pose_out, void *stream);

/* Reprojection-flow labels (lib/flow_c/gpu_flow_kernel.cu:32-69 flow_kernel, gpu_flow.pyx:24-41).
 * depth_src, depth_tgt f32[B,1,H,W]; KT f32[B,3,4] = K.T_src->tgt; Kinv host f32[9];
 * flow f32[B,2,H,W] (dh,dw), valid f32[B,1,H,W]. */
DIM_API int32_t dim_flow_fwd(dim_ctx *ctx, const float *depth_src, const float *depth_tgt,
                             const float *KT, const float *Kinv_host, int32_t B, float *flow,
                             float *valid, void *stream);

/* Transform3D forward / backward (deepim/operator_py/transform3d.py:34-151).
 * point_cloud f32[B,3,N], rotation f32[B,4], translation f32[B,3], pose_src f32[B,3,4]. */
DIM_API int32_t dim_transform3d_fwd(dim_ctx *ctx, const float *point_cloud, const float *rotation,
                                    const float *translation, const float *pose_src, int32_t B,
                                    int32_t N, const float *T_means_host, const float *T_stds_host,
                                    int32_t rot_coord, float *out_points, void *stream);
DIM_API int32_t dim_transform3d_bwd(dim_ctx *ctx, const float *out_grad, const float *point_cloud,
                                    const float *rotation, const float *translation,
                                    const float *pose_src, int32_t B, int32_t N,
                                    const float *T_means_host, const float *T_stds_host,
                                    int32_t rot_coord, float *rot_grad, float *trans_grad,
                                    void *stream);

/* Train-time inter-iteration update (lib/pair_matching/batch_updater_py_multi.py:91-328,
 * batchUpdaterPyMulti.forward): compose the predicted delta onto src_pose, re-render WITHOUT uint8
 * truncation (float32 image - float32 means, l.184,234), recompute the labels rot (quaternion of
 * calc_RT_delta(..., "QUAT")), trans, the reprojection flow against depth_gt_observed and its weights
 * (valid tiled to 2 channels).  mask_observed / image_observed / tgt_pose stay fixed (not touched).
 *   in : cls_idx i32[B], src_pose/tgt_pose f32[B,3,4], rot_est f32[B,4], trans_est f32[B,3],
 *        depth_gt_observed f32[B,1,H,W] (may be NULL when flow == NULL); K9 / T_means / T_stds /
 *        pixel_means_rgb host f64
 *   out: image_rendered f32[B,3,H,W], depth_rendered/mask_rendered f32[B,1,H,W], src_pose_new f32[B,3,4],
 *        rot_label f32[B,4], trans_label f32[B,3], flow f32[B,2,H,W], flow_weights f32[B,2,H,W]
 *        (flow, flow_weights may be NULL together). */
DIM_API int32_t dim_train_update(dim_ctx *ctx, const int32_t *cls_idx, const float *src_pose,
                                 const float *rot_est, const float *trans_est, const float *tgt_pose,
                                 const float *depth_gt_observed, int32_t B, const double *K9_host,
                                 float znear, float zfar, const double *pixel_means_rgb_host,
                                 const double *T_means_host, const double *T_stds_host,
                                 int32_t rot_coord, float *image_rendered, float *depth_rendered,
                                 float *mask_rendered, float *src_pose_new, float *rot_label,
                                 float *trans_label, float *flow, float *flow_weights, void *stream);

/* FlowNetS weights (deepim/symbols/deepIM_flownet.py:63-116,716-717; MXNet layouts: Convolution
 * (Cout,Cin,kh,kw), FullyConnected (out,in)).  Host float32 pointers, 14 (weight,bias) pairs in the
 * order flow_conv1, conv2, conv3, conv3_1, conv4, conv4_1, conv5, conv5_1, conv6, conv6_1, fc6,
 * fc7, rot, trans.  Replaces load_param + Module.init_params (deepim/core/tester.py:41-43). */
DIM_API int32_t dim_net_load(dim_ctx *ctx, const float *const *weights_host,
                             const float *const *biases_host);

/* precision of the conv stack */
#define DIM_PREC_BF16 0   /* one bf16 tcgen05 pass, fp32 accumulate (throughput mode)          */
#define DIM_PREC_BF16X3 1 /* hi/lo split, 3 tcgen05 passes, ~fp32 accuracy (parity mode)       */
#define DIM_PREC_FP16 2   /* one fp16 tcgen05 pass (11 significant bits), fp32 accumulate, saturating stores:
                             the single-pass mode that meets the 1e-4 rot / 1e-3 trans se3 tolerance (headline mode) */

/* Encoder + fc + heads on already-zoomed blobs (get_convs, deepIM_flownet.py:53-116; heads
 * l.716-717): inputs f32 NCHW as the op surface produces them; rot f32[B,4] raw quaternion,
 * trans f32[B,3] zoomed translation. */
DIM_API int32_t dim_net_fwd(dim_ctx *ctx, const float *zoom_image_observed,
                            const float *zoom_image_rendered, const float *zoom_mask_observed,
                            const float *zoom_mask_rendered, int32_t B, int32_t precision,
                            float *rot, float *trans, void *stream);

/* The fused test-time loop (deepim/core/tester.py:340-485 with FAST_TEST / UPDATE_MASK
 * box_rendered): n_iter x (render -> bbox+zoom -> FlowNetS -> ZoomTrans^-1 -> RT_transform),
 * everything on the device, no host sync.
 *   image_observed f32[B,3,H,W] RGB-mean (constant over iterations)
 *   cls_idx i32[B]; pose_init f64[B,3,4]
 *   outputs (device): poses f64[n_iter,B,3,4], se3 f32[n_iter,B,7], zoom_factor f32[n_iter,B,4],
 *   bbox i32[n_iter,B,8]; any of the last three may be NULL.
 *   pose_override f64[n_iter,B,3,4] or NULL: when given, iteration `it` starts from
 *   pose_override[it] instead of the previous estimate (teacher forcing for parity tests). */
DIM_API int32_t dim_refine(dim_ctx *ctx, const float *image_observed, const int32_t *cls_idx,
                           const double *pose_init, int32_t B, int32_t n_iter, const float *K9_host,
                           float znear, float zfar, const double *pixel_means_rgb_host,
                           int32_t precision, const double *pose_override, double *poses,
                           float *se3, float *zoom_factor, int32_t *bbox, void *stream);

/* Host-buffer convenience around dim_refine (what deepim/core/tester.py:pred_eval would call):
 * image_observed_u8 host u8[B,H,W,3] BGR (as cv2.imread returns; transformed on device as
 * lib/utils/image.py:583-594), cls_idx host, pose_init host f64; poses_out host f64[n_iter,B,3,4].
 * Pinned buffers are recommended.  Synchronises the stream before returning. */
DIM_API int32_t dim_refine_host(dim_ctx *ctx, const uint8_t *image_observed_u8_host,
                                const int32_t *cls_idx_host, const double *pose_init_host,
                                int32_t B, int32_t n_iter, const float *K9_host, float znear,
                                float zfar, const double *pixel_means_rgb_host, int32_t precision,
                                double *poses_out_host, float *se3_out_host, void *stream);

/* Per-iteration status of the LAST dim_refine / dim_refine_host(_async) call on this context, copied device -> host
 * asynchronously on `stream` (the stream that call ran on): [min(n_iter,8), B] int32.  0 = ok; bit 0 = the rendered mask
 * of that iteration was empty (the reference crashes there: np.min of an empty array, zoom_mask.py:55-58; here the
 * fallback zoom factor was used and the instance's pose is meaningless); bit 1 = class index out of range or no mesh
 * uploaded for that class (the reference indexes a python list and raises). */
DIM_API int32_t dim_refine_status(dim_ctx *ctx, int32_t B, int32_t n_iter, int32_t *status_host, void *stream);

/* Same as dim_refine_host but returns right after enqueueing the copies and kernels on `stream`
 * (no synchronisation): the host output buffers are valid once the stream has been synchronised.
 * Lets a caller overlap the H2D copy of batch k+1 (second context / stream) with the compute of k. */
DIM_API int32_t dim_refine_host_async(dim_ctx *ctx, const uint8_t *image_observed_u8_host,
                                      const int32_t *cls_idx_host, const double *pose_init_host,
                                      int32_t B, int32_t n_iter, const float *K9_host, float znear,
                                      float zfar, const double *pixel_means_rgb_host,
                                      int32_t precision, double *poses_out_host, float *se3_out_host,
                                      void *stream);

/* BGR u8 HWC -> RGB-mean f32 CHW on device (lib/utils/image.py:583-594 transform). */
DIM_API int32_t dim_transform_image_u8(dim_ctx *ctx, const uint8_t *bgr_u8, int32_t B,
                                       const double *pixel_means_rgb_host, float *image,
                                       void *stream);

/* Test hooks (not part of the drop-in surface): copy the bf16 NHWC activation buffer feeding conv
 * layer idx (10 = fc6 input) to the host, and its geometry
 * out8 = rows, cols, C, py, px, Ho, Wo, Cout. */
DIM_API int32_t dim_debug_activation(dim_ctx *ctx, int32_t idx, int32_t lo, void *host_dst,
                                     uint64_t bytes);
DIM_API int32_t dim_debug_layer_geometry(dim_ctx *ctx, int32_t idx, int32_t *out8);
/* tuning hooks (tools/conv_lab.py; not part of the drop-in surface).
 * dim_debug_set_option: run-time kernel-variant switch.  Keys: "pair_mask": bit i puts conv layer i (1..9) on the
 *   cta_group::2 kernel (default: conv2); "conv1_stack": 1 (default) = stacked-filter-rows conv1 kernel for the
 *   single-pass precisions, 0 = rolling-strip kernel; "graph": 1 (default) = replay the refinement chain as a CUDA graph.
 *   Synchronises the device and drops the cached launch descriptors / graphs.
 * dim_debug_layer_profile: enable = 1 records CUDA events around each conv layer of every dim_net_fwd / dim_refine
 *   iteration; ms10 (nullable) receives the 10 layer times of the LAST forward pass; enable = 0 stops. */
DIM_API int32_t dim_debug_set_option(dim_ctx *ctx, const char *key, int32_t value);
DIM_API int32_t dim_debug_layer_profile(dim_ctx *ctx, int32_t enable, float *ms10);

/* Stage profiling of dim_refine with CUDA events on the launching stream (used by bench.py for the
 * live roofline numbers).  enable=1 starts recording; dim_profile_read synchronises the device and
 * returns accumulated milliseconds since the last read as ms[4] = render, bbox+zoom, conv tower,
 * fc+head+compose, and the number of recorded iterations. */
DIM_API int32_t dim_profile_enable(dim_ctx *ctx, int32_t enable);
DIM_API int32_t dim_profile_read(dim_ctx *ctx, float *ms4, int32_t *iterations);

/* ADD / ADI pose error (lib/utils/pose_error.py:72-108; LM6D_REFINE.evaluate_pose_add l.418-424 picks ADI for the
 * symmetric classes): poses f64[M,3,4] (device), points f64[N,3] model points (device), err f64[M].
 * symmetric = 0: mean point distance; 1: mean distance from every GT-transformed point to the nearest
 * estimate-transformed point (brute force, float64). */
DIM_API int32_t dim_pose_error(dim_ctx *ctx, const double *poses_est, const double *poses_gt, int32_t M,
                               const double *points, int32_t N, int32_t symmetric, double *err,
                               void *stream);

/* Average 2D re-projection error and rotation / translation distances of pose pairs (float64, device pointers):
 * err3[m] = { arp_2d (pixels; lib/utils/pose_error.py:55-69), rotation distance in degrees, translation distance in metres
 * (lib/pair_matching/RT_transform.py:162-173 calc_rt_dist_m) } -- the inputs of LM6D_REFINE.evaluate_pose (5 cm 5 deg,
 * lib/dataset/LM6D_REFINE.py:278-371) and evaluate_pose_arp_2d (Proj. 2D, l.514-). K9_dev: 9 doubles on the device. */
DIM_API int32_t dim_pose_error_2d(dim_ctx *ctx, const double *poses_est, const double *poses_gt, int32_t M, const double *points,
                                  int32_t N, const double *K9_dev, double *err3, void *stream);

/* End-point error of a predicted flow (deepim/core/tester.py:573-589 calc_EPE_one_pair; the non-FAST_TEST evaluation):
 * flows [B,2,H,W], visible / bg [B,1,H,W] float32 device; out6 [B,6] float64 device =
 * { sum |gt - pred| over all pixels, pixel count, sum over visible == 1, sum(visible), sum over visible | bg, count }. */
DIM_API int32_t dim_flow_epe(dim_ctx *ctx, const float *flow_pred, const float *flow_gt, const float *visible, const float *bg,
                             int32_t B, double *out6, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Training step of the refiner network (train graph: deepim/symbols/deepIM_flownet.py:121-365 decoder,
 * flow / mask / point-matching losses; optimiser: deepim/train.py:296-304 SGD-momentum, one update per inner
 * iteration as deepim/core/module.py:1131-1137).  Replaces Module.forward_backward + Module.update for this
 * network; the zoom front of the train symbol (ZoomMask / ZoomImageWithFactor / ZoomFlow, symbol:391-489) is
 * the dim_zoom_* calls above, the inter-iteration batch update is dim_train_update.
 *
 * Parameters live in ONE flat fp32 vector in MXNet layouts: for each of the 22 trainable tensors in the order
 * returned by dim_train_param_info (flow_conv1 ... conv6_1, fc6, fc7, rot, trans, Convolution1, deconv5,
 * upsample_flow6to5, Convolution2, deconv4, upsample_flow5to4, Convolution3, mask_conv3) weight then bias,
 * followed by the frozen bilinear upsampling_weight (2,1,32,32) and mask_upsampling_weight (1,1,32,32):
 * 57 749 164 floats.  One tensor is permuted: fc6_weight is stored (256, h*10+w, c) -- the NHWC order of the
 * conv6_1 activation it multiplies -- instead of MXNet's (256, c*80 + h*10 + w) (the Python host permutes on
 * load / get; element-wise consumers such as the all-reduce and SGD do not care).  Gradients use the same layout (that is the buffer a data-parallel caller all-reduces
 * with NCCL between dim_train_forward_backward and dim_train_sgd_update; kvstore replacement,
 * deepim/core/module.py:616-635). */
DIM_API int32_t dim_train_create(dim_ctx *ctx, int32_t max_points);
DIM_API int64_t dim_train_param_count(dim_ctx *ctx);
DIM_API int32_t dim_train_param_info(int32_t idx, const char **name, int64_t *weight_numel,
                                     int64_t *bias_numel);
/* flat_host: host pointer.  Also (re)loads the inference network of this context. */
DIM_API int32_t dim_train_load_params(dim_ctx *ctx, const float *flat_host, int64_t n, void *stream);
/* which = 0: parameters, 1: momentum.  Synchronises the stream. */
DIM_API int32_t dim_train_get_params(dim_ctx *ctx, float *flat_host, int64_t n, int32_t which,
                                     void *stream);
/* Device pointers, fp32 NCHW: zoomed images (B,3,H,W), zoomed masks (B,1,H,W), zoom_factor (B,4), zoomed flow
 * label (B,2,H,W) and weights (B,2,H,W), zoomed GT mask (B,1,H,W), src_pose (B,3,4), point clouds (B,3,N).
 * Outputs: rot_est_norm (B,4) = L2Normalization(rot), trans_est (B,3) = invZoomTrans, flow_est (B,2,H,W)
 * (= flow_est_crop * NORMALIZE_FLOW, nullable), mask_prob (B,1,H,W) (nullable), losses4 = [sum flow_loss,
 * sum point_matching_loss, sum mask BCE, weighted objective], grads (flat, see above; NULL = forward only:
 * the non-FAST_TEST outputs of the test graph, symbol:624-713).
 * With grads == NULL the labels (zoom_flow ... point clouds, losses4) may all be NULL: pure test-time forward of
 * the full graph; rot_raw (B,4, nullable) is the un-normalised quaternion the test graph concatenates into se3.
 * bucket_events (nullable): n_buckets cudaEvent_t handles; event k is recorded (on an internal stream) as soon as
 * the gradients of every tensor with table index >= bucket_first_tensor[k] are complete, so a data-parallel
 * caller can start the NCCL all-reduce of that slice of `grads` while the rest of the backward pass still runs.
 * All work is joined back into `stream` before the call's stream order ends. */
DIM_API int32_t dim_train_forward_backward(
    dim_ctx *ctx, const float *zoom_image_observed, const float *zoom_image_rendered,
    const float *zoom_mask_observed, const float *zoom_mask_rendered, const float *zoom_factor,
    const float *zoom_flow, const float *zoom_flow_weights, const float *zoom_mask_gt_observed,
    const float *src_pose, const float *point_cloud_model, const float *point_cloud_weights,
    const float *point_cloud_observed, int32_t B, int32_t N, float *rot_est_norm, float *trans_est,
    float *flow_est, float *mask_prob, float *losses4, float *grads, float *rot_raw,
    void *const *bucket_events,
    const int32_t *bucket_first_tensor, int32_t n_buckets, void *stream);
/* Loss weights / normalisers of the training step and the pose parameterisation shared with the refinement loop: the
 * values the reference reads from its yaml (experiments/deepim/cfgs/...: train.LW_FLOW / LW_MASK / LW_PM, NUM_3D_SAMPLE,
 * NORMALIZE_3D_POINT, NORMALIZE_FLOW, network.TRANS_MEANS / TRANS_STDS, ROT_COORD).  Defaults = the shipped LM6d config
 * (0.25, 0.03, 0.1, 3000, 0.1, 20, means 0, stds 1, CAMERA).  The MakeLoss grad_scale of the point-matching loss is
 * lw_pm / num_3d_sample whatever the number of points passed per call (deepIM_flownet.py:330-336).
 * dim_train_set_config applies to dim_train_forward_backward of this context; trans_means / trans_stds / rot_coord also
 * drive dim_refine / dim_refine_host (RT_transform with T_means / T_stds / rot_coord, tester.py:452-461) and may be set
 * without dim_train_create (the other fields are then stored and used once a training state exists). */
typedef struct dim_train_config {
  float lw_flow, lw_mask, lw_pm;
  float num_3d_sample;
  float normalize_3d_point;
  float normalize_flow;
  float trans_means[3], trans_stds[3];
  int32_t rot_coord; /* 0 = MODEL, 1 = CAMERA */
} dim_train_config;
DIM_API int32_t dim_train_set_config(dim_ctx *ctx, const dim_train_config *cfg);
DIM_API int32_t dim_train_get_config(dim_ctx *ctx, dim_train_config *cfg);
/* mom = momentum*mom - lr*(rescale_grad*grad + wd*w); w += mom  (wd on *_weight only; the two bilinear kernels
 * are frozen), then refreshes every bf16 operand pack of the context from the new master weights. */
DIM_API int32_t dim_train_sgd_update(dim_ctx *ctx, const float *grads, float lr, float momentum, float wd,
                                     float rescale_grad, void *stream);
/* Test hooks: intermediates of the training step (ids in train.cu) and their geometry
 * out7 = Hp, Wp, py, px, C, H, W. */
DIM_API int32_t dim_train_debug_tensor(dim_ctx *ctx, int32_t id, void *host_dst, uint64_t bytes);
DIM_API int32_t dim_train_debug_geometry(dim_ctx *ctx, int32_t id, int32_t *out7);
/* ms7 = device time of the phases of the last dim_train_forward_backward (with gradients) on the caller's stream:
 * encoder fwd, decoder fwd, losses + pose heads, fc/head backward, decoder backward, encoder data-gradient chain,
 * wait for the weight-gradient stream.  Synchronises the device. */
DIM_API int32_t dim_train_debug_phases(dim_ctx *ctx, float *ms7);

/* number of kernel launches issued by this library since the counter was last reset */
DIM_API int64_t dim_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* DEEPIM_B200_H_ */
