// capi.cu -- the extern "C" surface declared in include/deepim_b200.h.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dim {

static thread_local char g_err[1024] = "";
long long g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// implemented in raster.cu / zoom.cu / geom.cu / net.cu
struct LitParams { const float *light_pos, *light_int; float a0, a1; };  // device [B,3] each; a0 = 1 - ratio, a1 = ratio
int render_launch(dim_ctx *, const int *, const float *, int, const float *, float, float, const double *, int, float *,
                  float *, float *, float *, int *, float4 *, cudaStream_t, const LitParams *lit = nullptr);
int zoom_gather_launch(dim_ctx *, int mode, const float *src, float *dst, const float *zf, int B, int C, int inv,
                       const float *param, cudaStream_t);
int zoom_factor_launch(dim_ctx *, const float *, const float *, int C, const float *, int B, const float *K9, float *,
                       int *, int *, cudaStream_t, const float *img_means = nullptr);
int pose_error_launch(const double *, const double *, int M, const double *, int N, int symmetric, double *, cudaStream_t);
int epe_launch(const float *, const float *, const float *, const float *, int B, int P, double *, cudaStream_t);
int pose_error2d_launch(const double *, const double *, int M, const double *, int N, const double *K9, double *, cudaStream_t);
int group_pick_launch(const float *, const float *, int B, int Ctot, int groups, size_t n, float *, int backward, cudaStream_t);
int zoom_factor_from_ren_launch(dim_ctx *, const int *, const float *, int B, const float *K9, float *, int *, int *,
                                cudaStream_t);
int box_mask_launch(dim_ctx *, const int *, int B, float *, cudaStream_t);
int zoom_fused_launch(dim_ctx *, const float4 *, const float4 *, const float *, const float *, int B, int Hs, int Ws,
                      int pad, __nv_bfloat16 *, __nv_bfloat16 *, cudaStream_t, int f16, const double *means_d);
int pack_obs4_launch(dim_ctx *, const float *, int B, float4 *, const double *means, cudaStream_t);
int transform_u8_obs4_launch(dim_ctx *, const uint8_t *, int B, const double *, float4 *, cudaStream_t);
int pack_nhwc8_launch(dim_ctx *, const float *, const float *, const float *, const float *, int B, int Hs, int Ws,
                      int pad, __nv_bfloat16 *, __nv_bfloat16 *, cudaStream_t, int f16);
int flow_launch(dim_ctx *, const float *, const float *, const float *, const float *, int B, float *, float *,
                float *, cudaStream_t);
int train_pose_launch(const float *, const float *, const float *, const float *, int B, const double *,
                      const double *, int, const double *, float *, float *, float *, float *, cudaStream_t);
int se3_compose_launch(const double *, const float *, int B, const double *, const double *, int, double *, float *,
                       cudaStream_t);
int f64_to_f32_launch(const double *, float *, int n, cudaStream_t);
int zoom_trans_launch(const float *, const float *, int B, int mul, int scale_xy, float *, cudaStream_t);
int transform3d_fwd_launch(const float *, const float *, const float *, const float *, int, int, const float *,
                           const float *, int, float *, cudaStream_t);
int transform3d_bwd_launch(const float *, const float *, const float *, const float *, const float *, int, int,
                           const float *, const float *, int, float *, float *, cudaStream_t);
int transform_u8_launch(dim_ctx *, const uint8_t *, int B, const double *, float *, cudaStream_t);
int net_create(dim_ctx *);
void net_destroy(dim_ctx *);
int net_load(dim_ctx *, const float *const *, const float *const *);
void net_input_geometry(dim_ctx *, int *rows, int *cols, int *pad, __nv_bfloat16 **hi, __nv_bfloat16 **lo);
int net_forward(dim_ctx *, int B, int precision, const float *zoom_factor, float *rot, float *trans, float *se3,
                cudaStream_t, cudaEvent_t after_conv);
int net_debug_activation(dim_ctx *, int idx, int lo, void *host_dst, size_t bytes);
void net_layer_geometry(dim_ctx *, int idx, int *out);
bool net_graph_safe(dim_ctx *);
int net_set_option(dim_ctx *, const char *key, int value);
int net_layer_profile(dim_ctx *, int enable, float *ms10);
// train.cu
struct TrainIO {
  const float *zio, *zir, *zmo, *zmr, *zoom_factor, *zflow, *zfw, *zmask_gt, *src_pose, *pc_model, *pc_weights, *pc_observed;
  int B, N;
  float *rot_est_norm, *trans_est, *flow_est, *mask_prob, *losses, *grads;
  float *rot_raw;
  void *const *bucket_events;
  const int *bucket_first_tensor;
  int n_buckets;
};
int train_create(dim_ctx *, int max_points);
void train_destroy(dim_ctx *);
int train_load_params(dim_ctx *, const float *flat_host, size_t n, cudaStream_t);
int train_get_params(dim_ctx *, float *flat_host, size_t n, int which, cudaStream_t);
size_t train_param_count(dim_ctx *);
int train_param_info(int idx, const char **name, long long *w_numel, long long *b_numel);
int train_forward_backward(dim_ctx *, const TrainIO &, cudaStream_t);
int train_sgd_update(dim_ctx *, const float *grads, float lr, float momentum, float wd, float rescale, cudaStream_t);
int train_debug_tensor(dim_ctx *, int id, void *host, size_t bytes);
void train_debug_geometry(dim_ctx *, int id, int *out);
int train_debug_phases(dim_ctx *, float *ms7);

template <typename T>
static int ctx_alloc(dim_ctx *ctx, T **p, size_t n) {
  void *q = nullptr;
  DIM_CHECK(cudaMalloc(&q, n * sizeof(T)));
  ctx->owned.push_back(q);
  *p = reinterpret_cast<T *>(q);
  return 0;
}

}  // namespace dim

using namespace dim;

// captured refinement graphs bake in launch dimensions and kernel choices: anything that changes them drops the graphs
static void drop_graphs(dim_ctx *ctx) {
  if (ctx->graphs.empty()) return;
  cudaDeviceSynchronize();
  for (auto &g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  ctx->graphs.clear();
}

extern "C" {

DIM_API int32_t dim_abi_version(void) { return DIM_ABI_VERSION; }
DIM_API const char *dim_last_error(void) { return g_err; }
DIM_API int64_t dim_launch_count(int32_t reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

DIM_API int32_t dim_ctx_create(int32_t device, int32_t max_batch, int32_t H, int32_t W, int32_t max_classes,
                               int32_t max_verts, int32_t max_faces, dim_ctx **out) {
  DIM_REQUIRE(out != nullptr, "dim_ctx_create: out is NULL");
  DIM_REQUIRE(max_batch >= 1 && H >= 16 && W >= 16 && (W % 4) == 0, "dim_ctx_create: bad sizes");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("dim_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
              cudaGetErrorString(e));
    return 10;
  }
  DIM_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  DIM_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("dim_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
              prop.minor);
    return 11;
  }
  dim_ctx *ctx = new dim_ctx();
  ctx->device = device; ctx->max_batch = max_batch; ctx->H = H; ctx->W = W;
  ctx->max_classes = max_classes; ctx->max_verts = max_verts; ctx->max_faces = max_faces;
  ctx->num_sms = prop.multiProcessorCount;
  const size_t P = (size_t)H * W, Bm = (size_t)max_batch;
  int rc = 0;
  rc |= ctx_alloc(ctx, &ctx->meshes, (size_t)max_classes);
  rc |= ctx_alloc(ctx, &ctx->pverts, Bm * (size_t)max_verts);
  rc |= ctx_alloc(ctx, &ctx->vis, Bm * P);
  rc |= ctx_alloc(ctx, &ctx->vbox, Bm * 4);
  rc |= ctx_alloc(ctx, &ctx->bbox8, Bm * 8);
  rc |= ctx_alloc(ctx, &ctx->status, Bm);
  rc |= ctx_alloc(ctx, &ctx->cls_flag, Bm);
  rc |= ctx_alloc(ctx, &ctx->status_hist, 8 * Bm);
  rc |= ctx_alloc(ctx, &ctx->zoom_factor, Bm * 4);
  rc |= ctx_alloc(ctx, &ctx->image_rendered, Bm * 3 * P);
  rc |= ctx_alloc(ctx, &ctx->depth_rendered, Bm * P);
  rc |= ctx_alloc(ctx, &ctx->mask_rendered, Bm * P);
  rc |= ctx_alloc(ctx, &ctx->bbox_ren, Bm * 4);
  rc |= ctx_alloc(ctx, &ctx->pose_cur, Bm * 12);
  rc |= ctx_alloc(ctx, &ctx->pose_cur_f32, Bm * 12);
  rc |= ctx_alloc(ctx, &ctx->se3_cur, Bm * 7);
  rc |= ctx_alloc(ctx, &ctx->ren4, Bm * P);
  rc |= ctx_alloc(ctx, &ctx->obs4, Bm * P);
  rc |= ctx_alloc(ctx, &ctx->image_observed_u8, Bm * 3 * P);
  rc |= ctx_alloc(ctx, &ctx->cls_dev, Bm);
  rc |= ctx_alloc(ctx, &ctx->poses_dev, 8 * Bm * 12);
  rc |= ctx_alloc(ctx, &ctx->se3_hist_dev, 8 * Bm * 7);
  if (rc) { dim_ctx_destroy(ctx); return 12; }
  ctx->meshes_host.assign(max_classes, MeshDev{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, nullptr});
  DIM_CHECK(cudaMemset(ctx->meshes, 0, sizeof(MeshDev) * max_classes));
  DIM_CHECK(cudaMemset(ctx->vis, 0xFF, sizeof(unsigned long long) * Bm * P));  // all pixels empty
  DIM_CHECK(cudaMemset(ctx->pverts, 0, sizeof(PVert) * Bm * max_verts));
  DIM_CHECK(cudaMemset(ctx->cls_flag, 0, sizeof(int) * Bm));
  DIM_CHECK(cudaMemset(ctx->status_hist, 0, sizeof(int) * 8 * Bm));
  if (net_create(ctx)) { dim_ctx_destroy(ctx); return 13; }
  DIM_CHECK(cudaDeviceSynchronize());
  *out = ctx;
  return 0;
}

DIM_API void dim_ctx_destroy(dim_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  train_destroy(ctx);
  net_destroy(ctx);
  for (auto &g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  for (void *p : ctx->owned) cudaFree(p);
  delete ctx;
}

DIM_API int32_t dim_mesh_upload(dim_ctx *ctx, int32_t cls, const float *verts, const float *uvs, int32_t V,
                                const int32_t *faces, int32_t F, const uint8_t *tex, int32_t Th, int32_t Tw) {
  DIM_REQUIRE(ctx && cls >= 0 && cls < ctx->max_classes, "dim_mesh_upload: bad class index");
  DIM_REQUIRE(V > 0 && V <= ctx->max_verts && F > 0 && F <= ctx->max_faces, "dim_mesh_upload: mesh exceeds ctx limits");
  drop_graphs(ctx);  // the render launch grid follows the largest uploaded mesh
  for (int32_t i = 0; i < 3 * F; ++i) DIM_REQUIRE(faces[i] >= 0 && faces[i] < V, "dim_mesh_upload: face index out of range");
  MeshDev m;
  float *dv, *du; int *df; uint8_t *dt;
  if (ctx_alloc(ctx, &dv, (size_t)3 * V) || ctx_alloc(ctx, &du, (size_t)2 * V) || ctx_alloc(ctx, &df, (size_t)3 * F) ||
      ctx_alloc(ctx, &dt, (size_t)3 * Th * Tw))
    return 12;
  DIM_CHECK(cudaMemcpy(dv, verts, sizeof(float) * 3 * V, cudaMemcpyHostToDevice));
  DIM_CHECK(cudaMemcpy(du, uvs, sizeof(float) * 2 * V, cudaMemcpyHostToDevice));
  DIM_CHECK(cudaMemcpy(df, faces, sizeof(int) * 3 * F, cudaMemcpyHostToDevice));
  DIM_CHECK(cudaMemcpy(dt, tex, (size_t)3 * Th * Tw, cudaMemcpyHostToDevice));
  m.verts = dv; m.uvs = du; m.faces = df; m.tex = dt; m.V = V; m.F = F; m.Th = Th; m.Tw = Tw; m.normals = nullptr;
  ctx->meshes_host[cls] = m;
  DIM_CHECK(cudaMemcpy(ctx->meshes + cls, &m, sizeof(MeshDev), cudaMemcpyHostToDevice));
  return 0;
}

DIM_API int32_t dim_render(dim_ctx *ctx, const int32_t *cls_idx, const float *pose, int32_t B, const float *K9,
                           float zn, float zf, const double *means, int32_t trunc_u8, float *out_image,
                           float *out_depth, float *out_mask, float *out_bgr, int32_t *out_bbox, void *stream) {
  DIM_REQUIRE(ctx && cls_idx && pose && K9, "dim_render: NULL argument");
  return render_launch(ctx, cls_idx, pose, B, K9, zn, zf, means, trunc_u8, out_image, out_depth, out_mask, out_bgr,
                       out_bbox, nullptr, (cudaStream_t)stream);
}

// lit renderer (lib/render_glumpy/render_py_light_modelnet_multi.py)
DIM_API int32_t dim_mesh_upload_normals(dim_ctx *ctx, int32_t cls, const float *normals, int32_t V) {
  DIM_REQUIRE(ctx && normals && cls >= 0 && cls < ctx->max_classes, "dim_mesh_upload_normals: bad argument");
  MeshDev &m = ctx->meshes_host[cls];
  DIM_REQUIRE(m.V > 0 && m.V == V, "dim_mesh_upload_normals: upload the mesh first; V must match");
  float *dn;
  if (ctx_alloc(ctx, &dn, (size_t)3 * V)) return 12;
  DIM_CHECK(cudaMemcpy(dn, normals, sizeof(float) * 3 * V, cudaMemcpyHostToDevice));
  m.normals = dn;
  DIM_CHECK(cudaMemcpy(ctx->meshes + cls, &m, sizeof(MeshDev), cudaMemcpyHostToDevice));
  return 0;
}
DIM_API int32_t dim_render_lit(dim_ctx *ctx, const int32_t *cls_idx, const float *pose, int32_t B, const float *K9, float zn,
                               float zf, const double *means, const float *light_pos, const float *light_int,
                               float brightness_ratio, float *out_image, float *out_depth, float *out_mask, float *out_bgr,
                               int32_t *out_bbox, void *stream) {
  DIM_REQUIRE(ctx && cls_idx && pose && K9 && light_pos && light_int, "dim_render_lit: NULL argument");
  for (auto &m : ctx->meshes_host) DIM_REQUIRE(m.V == 0 || m.normals != nullptr, "dim_render_lit: a mesh has no normals (dim_mesh_upload_normals)");
  LitParams lit{light_pos, light_int, (float)(1.0 - (double)brightness_ratio), brightness_ratio};
  return render_launch(ctx, cls_idx, pose, B, K9, zn, zf, means, 1, out_image, out_depth, out_mask, out_bgr, out_bbox, nullptr,
                       (cudaStream_t)stream, &lit);
}

DIM_API int32_t dim_zoom_mask_fwd(dim_ctx *ctx, const float *mo, const float *mgt, const float *mr,
                                  const float *src_pose, int32_t B, const float *K9, float *zmo, float *zmgt,
                                  float *zmr, float *zoom_factor, int32_t *bbox, int32_t *status, void *stream) {
  DIM_REQUIRE(ctx && mo && mgt && mr && src_pose && K9 && zoom_factor, "dim_zoom_mask_fwd: NULL argument");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "dim_zoom_mask_fwd: batch exceeds max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = zoom_factor_launch(ctx, mgt, mr, 1, src_pose, B, K9, zoom_factor, bbox, status, st)) return rc;
  if (zmo) if (int rc = zoom_gather_launch(ctx, 1, mo, zmo, zoom_factor, B, 1, 0, nullptr, st)) return rc;
  if (zmgt) if (int rc = zoom_gather_launch(ctx, 1, mgt, zmgt, zoom_factor, B, 1, 0, nullptr, st)) return rc;
  // rendered mask is binarised at 0.2 before sampling (zoom_mask.py:39-42)
  if (zmr) if (int rc = zoom_gather_launch(ctx, 2, mr, zmr, zoom_factor, B, 1, 0, nullptr, st)) return rc;
  return 0;
}

DIM_API int32_t dim_zoom_image_with_factor_fwd(dim_ctx *ctx, const float *zoom_factor, const float *io,
                                               const float *ir, int32_t B, const float *means, float *zio,
                                               float *zir, void *stream) {
  DIM_REQUIRE(ctx && zoom_factor && io && ir && means && zio && zir, "dim_zoom_image_with_factor_fwd: NULL argument");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "batch exceeds max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = zoom_gather_launch(ctx, 3, io, zio, zoom_factor, B, 3, 0, means, st)) return rc;
  return zoom_gather_launch(ctx, 3, ir, zir, zoom_factor, B, 3, 0, means, st);
}

// ZoomImage (zoom_image.py:26-107): bbox from the images themselves (INPUT_MASK: False graphs)
DIM_API int32_t dim_zoom_image_fwd(dim_ctx *ctx, const float *io, const float *ir, const float *src_pose, int32_t B,
                                   const float *K9, const float *means, float *zio, float *zir, float *zoom_factor,
                                   int32_t *bbox, int32_t *status, void *stream) {
  DIM_REQUIRE(ctx && io && ir && src_pose && K9 && means && zio && zir && zoom_factor, "dim_zoom_image_fwd: NULL argument");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "dim_zoom_image_fwd: batch exceeds max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = zoom_factor_launch(ctx, io, ir, 3, src_pose, B, K9, zoom_factor, bbox, status, st, means)) return rc;
  if (int rc = zoom_gather_launch(ctx, 3, io, zio, zoom_factor, B, 3, 0, means, st)) return rc;
  return zoom_gather_launch(ctx, 3, ir, zir, zoom_factor, B, 3, 0, means, st);
}

DIM_API int32_t dim_group_picker(dim_ctx *ctx, const float *in, const float *group_idx, int32_t B, int32_t channels,
                                 int32_t group_num, int64_t elems_per_channel, int32_t backward, float *out, void *stream) {
  DIM_REQUIRE(ctx && in && group_idx && out, "dim_group_picker: NULL argument");
  DIM_REQUIRE(group_num >= 1 && channels % group_num == 0 && elems_per_channel >= 1, "dim_group_picker: channels must divide into groups");
  return group_pick_launch(in, group_idx, B, channels, group_num, (size_t)elems_per_channel, out, backward, (cudaStream_t)stream);
}

DIM_API int32_t dim_zoom_mask_with_factor_fwd(dim_ctx *ctx, const float *zoom_factor, const float *mask, int32_t B,
                                              int32_t inv, float *out, void *stream) {
  DIM_REQUIRE(ctx && zoom_factor && mask && out, "dim_zoom_mask_with_factor_fwd: NULL argument");
  return zoom_gather_launch(ctx, 2, mask, out, zoom_factor, B, 1, inv, nullptr, (cudaStream_t)stream);
}

DIM_API int32_t dim_zoom_flow_fwd(dim_ctx *ctx, const float *zoom_factor, const float *flow, const float *fw,
                                  int32_t fw_channels, int32_t B, int32_t inv, float *zflow, float *zfw, void *stream) {
  DIM_REQUIRE(ctx && zoom_factor && flow && zflow, "dim_zoom_flow_fwd: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = zoom_gather_launch(ctx, inv ? 4 : 6, flow, zflow, zoom_factor, B, 2, inv, nullptr, st)) return rc;
  if (!inv && fw && zfw) {
    DIM_REQUIRE(fw_channels == 1 || fw_channels == 2, "dim_zoom_flow_fwd: flow_weights has 1 or 2 channels");
    return zoom_gather_launch(ctx, 5, fw, zfw, zoom_factor, B, fw_channels, 0, nullptr, st);
  }
  return 0;
}

DIM_API int32_t dim_zoom_depth_fwd(dim_ctx *ctx, const float *zoom_factor, const float *dobs, const float *dren,
                                   int32_t B, float *zobs, float *zren, void *stream) {
  DIM_REQUIRE(ctx && zoom_factor && dobs && dren && zobs && zren, "dim_zoom_depth_fwd: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = zoom_gather_launch(ctx, 0, dobs, zobs, zoom_factor, B, 1, 0, nullptr, st)) return rc;
  return zoom_gather_launch(ctx, 0, dren, zren, zoom_factor, B, 1, 0, nullptr, st);
}

DIM_API int32_t dim_zoom_trans_fwd(dim_ctx *ctx, const float *zoom_factor, const float *trans, int32_t B, int32_t inv,
                                   float *out, void *stream) {
  DIM_REQUIRE(ctx && zoom_factor && trans && out, "dim_zoom_trans_fwd: NULL argument");
  return zoom_trans_launch(zoom_factor, trans, B, inv ? 1 : 0, 1, out, (cudaStream_t)stream);
}
DIM_API int32_t dim_zoom_trans_bwd(dim_ctx *ctx, const float *zoom_factor, const float *og, int32_t B, int32_t inv,
                                   int32_t zoom_grad, float *out, void *stream) {
  DIM_REQUIRE(ctx && zoom_factor && og && out, "dim_zoom_trans_bwd: NULL argument");
  return zoom_trans_launch(zoom_factor, og, B, inv ? 1 : 0, zoom_grad ? 1 : 0, out, (cudaStream_t)stream);
}

DIM_API int32_t dim_update_mask_box(dim_ctx *ctx, const int32_t *bbox, int32_t B, float *mask, void *stream) {
  DIM_REQUIRE(ctx && bbox && mask, "dim_update_mask_box: NULL argument");
  return box_mask_launch(ctx, bbox, B, mask, (cudaStream_t)stream);
}

DIM_API int32_t dim_se3_compose(dim_ctx *ctx, const double *pose_src, const float *se3, int32_t B, const double *Tm,
                                const double *Ts, int32_t rot_coord, double *pose_out, void *stream) {
  DIM_REQUIRE(ctx && pose_src && se3 && Tm && Ts && pose_out, "dim_se3_compose: NULL argument");
  DIM_REQUIRE(rot_coord >= 0 && rot_coord <= 2, "dim_se3_compose: unknown rot_coord");
  return se3_compose_launch(pose_src, se3, B, Tm, Ts, rot_coord, pose_out, nullptr, (cudaStream_t)stream);
}

DIM_API int32_t dim_flow_fwd(dim_ctx *ctx, const float *ds, const float *dt, const float *KT, const float *Kinv,
                             int32_t B, float *flow, float *valid, void *stream) {
  DIM_REQUIRE(ctx && ds && dt && KT && Kinv && flow && valid, "dim_flow_fwd: NULL argument");
  return flow_launch(ctx, ds, dt, KT, Kinv, B, flow, valid, nullptr, (cudaStream_t)stream);
}

DIM_API int32_t dim_transform3d_fwd(dim_ctx *ctx, const float *pc, const float *rot, const float *tr,
                                    const float *ps, int32_t B, int32_t N, const float *Tm, const float *Ts,
                                    int32_t rot_coord, float *out, void *stream) {
  DIM_REQUIRE(ctx && pc && rot && tr && ps && Tm && Ts && out, "dim_transform3d_fwd: NULL argument");
  return transform3d_fwd_launch(pc, rot, tr, ps, B, N, Tm, Ts, rot_coord, out, (cudaStream_t)stream);
}
DIM_API int32_t dim_transform3d_bwd(dim_ctx *ctx, const float *og, const float *pc, const float *rot, const float *tr,
                                    const float *ps, int32_t B, int32_t N, const float *Tm, const float *Ts,
                                    int32_t rot_coord, float *rg, float *tg, void *stream) {
  DIM_REQUIRE(ctx && og && pc && rot && tr && ps && Tm && Ts && rg && tg, "dim_transform3d_bwd: NULL argument");
  return transform3d_bwd_launch(og, pc, rot, tr, ps, B, N, Tm, Ts, rot_coord, rg, tg, (cudaStream_t)stream);
}

DIM_API int32_t dim_net_load(dim_ctx *ctx, const float *const *W, const float *const *Bv) {
  DIM_REQUIRE(ctx && W && Bv, "dim_net_load: NULL argument");
  drop_graphs(ctx);
  return net_load(ctx, W, Bv);
}

DIM_API int32_t dim_net_fwd(dim_ctx *ctx, const float *zio, const float *zir, const float *zmo, const float *zmr,
                            int32_t B, int32_t precision, float *rot, float *trans, void *stream) {
  DIM_REQUIRE(ctx && zio && zir && zmo && zmr && rot && trans, "dim_net_fwd: NULL argument");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "dim_net_fwd: batch exceeds max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  int rows, cols, pad; __nv_bfloat16 *hi, *lo;
  net_input_geometry(ctx, &rows, &cols, &pad, &hi, &lo);
  if (int rc = pack_nhwc8_launch(ctx, zio, zir, zmo, zmr, B, rows, cols, pad, hi,
                                 precision == DIM_PREC_BF16X3 ? lo : nullptr, st, precision == DIM_PREC_FP16))
    return rc;
  return net_forward(ctx, B, precision, nullptr, rot, trans, nullptr, st, nullptr);
}

DIM_API int32_t dim_transform_image_u8(dim_ctx *ctx, const uint8_t *bgr, int32_t B, const double *means, float *image,
                                       void *stream) {
  DIM_REQUIRE(ctx && bgr && means && image, "dim_transform_image_u8: NULL argument");
  return transform_u8_launch(ctx, bgr, B, means, image, (cudaStream_t)stream);
}

static int refine_core(dim_ctx *ctx, const float4 *obs4, const int32_t *cls_idx, const double *pose_init, int32_t B,
                       int32_t n_iter, const float *K9, float zn, float zf, const double *means, int32_t precision,
                       const double *pose_override, double *poses, float *se3, float *zoom_factor, int32_t *bbox,
                       cudaStream_t st) {
  const double Tm[3] = {ctx->cfg.trans_means[0], ctx->cfg.trans_means[1], ctx->cfg.trans_means[2]};
  const double Ts[3] = {ctx->cfg.trans_stds[0], ctx->cfg.trans_stds[1], ctx->cfg.trans_stds[2]};
  const float means_f[3] = {(float)means[0], (float)means[1], (float)means[2]};
  int rows, cols, pad; __nv_bfloat16 *hi, *lo;
  net_input_geometry(ctx, &rows, &cols, &pad, &hi, &lo);
  const double *pose_src = pose_init;
  for (int it = 0; it < n_iter; ++it) {
    DimNvtxRange r_it("dim_refine iteration");
    cudaEvent_t *ev = nullptr;
    if (ctx->prof) {
      while (ctx->prof_events.size() < ctx->prof_used + 5) {
        cudaEvent_t e;
        DIM_CHECK(cudaEventCreate(&e));
        ctx->prof_events.push_back(e);
      }
      ev = &ctx->prof_events[ctx->prof_used];
      ctx->prof_used += 5;
      DIM_CHECK(cudaEventRecord(ev[0], st));
    }
    if (pose_override) pose_src = pose_override + (size_t)it * B * 12;
    // src_pose blob is float32 (nd.array), the host pose stays float64 (tester.py:391)
    if (int rc = f64_to_f32_launch(pose_src, ctx->pose_cur_f32, B * 12, st)) return rc;
    // render at the current pose (tester.py:427-442) straight into the pixel-interleaved
    // (R,G,B,mask) image the zoom kernel samples; mask_observed := box(mask_rendered) is analytic
    {
      DimNvtxRange r("render");
      if (int rc = render_launch(ctx, cls_idx, ctx->pose_cur_f32, B, K9, zn, zf, means, 1, nullptr, nullptr, nullptr,
                                 nullptr, nullptr, ctx->ren4, st))
        return rc;
    }
    if (ev) DIM_CHECK(cudaEventRecord(ev[1], st));
    float *zf_it = zoom_factor ? zoom_factor + (size_t)it * B * 4 : ctx->zoom_factor;
    int *bbox_it = bbox ? bbox + (size_t)it * B * 8 : nullptr;
    // per-iteration status (bit 0: rendered / observed mask empty -> fallback zoom factor; bit 1: bad class index)
    {
      DimNvtxRange r("bbox + zoom");
      if (int rc = zoom_factor_from_ren_launch(ctx, ctx->bbox_ren, ctx->pose_cur_f32, B, K9, zf_it, bbox_it,
                                               ctx->status_hist + (size_t)(it < 8 ? it : 7) * B, st))
        return rc;
      if (int rc = zoom_fused_launch(ctx, obs4, ctx->ren4, zf_it, means_f, B, rows, cols, pad, hi,
                                     precision == DIM_PREC_BF16X3 ? lo : nullptr, st, precision == DIM_PREC_FP16, means))
        return rc;
    }
    if (ev) DIM_CHECK(cudaEventRecord(ev[2], st));
    float *se3_it = se3 ? se3 + (size_t)it * B * 7 : ctx->se3_cur;
    {
      DimNvtxRange r("FlowNetS tower + heads");
      if (int rc = net_forward(ctx, B, precision, zf_it, nullptr, nullptr, se3_it, st, ev ? ev[3] : nullptr)) return rc;
    }
    double *pose_out = poses + (size_t)it * B * 12;
    {
      DimNvtxRange r("SE(3) compose");
      if (int rc = se3_compose_launch(pose_src, se3_it, B, Tm, Ts, ctx->cfg.rot_coord, pose_out, nullptr, st)) return rc;
    }
    if (ev) DIM_CHECK(cudaEventRecord(ev[4], st));
    pose_src = pose_out;
  }
  return 0;
}

// The 4-iteration chain is ~90 kernel launches whose arguments do not change from call to call when the caller reuses
// its buffers (PoseRefiner does; so does bench.py).  After one eager run of an argument set the chain is captured into a
// CUDA graph (stream capture, thread-local mode) and replayed with one cudaGraphLaunch: the data-dependent parts of the
// loop are already branch-free on the device.  Disabled while stage / layer profiling is on or the context trains.
static int refine_graphed(dim_ctx *ctx, const float4 *obs4, const int32_t *cls_idx, const double *pose_init, int32_t B,
                          int32_t n_iter, const float *K9, float zn, float zf, const double *means, int32_t precision,
                          const double *pose_override, double *poses, float *se3, float *zoom_factor, int32_t *bbox,
                          cudaStream_t st) {
  // the legacy default stream (and the per-thread default stream handle) cannot be captured
  const bool capturable = st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
  if (!ctx->use_graph || ctx->prof || !capturable || !net_graph_safe(ctx))
    return refine_core(ctx, obs4, cls_idx, pose_init, B, n_iter, K9, zn, zf, means, precision, pose_override, poses, se3,
                       zoom_factor, bbox, st);
  std::vector<unsigned char> key;
  auto put = [&key](const void *p, size_t n) { key.insert(key.end(), (const unsigned char *)p, (const unsigned char *)p + n); };
  const void *ptrs[8] = {obs4, cls_idx, pose_init, pose_override, poses, se3, zoom_factor, bbox};
  const int32_t ints[3] = {B, n_iter, precision};
  put(ptrs, sizeof(ptrs)); put(ints, sizeof(ints)); put(K9, 9 * sizeof(float)); put(&zn, sizeof(zn)); put(&zf, sizeof(zf));
  put(means, 3 * sizeof(double));
  dim_ctx::RefineGraph *g = nullptr;
  for (auto &e : ctx->graphs)
    if (e.key == key) { g = &e; break; }
  if (g && g->exec) {
    DIM_CHECK(cudaGraphLaunch(g->exec, st));
    g_launches += g->kernels;
    return 0;
  }
  if (!g) {  // first sight of this argument set: eager run (also builds tensor maps, sets function attributes)
    if (ctx->graphs.size() >= 32) ctx->graphs.erase(ctx->graphs.begin());
    ctx->graphs.push_back(dim_ctx::RefineGraph{key, nullptr, 0});
    return refine_core(ctx, obs4, cls_idx, pose_init, B, n_iter, K9, zn, zf, means, precision, pose_override, poses, se3,
                       zoom_factor, bbox, st);
  }
  const long long before = g_launches;
  DIM_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = refine_core(ctx, obs4, cls_idx, pose_init, B, n_iter, K9, zn, zf, means, precision, pose_override, poses,
                             se3, zoom_factor, bbox, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == 0) set_error("dim_refine: stream capture failed: %s", cudaGetErrorString(ce));
    cudaGetLastError();
    return rc ? rc : 1;
  }
  const long long kernels = g_launches - before;
  g_launches = before;
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) { set_error("dim_refine: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); return 1; }
  g->exec = exec;
  g->kernels = kernels;
  DIM_CHECK(cudaGraphLaunch(exec, st));
  g_launches += kernels;
  return 0;
}

DIM_API int32_t dim_refine(dim_ctx *ctx, const float *image_observed, const int32_t *cls_idx, const double *pose_init,
                           int32_t B, int32_t n_iter, const float *K9, float zn, float zf, const double *means,
                           int32_t precision, const double *pose_override, double *poses, float *se3,
                           float *zoom_factor, int32_t *bbox, void *stream) {
  DIM_REQUIRE(ctx && image_observed && cls_idx && pose_init && K9 && means && poses, "dim_refine: NULL argument");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "dim_refine: batch exceeds max_batch");
  DIM_REQUIRE(n_iter >= 1, "dim_refine: n_iter must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = pack_obs4_launch(ctx, image_observed, B, ctx->obs4, means, st)) return rc;
  return refine_graphed(ctx, ctx->obs4, cls_idx, pose_init, B, n_iter, K9, zn, zf, means, precision, pose_override, poses,
                        se3, zoom_factor, bbox, st);
}

DIM_API int32_t dim_refine_host_async(dim_ctx *ctx, const uint8_t *img_u8, const int32_t *cls_host,
                                      const double *pose_host, int32_t B, int32_t n_iter, const float *K9, float zn,
                                      float zf, const double *means, int32_t precision, double *poses_out,
                                      float *se3_out, void *stream) {
  DIM_REQUIRE(ctx && img_u8 && cls_host && pose_host && K9 && means && poses_out, "dim_refine_host: NULL argument");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "dim_refine_host: batch exceeds max_batch");
  DIM_REQUIRE(n_iter >= 1 && n_iter <= 8, "dim_refine_host: n_iter must be in [1,8]");
  for (int32_t i = 0; i < B; ++i) {  // the class indices are on the host here: fail loudly (the reference indexes a python list)
    const int32_t c = cls_host[i];
    if (c < 0 || c >= ctx->max_classes || ctx->meshes_host[c].V <= 0) {
      set_error("dim_refine_host: instance %d has class index %d: out of range [0,%d) or no mesh uploaded for it", (int)i, (int)c,
                (int)ctx->max_classes);
      return 2;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P = (size_t)ctx->H * ctx->W;
  DIM_CHECK(cudaMemcpyAsync(ctx->image_observed_u8, img_u8, (size_t)B * 3 * P, cudaMemcpyHostToDevice, st));
  DIM_CHECK(cudaMemcpyAsync(ctx->cls_dev, cls_host, sizeof(int) * B, cudaMemcpyHostToDevice, st));
  DIM_CHECK(cudaMemcpyAsync(ctx->pose_cur, pose_host, sizeof(double) * B * 12, cudaMemcpyHostToDevice, st));
  if (int rc = transform_u8_obs4_launch(ctx, ctx->image_observed_u8, B, means, ctx->obs4, st)) return rc;
  if (int rc = refine_graphed(ctx, ctx->obs4, ctx->cls_dev, ctx->pose_cur, B, n_iter, K9, zn, zf, means, precision,
                              nullptr, ctx->poses_dev, ctx->se3_hist_dev, nullptr, nullptr, st))
    return rc;
  DIM_CHECK(cudaMemcpyAsync(poses_out, ctx->poses_dev, sizeof(double) * (size_t)n_iter * B * 12, cudaMemcpyDeviceToHost, st));
  if (se3_out)
    DIM_CHECK(cudaMemcpyAsync(se3_out, ctx->se3_hist_dev, sizeof(float) * (size_t)n_iter * B * 7, cudaMemcpyDeviceToHost, st));
  return 0;
}

// status of the LAST dim_refine / dim_refine_host(_async) call on this context: [min(n_iter, 8), B] int32, device -> host
// (asynchronous on `stream`, which must be the stream that call ran on).  0 = ok; bit 0 = the rendered mask of that
// iteration was empty (object left the frustum: the reference crashes in ZoomMask, np.min of an empty array; here the
// fallback zoom factor (1,1,0,0) was used and the pose of that instance is meaningless); bit 1 = class index out of range
// or no mesh uploaded for it.
DIM_API int32_t dim_refine_status(dim_ctx *ctx, int32_t B, int32_t n_iter, int32_t *status_host, void *stream) {
  DIM_REQUIRE(ctx && status_host && B >= 1 && B <= ctx->max_batch && n_iter >= 1, "dim_refine_status: bad argument");
  const int n = n_iter < 8 ? n_iter : 8;
  DIM_CHECK(cudaMemcpyAsync(status_host, ctx->status_hist, sizeof(int) * (size_t)n * B, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}

DIM_API int32_t dim_refine_host(dim_ctx *ctx, const uint8_t *img_u8, const int32_t *cls_host, const double *pose_host,
                                int32_t B, int32_t n_iter, const float *K9, float zn, float zf, const double *means,
                                int32_t precision, double *poses_out, float *se3_out, void *stream) {
  if (int rc = dim_refine_host_async(ctx, img_u8, cls_host, pose_host, B, n_iter, K9, zn, zf, means, precision, poses_out,
                                     se3_out, stream))
    return rc;
  DIM_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

DIM_API int32_t dim_train_update(dim_ctx *ctx, const int32_t *cls_idx, const float *src_pose, const float *rot_est,
                                 const float *trans_est, const float *tgt_pose, const float *depth_gt_observed,
                                 int32_t B, const double *K9, float zn, float zf, const double *means,
                                 const double *Tm, const double *Ts, int32_t rot_coord, float *image_rendered,
                                 float *depth_rendered, float *mask_rendered, float *src_pose_new, float *rot_label,
                                 float *trans_label, float *flow, float *flow_weights, void *stream) {
  DIM_REQUIRE(ctx && cls_idx && src_pose && rot_est && trans_est && tgt_pose && K9 && means && Tm && Ts,
              "dim_train_update: NULL argument");
  DIM_REQUIRE(image_rendered && depth_rendered && mask_rendered && src_pose_new && rot_label && trans_label,
              "dim_train_update: NULL output");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "dim_train_update: batch exceeds max_batch");
  DIM_REQUIRE(rot_coord >= 0 && rot_coord <= 2, "dim_train_update: unknown rot_coord");
  cudaStream_t st = (cudaStream_t)stream;
  float *KT = ctx->pose_cur_f32;  // [B,12] scratch
  if (int rc = train_pose_launch(src_pose, rot_est, trans_est, tgt_pose, B, Tm, Ts, rot_coord, K9, src_pose_new,
                                 rot_label, trans_label, KT, st))
    return rc;
  const float K9f[9] = {(float)K9[0], (float)K9[1], (float)K9[2], (float)K9[3], (float)K9[4],
                        (float)K9[5], (float)K9[6], (float)K9[7], (float)K9[8]};
  // no uint8 truncation on the train path (batch_updater_py_multi.py:184,234)
  if (int rc = render_launch(ctx, cls_idx, src_pose_new, B, K9f, zn, zf, means, 0, image_rendered, depth_rendered,
                             mask_rendered, nullptr, nullptr, nullptr, st))
    return rc;
  if (flow && flow_weights) {
    DIM_REQUIRE(depth_gt_observed != nullptr, "dim_train_update: flow labels need depth_gt_observed");
    // Kinv = inverse of K (np.linalg.inv, l.60) in float64, cast to float32 (l.292)
    const double a = K9[0], b_ = K9[1], c = K9[2], d = K9[3], e = K9[4], f = K9[5], g = K9[6], h = K9[7], i = K9[8];
    const double det = a * (e * i - f * h) - b_ * (d * i - f * g) + c * (d * h - e * g);
    const float Kinv[9] = {(float)((e * i - f * h) / det), (float)((c * h - b_ * i) / det), (float)((b_ * f - c * e) / det),
                           (float)((f * g - d * i) / det), (float)((a * i - c * g) / det), (float)((c * d - a * f) / det),
                           (float)((d * h - e * g) / det), (float)((b_ * g - a * h) / det), (float)((a * e - b_ * d) / det)};
    const size_t P = (size_t)ctx->H * ctx->W;
    // flow_weights = tile(valid, [1,2,1,1]) (l.296): the kernel writes both copies
    if (int rc = flow_launch(ctx, depth_rendered, depth_gt_observed, KT, Kinv, B, flow, ctx->mask_rendered,
                             nullptr, st))
      return rc;
    // interleave valid into the two weight planes per instance
    for (int pl = 0; pl < 2; ++pl)
      DIM_CHECK(cudaMemcpy2DAsync(flow_weights + pl * P, 2 * P * sizeof(float), ctx->mask_rendered, P * sizeof(float),
                                  P * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

DIM_API int32_t dim_profile_enable(dim_ctx *ctx, int32_t enable) {
  DIM_REQUIRE(ctx, "dim_profile_enable: NULL ctx");
  DIM_CHECK(cudaDeviceSynchronize());
  ctx->prof = enable != 0;
  ctx->prof_used = 0;
  return 0;
}
DIM_API int32_t dim_profile_read(dim_ctx *ctx, float *ms4, int32_t *iterations) {
  DIM_REQUIRE(ctx && ms4, "dim_profile_read: NULL argument");
  DIM_CHECK(cudaDeviceSynchronize());
  for (int k = 0; k < 4; ++k) ms4[k] = 0.f;
  const size_t n = ctx->prof_used / 5;
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) {
      float ms = 0.f;
      DIM_CHECK(cudaEventElapsedTime(&ms, ctx->prof_events[5 * i + k], ctx->prof_events[5 * i + k + 1]));
      ms4[k] += ms;
    }
  if (iterations) *iterations = (int32_t)n;
  ctx->prof_used = 0;
  return 0;
}

// ---- test hooks (not part of the drop-in surface; used by tests/ to look inside the conv tower)
DIM_API int32_t dim_debug_activation(dim_ctx *ctx, int32_t idx, int32_t lo, void *host_dst, uint64_t bytes) {
  return net_debug_activation(ctx, idx, lo, host_dst, (size_t)bytes);
}
DIM_API int32_t dim_debug_layer_geometry(dim_ctx *ctx, int32_t idx, int32_t *out8) {
  DIM_REQUIRE(ctx && out8 && idx >= 0 && idx <= 10, "dim_debug_layer_geometry: bad argument");
  net_layer_geometry(ctx, idx, out8);
  return 0;
}


DIM_API int32_t dim_debug_set_option(dim_ctx *ctx, const char *key, int32_t value) {
  DIM_REQUIRE(ctx && key, "dim_debug_set_option: NULL argument");
  drop_graphs(ctx);  // captured graphs hold the old kernel choice
  if (!strcmp(key, "graph")) { ctx->use_graph = value != 0; return 0; }
  return net_set_option(ctx, key, value);
}
DIM_API int32_t dim_debug_layer_profile(dim_ctx *ctx, int32_t enable, float *ms10) {
  DIM_REQUIRE(ctx, "dim_debug_layer_profile: NULL ctx");
  return net_layer_profile(ctx, enable, ms10);
}

// ADD / ADI (lib/utils/pose_error.py:72-108)
DIM_API int32_t dim_pose_error(dim_ctx *ctx, const double *poses_est, const double *poses_gt, int32_t M, const double *points,
                               int32_t N, int32_t symmetric, double *err, void *stream) {
  DIM_REQUIRE(ctx && poses_est && poses_gt && points && err && M >= 1 && N >= 1, "dim_pose_error: bad argument");
  return pose_error_launch(poses_est, poses_gt, M, points, N, symmetric, err, (cudaStream_t)stream);
}

// flow end-point error sums (deepim/core/tester.py:573-589)
DIM_API int32_t dim_flow_epe(dim_ctx *ctx, const float *flow_pred, const float *flow_gt, const float *visible, const float *bg,
                             int32_t B, double *out6, void *stream) {
  DIM_REQUIRE(ctx && flow_pred && flow_gt && visible && bg && out6 && B >= 1, "dim_flow_epe: bad argument");
  return epe_launch(flow_pred, flow_gt, visible, bg, B, ctx->H * ctx->W, out6, (cudaStream_t)stream);
}

// Proj. 2D / (rot, trans) distances (lib/utils/pose_error.py:55-69, lib/pair_matching/RT_transform.py:162-173)
DIM_API int32_t dim_pose_error_2d(dim_ctx *ctx, const double *poses_est, const double *poses_gt, int32_t M, const double *points,
                                  int32_t N, const double *K9_dev, double *err3, void *stream) {
  DIM_REQUIRE(ctx && poses_est && poses_gt && points && K9_dev && err3 && M >= 1 && N >= 1, "dim_pose_error_2d: bad argument");
  return pose_error2d_launch(poses_est, poses_gt, M, points, N, K9_dev, err3, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ training step
DIM_API int32_t dim_train_create(dim_ctx *ctx, int32_t max_points) {
  DIM_REQUIRE(ctx && max_points >= 1, "dim_train_create: bad argument");
  return train_create(ctx, max_points);
}
DIM_API int64_t dim_train_param_count(dim_ctx *ctx) { return ctx ? (int64_t)train_param_count(ctx) : 0; }
DIM_API int32_t dim_train_param_info(int32_t idx, const char **name, int64_t *w_numel, int64_t *b_numel) {
  long long w = 0, b = 0;
  DIM_REQUIRE(name && w_numel && b_numel, "dim_train_param_info: NULL argument");
  if (train_param_info(idx, name, &w, &b)) return 2;
  *w_numel = w; *b_numel = b;
  return 0;
}
DIM_API int32_t dim_train_load_params(dim_ctx *ctx, const float *flat_host, int64_t n, void *stream) {
  DIM_REQUIRE(ctx && flat_host && n > 0, "dim_train_load_params: bad argument");
  return train_load_params(ctx, flat_host, (size_t)n, (cudaStream_t)stream);
}
DIM_API int32_t dim_train_get_params(dim_ctx *ctx, float *flat_host, int64_t n, int32_t which, void *stream) {
  DIM_REQUIRE(ctx && flat_host && n > 0, "dim_train_get_params: bad argument");
  return train_get_params(ctx, flat_host, (size_t)n, which, (cudaStream_t)stream);
}
DIM_API int32_t dim_train_forward_backward(dim_ctx *ctx, const float *zio, const float *zir, const float *zmo, const float *zmr,
                                           const float *zoom_factor, const float *zflow, const float *zfw, const float *zmask_gt,
                                           const float *src_pose, const float *pc_model, const float *pc_weights,
                                           const float *pc_observed, int32_t B, int32_t N, float *rot_est_norm, float *trans_est,
                                           float *flow_est, float *mask_prob, float *losses4, float *grads, float *rot_raw,
                                           void *const *bucket_events, const int32_t *bucket_first_tensor, int32_t n_buckets,
                                           void *stream) {
  DIM_REQUIRE(ctx && zio && zir && zmo && zmr && zoom_factor, "dim_train_forward_backward: NULL argument");
  TrainIO io{zio, zir, zmo, zmr, zoom_factor, zflow, zfw, zmask_gt, src_pose, pc_model, pc_weights, pc_observed, B, N,
             rot_est_norm, trans_est, flow_est, mask_prob, losses4, grads, rot_raw, bucket_events, bucket_first_tensor,
             (bucket_events && bucket_first_tensor) ? n_buckets : 0};
  return train_forward_backward(ctx, io, (cudaStream_t)stream);
}
DIM_API int32_t dim_train_set_config(dim_ctx *ctx, const dim_train_config *cfg) {
  DIM_REQUIRE(ctx && cfg, "dim_train_set_config: NULL argument");
  DIM_REQUIRE(cfg->rot_coord == 0 || cfg->rot_coord == 1, "dim_train_set_config: rot_coord must be 0 (MODEL) or 1 (CAMERA)");
  DIM_REQUIRE(cfg->num_3d_sample > 0.f && cfg->normalize_3d_point > 0.f && cfg->normalize_flow > 0.f,
              "dim_train_set_config: num_3d_sample, normalize_3d_point and normalize_flow must be positive");
  for (int i = 0; i < 3; ++i) DIM_REQUIRE(cfg->trans_stds[i] != 0.f, "dim_train_set_config: trans_stds must be non-zero");
  ctx->cfg = *cfg;
  drop_graphs(ctx);  // the captured refinement chains carry trans_means / trans_stds / rot_coord as kernel arguments
  return 0;
}
DIM_API int32_t dim_train_get_config(dim_ctx *ctx, dim_train_config *cfg) {
  DIM_REQUIRE(ctx && cfg, "dim_train_get_config: NULL argument");
  *cfg = ctx->cfg;
  return 0;
}
DIM_API int32_t dim_train_sgd_update(dim_ctx *ctx, const float *grads, float lr, float momentum, float wd, float rescale_grad,
                                     void *stream) {
  DIM_REQUIRE(ctx && grads, "dim_train_sgd_update: NULL argument");
  return train_sgd_update(ctx, grads, lr, momentum, wd, rescale_grad, (cudaStream_t)stream);
}
DIM_API int32_t dim_train_debug_tensor(dim_ctx *ctx, int32_t id, void *host_dst, uint64_t bytes) {
  DIM_REQUIRE(ctx && host_dst, "dim_train_debug_tensor: NULL argument");
  return train_debug_tensor(ctx, id, host_dst, (size_t)bytes);
}
DIM_API int32_t dim_train_debug_phases(dim_ctx *ctx, float *ms7) {
  DIM_REQUIRE(ctx && ms7, "dim_train_debug_phases: NULL argument");
  return train_debug_phases(ctx, ms7);
}
DIM_API int32_t dim_train_debug_geometry(dim_ctx *ctx, int32_t id, int32_t *out7) {
  DIM_REQUIRE(ctx && out7, "dim_train_debug_geometry: NULL argument");
  train_debug_geometry(ctx, id, out7);
  return 0;
}
}  // extern "C"
