/*
 * deepim_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; never linked into or
 * called from the product path).
 *
 * Plain-C restatement of the geometry half of the mx-DeepIM test-time hot path
 * (render -> bbox -> zoom -> se3 compose, plus the reprojection-flow label kernel).
 * Every function cites the reference file:line it follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.
 *
 * Parity status (see DESIGN.md "Oracle pinning"):
 *   - se3 compose / delta     : PINNED against the live reference
 *                               lib/pair_matching/RT_transform.py (tests/golden/se3_*.npz)
 *   - reprojection flow       : PINNED against lib/pair_matching/flow.py:calc_flow and a
 *                               restatement of lib/flow_c/gpu_flow_kernel.cu
 *   - zoom bbox / zoom factor : restated from deepim/operator_py/zoom_mask.py (numpy part is
 *                               pinned by replaying the reference's numpy expressions in
 *                               tests/golden/make_golden.py)
 *   - GridGenerator/BilinearSampler, OpenGL rasterisation (unlit orc_render and Lambert-lit
 *     orc_render_lit, render_py_light_modelnet_multi.py:36-79): third-party (MXNet, glumpy/GL,
 *     not vendored) -> PARITY UNPINNED; the formulas below define the contract.
 *   - training graph / gradients / SGD: oracle/train_oracle.py (torch-CPU fp32; unpinned, see its header)
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: the CUDA kernels are compiled with -fmad=false so that the
 * float32 sequences below are reproduced bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * 1. Rasteriser  (replaces lib/render_glumpy/render_py_multi.py:101-160 + GLSL l.22-52)
 *
 * Conventions restated from the reference:
 *   - camera = OpenCV, view = yz-flip * [R|t]            (render_py_multi.py:153-160)
 *   - projection with u0 = cx + 0.5, v0 = cy + 0.5        (render_py_multi.py:134-151)
 *     => integer pixel (i,j) is sampled at image-plane point (u,v) = (j,i)
 *   - unlit texture lookup, background colour 0, background depth 0 (l.120-129)
 *   - depth returned is metric camera Z (the GL depth linearisation l.126-127 inverts the
 *     projection exactly up to depth-buffer quantisation)
 *   - near/far 0.25 / 6.0 fragments outside are dropped (GL clip volume)
 *   - no face culling (GL_CULL_FACE is never enabled), GL_LESS depth test, first-drawn wins ties
 * Oracle-defined (GL leaves them to the implementation):
 *   - 24.8 fixed-point vertex snapping, int64 edge functions, antisymmetric tie rule
 *   - nearest texel, clamp to edge  (glumpy Texture2D default, SURVEY App.B-24)
 *   - triangles with a vertex at Z <= 1e-6 or |screen coord| > 1e6 px are dropped (no clipping)
 * ---------------------------------------------------------------------------------------- */

typedef struct {
  int32_t X, Y;   /* 24.8 fixed point screen position               */
  float iz;       /* 1 / Zc                                          */
  float uz, vz;   /* u / Zc, v / Zc                                  */
  int32_t ok;
} orc_pvert;

static void orc_project_vertex(const float *pose /*3x4 row major*/, float fx, float fy, float cx,
                               float cy, const float *p, const float *uv, orc_pvert *o) {
  float x = p[0], y = p[1], z = p[2];
  float xc = ((pose[0] * x + pose[1] * y) + pose[2] * z) + pose[3];
  float yc = ((pose[4] * x + pose[5] * y) + pose[6] * z) + pose[7];
  float zc = ((pose[8] * x + pose[9] * y) + pose[10] * z) + pose[11];
  int ok = (zc > 1e-6f);
  float sx = 0.f, sy = 0.f, iz = 0.f;
  if (ok) {
    sx = (fx * xc) / zc + cx;
    sy = (fy * yc) / zc + cy;
    ok = (fabsf(sx) <= 1e6f) && (fabsf(sy) <= 1e6f);
    iz = 1.0f / zc;
  }
  o->ok = ok;
  if (ok) {
    o->X = (int32_t)lrintf(sx * 256.0f);
    o->Y = (int32_t)lrintf(sy * 256.0f);
    o->iz = iz;
    o->uz = uv[0] * iz;
    o->vz = uv[1] * iz;
  } else {
    o->X = o->Y = 0;
    o->iz = o->uz = o->vz = 0.f;
  }
}

static inline int64_t orc_edge(int32_t ax, int32_t ay, int32_t bx, int32_t by, int32_t px,
                               int32_t py) {
  return (int64_t)(bx - ax) * (int64_t)(py - ay) - (int64_t)(by - ay) * (int64_t)(px - ax);
}

/* tie rule for a pixel centre exactly on edge a->b: owned iff dy>0 or (dy==0 and dx<0) */
static inline int orc_edge_owns(int32_t ax, int32_t ay, int32_t bx, int32_t by) {
  int32_t dx = bx - ax, dy = by - ay;
  return (dy > 0) || (dy == 0 && dx < 0);
}

static inline uint32_t orc_fbits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

typedef struct {
  orc_pvert a, b, c;
  int64_t area;
  int valid;
  int swapped; /* b and c exchanged to make the area positive */
} orc_tri;

static void orc_setup_tri(const orc_pvert *pv, const int32_t *face, orc_tri *t) {
  t->a = pv[face[0]];
  t->b = pv[face[1]];
  t->c = pv[face[2]];
  t->valid = t->a.ok && t->b.ok && t->c.ok;
  t->swapped = 0;
  if (!t->valid) return;
  int64_t area = orc_edge(t->a.X, t->a.Y, t->b.X, t->b.Y, t->c.X, t->c.Y);
  if (area == 0) {
    t->valid = 0;
    return;
  }
  if (area < 0) {
    orc_pvert tmp = t->b;
    t->b = t->c;
    t->c = tmp;
    area = -area;
    t->swapped = 1;
  }
  t->area = area;
}

/* coverage + barycentrics + depth at pixel (i,j); returns 1 when the fragment survives */
static inline int orc_fragment(const orc_tri *t, int i, int j, float zn, float zf, float *b0,
                               float *b1, float *b2, float *iz_out, float *z_out) {
  int32_t px = j << 8, py = i << 8;
  int64_t w0 = orc_edge(t->b.X, t->b.Y, t->c.X, t->c.Y, px, py);
  int64_t w1 = orc_edge(t->c.X, t->c.Y, t->a.X, t->a.Y, px, py);
  int64_t w2 = orc_edge(t->a.X, t->a.Y, t->b.X, t->b.Y, px, py);
  if (w0 < 0 || w1 < 0 || w2 < 0) return 0;
  if (w0 == 0 && !orc_edge_owns(t->b.X, t->b.Y, t->c.X, t->c.Y)) return 0;
  if (w1 == 0 && !orc_edge_owns(t->c.X, t->c.Y, t->a.X, t->a.Y)) return 0;
  if (w2 == 0 && !orc_edge_owns(t->a.X, t->a.Y, t->b.X, t->b.Y)) return 0;
  float fa = (float)t->area;
  float c0 = (float)w0 / fa, c1 = (float)w1 / fa, c2 = (float)w2 / fa;
  float iz = (c0 * t->a.iz + c1 * t->b.iz) + c2 * t->c.iz;
  float z = 1.0f / iz;
  if (!(z >= zn && z <= zf)) return 0;
  *b0 = c0;
  *b1 = c1;
  *b2 = c2;
  *iz_out = iz;
  *z_out = z;
  return 1;
}

/* colour chain of the reference: u8 texel -> GL float (c/255) -> "bgr_gl *= 255"
 * (render_py_multi.py:124) -> optional uint8 truncation (deepim/core/tester.py:188). */
static inline float orc_colour(uint8_t c, int trunc_u8) {
  float f = ((float)c / 255.0f) * 255.0f;
  if (trunc_u8) f = (float)(uint8_t)f;
  return f;
}

/*
 * orc_render: one instance.
 *  out_bgr   : [H,W,3] float32, BGR in [0,255]          (Render_Py.render return value 0)
 *  out_depth : [H,W]   float32, metres, background 0    (Render_Py.render return value 1)
 *  out_image : [3,H,W] float32, RGB - pixel_means_rgb (float64 subtract, float32 store;
 *              lib/utils/image.py:583-594 transform + nd.array)
 *  out_mask  : [H,W]   float32, depth > 0.2             (deepim/core/tester.py:440)
 *  bbox_ren  : 4 ints  x0,x1,y0,y1 of out_mask (min/max nonzero col/row), or -1 if empty
 *  any output pointer may be NULL.
 */
/* Lambert shading of the lit renderer (lib/render_glumpy/render_py_light_modelnet_multi.py:36-79 fragment shader):
 * position / normal interpolated perspective-correctly in model space, moved to the GL camera frame (x, -y, -z of
 * the OpenCV frame, _get_view_mtx), brightness = clamp(cos(normal, light - position), 0, 1),
 * colour = texel * ((1 - ratio) + ratio * brightness) * light_intensity, quantised like the 8-bit framebuffer the
 * reference reads back (np.round(rgb * 255), l.160).  float32, one operation per line (no contraction). */
static void orc_shade_lit(float b0, float b1, float b2, float iz, float izA, float izB, float izC, const float *vA,
                          const float *vB, const float *vC, const float *nA, const float *nB, const float *nC,
                          const float *pose, const float *light_pos, const float *light_int, float a0, float a1,
                          const uint8_t *tp, float *rgb) {
  float w0 = b0 * izA, w1 = b1 * izB, w2 = b2 * izC;
  float pm[3], nm[3], pc[3], nc[3];
  for (int k = 0; k < 3; ++k) {
    pm[k] = ((w0 * vA[k] + w1 * vB[k]) + w2 * vC[k]) / iz;
    nm[k] = ((w0 * nA[k] + w1 * nB[k]) + w2 * nC[k]) / iz;
  }
  for (int r = 0; r < 3; ++r) {
    pc[r] = ((pose[4 * r] * pm[0] + pose[4 * r + 1] * pm[1]) + pose[4 * r + 2] * pm[2]) + pose[4 * r + 3];
    nc[r] = (pose[4 * r] * nm[0] + pose[4 * r + 1] * nm[1]) + pose[4 * r + 2] * nm[2];
  }
  float s0 = light_pos[0] - pc[0], s1 = light_pos[1] - (0.f - pc[1]), s2 = light_pos[2] - (0.f - pc[2]);
  float g0 = nc[0], g1 = 0.f - nc[1], g2 = 0.f - nc[2];
  float dot = (g0 * s0 + g1 * s1) + g2 * s2;
  float ls = sqrtf((s0 * s0 + s1 * s1) + s2 * s2), ln = sqrtf((g0 * g0 + g1 * g1) + g2 * g2);
  float den = ls * ln, br = 0.f;
  if (den > 0.f) br = dot / den;
  br = br < 1.f ? br : 1.f;
  br = br > 0.f ? br : 0.f;
  float scale = a0 + a1 * br;
  for (int c = 0; c < 3; ++c) {
    float col = ((float)tp[c] / 255.0f) * (scale * light_int[c]);
    col = col < 1.f ? col : 1.f;
    col = col > 0.f ? col : 0.f;
    rgb[c] = rintf(col * 255.0f);
  }
}

static void orc_render_impl(const float *verts, const float *uvs, int32_t V, const int32_t *faces,
                        int32_t F, const uint8_t *tex, int32_t Th, int32_t Tw, const float *pose,
                        const float *K4 /*fx,fy,cx,cy*/, float zn, float zf, int32_t H, int32_t W,
                        const double *means_rgb, int32_t trunc_u8, float *out_bgr,
                        float *out_depth, float *out_image, float *out_mask, int32_t *bbox_ren,
                        const float *normals, const float *light_pos, const float *light_int, float a0, float a1) {
  orc_pvert *pv = (orc_pvert *)malloc(sizeof(orc_pvert) * (size_t)V);
  uint64_t *zb = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)H * W);
  for (size_t k = 0; k < (size_t)H * W; ++k) zb[k] = ~(uint64_t)0;
  for (int32_t v = 0; v < V; ++v)
    orc_project_vertex(pose, K4[0], K4[1], K4[2], K4[3], verts + 3 * v, uvs + 2 * v, pv + v);

  for (int32_t f = 0; f < F; ++f) {
    orc_tri t;
    orc_setup_tri(pv, faces + 3 * f, &t);
    if (!t.valid) continue;
    int32_t minX = t.a.X < t.b.X ? t.a.X : t.b.X;
    if (t.c.X < minX) minX = t.c.X;
    int32_t maxX = t.a.X > t.b.X ? t.a.X : t.b.X;
    if (t.c.X > maxX) maxX = t.c.X;
    int32_t minY = t.a.Y < t.b.Y ? t.a.Y : t.b.Y;
    if (t.c.Y < minY) minY = t.c.Y;
    int32_t maxY = t.a.Y > t.b.Y ? t.a.Y : t.b.Y;
    if (t.c.Y > maxY) maxY = t.c.Y;
    int32_t j0 = (minX + 255) >> 8, j1 = maxX >> 8, i0 = (minY + 255) >> 8, i1 = maxY >> 8;
    if (j0 < 0) j0 = 0;
    if (i0 < 0) i0 = 0;
    if (j1 > W - 1) j1 = W - 1;
    if (i1 > H - 1) i1 = H - 1;
    for (int32_t i = i0; i <= i1; ++i)
      for (int32_t j = j0; j <= j1; ++j) {
        float b0, b1, b2, iz, z;
        if (!orc_fragment(&t, i, j, zn, zf, &b0, &b1, &b2, &iz, &z)) continue;
        uint64_t key = ((uint64_t)orc_fbits(z) << 32) | (uint32_t)f;
        if (key < zb[(size_t)i * W + j]) zb[(size_t)i * W + j] = key;
      }
  }

  int32_t bx0 = W, bx1 = -1, by0 = H, by1 = -1;
  for (int32_t i = 0; i < H; ++i)
    for (int32_t j = 0; j < W; ++j) {
      size_t p = (size_t)i * W + j;
      uint64_t key = zb[p];
      float rgb[3] = {0.f, 0.f, 0.f}, depth = 0.f;
      if (key != ~(uint64_t)0) {
        int32_t f = (int32_t)(uint32_t)(key & 0xffffffffu);
        orc_tri t;
        orc_setup_tri(pv, faces + 3 * f, &t);
        float b0 = 0.f, b1 = 0.f, b2 = 0.f, iz = 1.f, z = 0.f;
        orc_fragment(&t, i, j, zn, zf, &b0, &b1, &b2, &iz, &z);
        float un = (b0 * t.a.uz + b1 * t.b.uz) + b2 * t.c.uz;
        float vn = (b0 * t.a.vz + b1 * t.b.vz) + b2 * t.c.vz;
        float u = un / iz, v = vn / iz;
        int32_t tx = (int32_t)floorf(u * (float)Tw), ty = (int32_t)floorf(v * (float)Th);
        if (tx < 0) tx = 0;
        if (tx > Tw - 1) tx = Tw - 1;
        if (ty < 0) ty = 0;
        if (ty > Th - 1) ty = Th - 1;
        const uint8_t *tp = tex + ((size_t)ty * Tw + tx) * 3;
        if (normals) {
          const int32_t *fi = faces + 3 * f;
          const int32_t iA = fi[0], iB = t.swapped ? fi[2] : fi[1], iC = t.swapped ? fi[1] : fi[2];
          orc_shade_lit(b0, b1, b2, iz, t.a.iz, t.b.iz, t.c.iz, verts + 3 * iA, verts + 3 * iB, verts + 3 * iC,
                        normals + 3 * iA, normals + 3 * iB, normals + 3 * iC, pose, light_pos, light_int, a0, a1, tp, rgb);
        } else {
          rgb[0] = orc_colour(tp[0], trunc_u8);
          rgb[1] = orc_colour(tp[1], trunc_u8);
          rgb[2] = orc_colour(tp[2], trunc_u8);
        }
        depth = z;
      }
      float m = depth > 0.2f ? 1.f : 0.f;
      if (m > 0.f) {
        if (j < bx0) bx0 = j;
        if (j > bx1) bx1 = j;
        if (i < by0) by0 = i;
        if (i > by1) by1 = i;
      }
      if (out_bgr) {
        out_bgr[p * 3 + 0] = rgb[2];
        out_bgr[p * 3 + 1] = rgb[1];
        out_bgr[p * 3 + 2] = rgb[0];
      }
      if (out_depth) out_depth[p] = depth;
      if (out_image) {
        size_t P = (size_t)H * W;
        if (trunc_u8) {
          /* test path: image.transform works in float64 (np.zeros tensor), nd.array casts to float32 */
          out_image[p] = (float)((double)rgb[0] - means_rgb[0]);
          out_image[P + p] = (float)((double)rgb[1] - means_rgb[1]);
          out_image[2 * P + p] = (float)((double)rgb[2] - means_rgb[2]);
        } else {
          /* train path (batch_updater_py_multi.py:234-235): float32 image -= float32 pixel_means */
          out_image[p] = rgb[0] - (float)means_rgb[0];
          out_image[P + p] = rgb[1] - (float)means_rgb[1];
          out_image[2 * P + p] = rgb[2] - (float)means_rgb[2];
        }
      }
      if (out_mask) out_mask[p] = m;
    }
  if (bbox_ren) {
    if (bx1 < 0) {
      bbox_ren[0] = bbox_ren[1] = bbox_ren[2] = bbox_ren[3] = -1;
    } else {
      bbox_ren[0] = bx0;
      bbox_ren[1] = bx1;
      bbox_ren[2] = by0;
      bbox_ren[3] = by1;
    }
  }
  free(pv);
  free(zb);
}

/* ------------------------------------------------------------------------------------------
 * 2. Mask bbox + zoom factor   (deepim/operator_py/zoom_mask.py:29-103)
 * ---------------------------------------------------------------------------------------- */

/* min/max nonzero column/row of (mask > thresh); x0,x1,y0,y1 or -1 (zoom_mask.py:51-58,62-66) */

ORC_API void orc_render(const float *verts, const float *uvs, int32_t V, const int32_t *faces,
                        int32_t F, const uint8_t *tex, int32_t Th, int32_t Tw, const float *pose,
                        const float *K4 /*fx,fy,cx,cy*/, float zn, float zf, int32_t H, int32_t W,
                        const double *means_rgb, int32_t trunc_u8, float *out_bgr,
                        float *out_depth, float *out_image, float *out_mask, int32_t *bbox_ren) {
  orc_render_impl(verts, uvs, V, faces, F, tex, Th, Tw, pose, K4, zn, zf, H, W, means_rgb, trunc_u8, out_bgr, out_depth,
                  out_image, out_mask, bbox_ren, NULL, NULL, NULL, 0.f, 0.f);
}

/* Render_Py_Light_ModelNet_Multi.render (render_py_light_modelnet_multi.py:131-175): per-vertex normals,
 * light_pos / light_int (3 floats each, GL camera frame), a0 = 1 - brightness_ratio, a1 = brightness_ratio. */
ORC_API void orc_render_lit(const float *verts, const float *uvs, const float *normals, int32_t V, const int32_t *faces,
                            int32_t F, const uint8_t *tex, int32_t Th, int32_t Tw, const float *pose, const float *K4,
                            float zn, float zf, int32_t H, int32_t W, const double *means_rgb, const float *light_pos,
                            const float *light_int, float a0, float a1, float *out_bgr, float *out_depth,
                            float *out_image, float *out_mask, int32_t *bbox_ren) {
  orc_render_impl(verts, uvs, V, faces, F, tex, Th, Tw, pose, K4, zn, zf, H, W, means_rgb, 1, out_bgr, out_depth, out_image,
                  out_mask, bbox_ren, normals, light_pos, light_int, a0, a1);
}

ORC_API void orc_mask_bbox(const float *mask, int32_t H, int32_t W, float thresh, int32_t *bbox) {
  int32_t bx0 = W, bx1 = -1, by0 = H, by1 = -1;
  for (int32_t i = 0; i < H; ++i)
    for (int32_t j = 0; j < W; ++j)
      if (mask[(size_t)i * W + j] > thresh) {
        if (j < bx0) bx0 = j;
        if (j > bx1) bx1 = j;
        if (i < by0) by0 = i;
        if (i > by1) by1 = i;
      }
  if (bx1 < 0) {
    bbox[0] = bbox[1] = bbox[2] = bbox[3] = -1;
  } else {
    bbox[0] = bx0;
    bbox[1] = bx1;
    bbox[2] = by0;
    bbox[3] = by1;
  }
}

static inline double orc_dmax(double a, double b) { return a > b ? a : b; }

/*
 * zoom factor from the two bboxes and src_pose (zoom_mask.py:59-103).  Mixed precision exactly as
 * the reference's numpy 1.x (legacy scalar promotion, the reference predates NEP 50): K and
 * src_pose are float32 arrays so c = K.t and c_x = c0/c2 are float32; every later expression
 * mixes that float32 scalar with int64 scalars or python numbers and is therefore evaluated in
 * float64 (distances, crop_height, wx, tx, ty); the four results are stored as float32
 * (zoom_mask.py:96-103).
 * Returns 0 ok, 1 = observed bbox empty (the reference raises on np.min of an empty array).
 */
ORC_API int32_t orc_zoom_factor(const int32_t *bbox_real, const int32_t *bbox_ren,
                                const float *src_pose, const float *K9, int32_t H, int32_t W,
                                float *zoom_factor) {
  if (bbox_real[1] < 0) {
    zoom_factor[0] = zoom_factor[1] = 1.f;
    zoom_factor[2] = zoom_factor[3] = 0.f;
    return 1;
  }
  double real_x0 = bbox_real[0], real_x1 = bbox_real[1], real_y0 = bbox_real[2],
         real_y1 = bbox_real[3];
  float t0 = src_pose[3], t1 = src_pose[7], t2 = src_pose[11];
  float c0 = (K9[0] * t0 + K9[1] * t1) + K9[2] * t2;
  float c1 = (K9[3] * t0 + K9[4] * t1) + K9[5] * t2;
  float c2 = (K9[6] * t0 + K9[7] * t1) + K9[8] * t2;
  float cxf = c0 / c2, cyf = c1 / c2;
  double ren_x0, ren_x1, ren_y0, ren_y1, zcx, zcy;
  float tx, ty;
  if (bbox_ren[1] < 0) { /* "NO POINT VALID IN MASK rendered" branch, zoom_mask.py:70-77 */
    ren_x0 = real_x0;
    ren_x1 = real_x1;
    ren_y0 = real_y0;
    ren_y1 = real_y1;
    zcx = (real_x0 + real_x1) * 0.5;
    zcy = (real_y0 + real_y1) * 0.5;
  } else {
    ren_x0 = bbox_ren[0];
    ren_x1 = bbox_ren[1];
    ren_y0 = bbox_ren[2];
    ren_y1 = bbox_ren[3];
    zcx = (double)cxf;
    zcy = (double)cyf;
  }
  tx = (float)(zcx / (double)W * 2.0 - 1.0);
  ty = (float)(zcy / (double)H * 2.0 - 1.0);
  double left = orc_dmax(zcx - ren_x0, zcx - real_x0);
  double right = orc_dmax(ren_x1 - zcx, real_x1 - zcx);
  double up = orc_dmax(zcy - ren_y0, zcy - real_y0);
  double down = orc_dmax(real_y1 - zcy, ren_y1 - zcy);
  double m = orc_dmax(orc_dmax(0.75 * right, 0.75 * left), orc_dmax(up, down));
  double crop_height = m * 1.4 * 2;
  double wx = crop_height / (double)H;
  zoom_factor[0] = (float)wx;
  zoom_factor[1] = (float)wx;
  zoom_factor[2] = tx;
  zoom_factor[3] = ty;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * 3. Affine grid + bilinear sampler.  Third-party MXNet ops (GridGenerator 'affine',
 *    BilinearSampler) -- PARITY UNPINNED, restated:
 *      x_t = -1 + j*(2/(W-1)),  x_s = wx*x_t + tx,  x = (x_s+1)*(W-1)/2, zero padding.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t x0, y0;
  float wx1, wy1; /* weight of the top-left tap along x / y */
} orc_tap;

static inline orc_tap orc_src_coord(int i, int j, float wx, float wy, float tx, float ty, int H,
                                    int W) {
  float stepx = (float)(2.0 / (double)(W - 1)), stepy = (float)(2.0 / (double)(H - 1));
  float xt = -1.0f + (float)j * stepx;
  float yt = -1.0f + (float)i * stepy;
  float xs = wx * xt + tx;
  float ys = wy * yt + ty;
  float xr = ((xs + 1.0f) * (float)(W - 1)) / 2.0f;
  float yr = ((ys + 1.0f) * (float)(H - 1)) / 2.0f;
  float fx0 = floorf(xr), fy0 = floorf(yr);
  orc_tap t;
  /* saturating casts keep far-out-of-frame samples out of bounds without UB */
  t.x0 = fx0 < -4.0f ? -4 : (fx0 > (float)(W + 4) ? W + 4 : (int32_t)fx0);
  t.y0 = fy0 < -4.0f ? -4 : (fy0 > (float)(H + 4) ? H + 4 : (int32_t)fy0);
  t.wx1 = 1.0f - (xr - fx0);
  t.wy1 = 1.0f - (yr - fy0);
  return t;
}

static inline float orc_fetch(const float *img, int H, int W, int y, int x, float add) {
  if (x < 0 || x > W - 1 || y < 0 || y > H - 1) return 0.0f;
  return img[(size_t)y * W + x] + add;
}

static inline float orc_bilinear(const float *img, int H, int W, orc_tap t, float add) {
  float tl = orc_fetch(img, H, W, t.y0, t.x0, add);
  float tr = orc_fetch(img, H, W, t.y0, t.x0 + 1, add);
  float bl = orc_fetch(img, H, W, t.y0 + 1, t.x0, add);
  float br = orc_fetch(img, H, W, t.y0 + 1, t.x0 + 1, add);
  /* the four tap weights are formed once per output pixel and accumulated with fused
   * multiply-adds (same bilinear formula as MXNet's BilinearSampler; expression order is ours, the
   * third-party kernel's own rounding order is unknowable -> parity unpinned, <= 1 ulp apart) */
  float wx1 = t.wx1, wy1 = t.wy1, ax = 1.0f - wx1, ay = 1.0f - wy1;
  float a = wy1 * wx1, b = wy1 * ax, c = ay * wx1, d = ay * ax;
  return fmaf(br, d, fmaf(bl, c, fmaf(tr, b, tl * a)));
}

/*
 * mode 0: plain sample                          (ZoomDepth, zoom_depth.py:34-42)
 * mode 1: round(sample)  roundf half-away       (ZoomMask zoom_mask.py:105-107)
 * mode 2: binarise input (>0.2) then round      (ZoomMaskWithFactor zoom_mask_with_factor.py:36-62)
 * mode 3: (img+mean) sample - mean              (ZoomImageWithFactor zoom_image_with_factor.py:44-62)
 * mode 4: sample * scale                        (ZoomFlow flow, zoom_flow.py:55-64)
 * mode 5: round(sample - 0.45)                  (ZoomFlow weights, zoom_flow.py:66-71)
 * affine = (wx, wy, tx, ty) actually used by the GridGenerator.
 */
ORC_API void orc_zoom_plane(const float *src, float *dst, int32_t H, int32_t W, const float *affine,
                            int32_t mode, float param) {
  float wx = affine[0], wy = affine[1], tx = affine[2], ty = affine[3];
  float *tmp = NULL;
  if (mode == 2) {
    tmp = (float *)malloc(sizeof(float) * (size_t)H * W);
    for (size_t k = 0; k < (size_t)H * W; ++k) tmp[k] = src[k] > 0.2f ? 1.0f : 0.0f;
    src = tmp;
  }
  for (int32_t i = 0; i < H; ++i)
    for (int32_t j = 0; j < W; ++j) {
      orc_tap t = orc_src_coord(i, j, wx, wy, tx, ty, H, W);
      float v;
      switch (mode) {
        case 1:
        case 2:
          v = roundf(orc_bilinear(src, H, W, t, 0.f));
          break;
        case 3:
          v = orc_bilinear(src, H, W, t, param) - param;
          break;
        case 4:
          v = orc_bilinear(src, H, W, t, 0.f) * param;
          break;
        case 5:
          v = roundf(orc_bilinear(src, H, W, t, 0.f) - 0.45f);
          break;
        default:
          v = orc_bilinear(src, H, W, t, 0.f);
      }
      dst[(size_t)i * W + j] = v;
    }
  if (tmp) free(tmp);
}

/* inverse-zoom affine (zoom_flow.py:35-44, zoom_mask_with_factor.py:43-52); float32 zoom_factor
 * scalars combined with python numbers -> float64 under numpy 1.x, stored float32 */
ORC_API void orc_inv_zoom_affine(const float *zf, int32_t H, int32_t W, float *affine) {
  double wx_in = zf[0], wy_in = zf[1], tx_in = zf[2], ty_in = zf[3];
  double wx = 1.0 / wx_in, wy = 1.0 / wy_in;
  double crop_w = wx_in * (double)W, crop_h = wy_in * (double)H;
  double cx = tx_in * 0.5 * (double)W + 0.5 * (double)W;
  double cy = ty_in * 0.5 * (double)H + 0.5 * (double)H;
  double tx = ((double)W * 0.5 - cx) / crop_w * 2.0;
  double ty = ((double)H * 0.5 - cy) / crop_h * 2.0;
  affine[0] = (float)wx;
  affine[1] = (float)wy;
  affine[2] = (float)tx;
  affine[3] = (float)ty;
}

/* box mask from a bbox, END-EXCLUSIVE (lib/pair_matching/data_pair.py:93-105) */
ORC_API void orc_box_mask(const int32_t *bbox, int32_t H, int32_t W, float *mask) {
  memset(mask, 0, sizeof(float) * (size_t)H * W);
  if (bbox[1] < 0) return;
  for (int32_t i = bbox[2]; i < bbox[3]; ++i)
    for (int32_t j = bbox[0]; j < bbox[1]; ++j) mask[(size_t)i * W + j] = 1.0f;
}

/* ZoomTrans forward (zoom_trans.py:22-46): wx used for both axes; numpy float32 in, float64
 * zeros array out, stored float32 */
ORC_API void orc_zoom_trans(const float *zoom_factor, const float *trans, int32_t B, int32_t inv,
                            float *out) {
  for (int32_t b = 0; b < B; ++b) {
    float w = zoom_factor[4 * b];
    float dx = trans[3 * b], dy = trans[3 * b + 1], dz = trans[3 * b + 2];
    out[3 * b + 0] = inv ? dx * w : dx / w;
    out[3 * b + 1] = inv ? dy * w : dy / w;
    out[3 * b + 2] = dz;
  }
}

/* ------------------------------------------------------------------------------------------
 * 4. SE(3) compose / delta in float64  (lib/pair_matching/RT_transform.py:127-151, 383-429,
 *    47-61, 74-95, 16-44).  rot_coord: 0 MODEL, 1 CAMERA, 2 CAMERA_NEW.
 * ---------------------------------------------------------------------------------------- */
static void orc_quat2mat(const double *q, double *M) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double Nq = w * w + x * x + y * y + z * z;
  if (Nq < 2.220446049250313e-16 * 4.0) { /* _FLOAT_EPS = finfo(float).eps * 4.0 (l.236-238) */
    M[0] = M[4] = M[8] = 1.0;
    M[1] = M[2] = M[3] = M[5] = M[6] = M[7] = 0.0;
    return;
  }
  double s = 2.0 / Nq;
  double X = x * s, Y = y * s, Z = z * s;
  double wX = w * X, wY = w * Y, wZ = w * Z, xX = x * X, xY = x * Y, xZ = x * Z, yY = y * Y,
         yZ = y * Z, zZ = z * Z;
  M[0] = 1.0 - (yY + zZ);
  M[1] = xY - wZ;
  M[2] = xZ + wY;
  M[3] = xY + wZ;
  M[4] = 1.0 - (xX + zZ);
  M[5] = yZ - wX;
  M[6] = xZ - wY;
  M[7] = yZ + wX;
  M[8] = 1.0 - (xX + yY);
}

ORC_API void orc_rt_transform(const double *pose_src /*3x4*/, const double *quat,
                              const double *t_delta, const double *T_means, const double *T_stds,
                              int32_t rot_coord, double *pose_out) {
  double n = sqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
  double q[4] = {quat[0] / n, quat[1] / n, quat[2] / n, quat[3] / n};
  double Rd[9];
  orc_quat2mat(q, Rd);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k)
        acc += (rot_coord == 0) ? pose_src[i * 4 + k] * Rd[k * 3 + j]
                                : Rd[i * 3 + k] * pose_src[k * 4 + j];
      pose_out[i * 4 + j] = acc;
    }
  double d0 = t_delta[0] * T_stds[0] + T_means[0];
  double d1 = t_delta[1] * T_stds[1] + T_means[1];
  double d2 = t_delta[2] * T_stds[2] + T_means[2];
  double sx = pose_src[3], sy = pose_src[7], sz = pose_src[11];
  double z2 = sz / exp(d2);
  pose_out[11] = z2;
  if (rot_coord == 2) {
    pose_out[3] = sz * d0 + sx;
    pose_out[7] = sz * d1 + sy;
  } else {
    pose_out[3] = z2 * (d0 + sx / sz);
    pose_out[7] = z2 * (d1 + sy / sz);
  }
}

/* calc_RT_delta with rot_type MATRIX (RT_transform.py:16-44): returns R_delta (3x3), T_delta(3) */
ORC_API void orc_rt_delta(const double *pose_src, const double *pose_tgt, const double *T_means,
                          const double *T_stds, int32_t rot_coord, double *R_delta,
                          double *T_delta) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k)
        acc += (rot_coord == 0) ? pose_src[k * 4 + i] * pose_tgt[k * 4 + j]  /* Rs^T Rt */
                                : pose_tgt[i * 4 + k] * pose_src[j * 4 + k]; /* Rt Rs^T */
      R_delta[i * 3 + j] = acc;
    }
  double sx = pose_src[3], sy = pose_src[7], sz = pose_src[11];
  double tx = pose_tgt[3], ty = pose_tgt[7], tz = pose_tgt[11];
  double d[3];
  if (rot_coord == 2) {
    d[0] = (tx - sx) / sz;
    d[1] = (ty - sy) / sz;
  } else {
    d[0] = tx / tz - sx / sz;
    d[1] = ty / tz - sy / sz;
  }
  d[2] = log(sz / tz);
  for (int k = 0; k < 3; ++k) T_delta[k] = (d[k] - T_means[k]) / T_stds[k];
}

/* ------------------------------------------------------------------------------------------
 * 5. Reprojection-flow label kernel, float32  (lib/flow_c/gpu_flow_kernel.cu:32-69)
 *    flow[b,0]=dh, flow[b,1]=dw, valid[b]; KT = K.T_src->tgt (B,3,4), Kinv (3,3).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_flow(const float *depth_src, const float *depth_tgt, const float *KT,
                      const float *Kinv, int32_t B, int32_t H, int32_t W, float *flow,
                      float *valid) {
  for (int32_t b = 0; b < B; ++b) {
    const float *kt = KT + 12 * b;
    for (int32_t h = 0; h < H; ++h)
      for (int32_t w = 0; w < W; ++w) {
        size_t idx = ((size_t)b * H + h) * W + w;
        float d = depth_src[idx];
        float x = (((float)w * Kinv[0] + (float)h * Kinv[1]) + Kinv[2]) * d;
        float y = (((float)w * Kinv[3] + (float)h * Kinv[4]) + Kinv[5]) * d;
        float z = d;
        float fh = 0.f, fw = 0.f, ok = 0.f;
        /* double literals in the reference (gpu_flow_kernel.cu:45,49,56): the float operands are promoted */
        if ((double)d > 1e-3) {
          float xp = ((x * kt[0] + y * kt[1]) + z * kt[2]) + kt[3];
          float yp = ((x * kt[4] + y * kt[5]) + z * kt[6]) + kt[7];
          float zp = (float)((double)(((x * kt[8] + y * kt[9]) + z * kt[10]) + kt[11]) + 1e-15);
          float wp = xp / zp, hp = yp / zp;
          if (wp >= 0.f && wp <= (float)(W - 1) && hp >= 0.f && hp <= (float)(H - 1)) {
            int32_t wi = (int32_t)roundf(wp), hi = (int32_t)roundf(hp);
            float dt = depth_tgt[((size_t)b * H + hi) * W + wi];
            if ((double)fabsf(zp - dt) < 3e-3) {
              fh = hp - (float)h;
              fw = wp - (float)w;
              ok = 1.f;
            }
          }
        }
        flow[(((size_t)b * 2 + 0) * H + h) * W + w] = fh;
        flow[(((size_t)b * 2 + 1) * H + h) * W + w] = fw;
        valid[idx] = ok;
      }
  }
}

ORC_API int32_t orc_abi_version(void) { return 1; }
