// common.cuh -- context, error handling and launch bookkeeping shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only; no-ops unless a profiler injects itself

#include "../../include/deepim_b200.h"

// NVTX range for the host-side enqueue of one stage (Nsight Systems timelines: SURVEY 5 "tracing"): scoped push / pop
struct DimNvtxRange {
  bool open = true;
  explicit DimNvtxRange(const char *name) { nvtxRangePushA(name); }
  void end() {  // close before the scope does (ranges must nest: end inner ranges first)
    if (open) { nvtxRangePop(); open = false; }
  }
  ~DimNvtxRange() { end(); }
  DimNvtxRange(const DimNvtxRange &) = delete;
  DimNvtxRange &operator=(const DimNvtxRange &) = delete;
};

namespace dim {

void set_error(const char *fmt, ...);
extern long long g_launches;

#define DIM_CHECK(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      dim::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define DIM_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    ++dim::g_launches;                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      dim::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define DIM_REQUIRE(cond, msg)                                        \
  do {                                                                \
    if (!(cond)) {                                                    \
      dim::set_error("%s:%d: %s", __FILE__, __LINE__, msg);           \
      return 2;                                                       \
    }                                                                 \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------- rasteriser
struct PVert {  // projected vertex, 24 B
  int X, Y;     // 24.8 fixed-point screen position
  float iz;     // 1/Zc
  float uz, vz; // u/Zc, v/Zc
  int ok;
};

struct MeshDev {
  const float *verts;   // [V,3]
  const float *uvs;     // [V,2]
  const int *faces;     // [F,3]
  const uint8_t *tex;   // [Th,Tw,3]
  int V, F, Th, Tw;
  const float *normals; // [V,3] per-vertex normals (lit renderer only; nullptr otherwise)
};

// ------------------------------------------------------------------------------------ network
struct ConvLayer {
  int Cin, Cout, k, stride, pad;
  int Hin, Win, Hout, Wout;
  // padded NHWC input buffer geometry (see conv.cu)
  int Hp, Wp, py, px;
};

struct NetState;  // conv.cu

}  // namespace dim

struct dim_ctx {
  int device = 0, max_batch = 0, H = 0, W = 0, max_classes = 0, max_verts = 0, max_faces = 0;
  int num_sms = 148;
  // meshes
  std::vector<dim::MeshDev> meshes_host;
  dim::MeshDev *meshes = nullptr;  // device table [max_classes]
  std::vector<void *> owned;       // device allocations to free
  // raster scratch
  dim::PVert *pverts = nullptr;          // [max_batch, max_verts]
  unsigned long long *vis = nullptr;     // [max_batch, H*W]
  int *vbox = nullptr;                   // [max_batch,4] screen bbox of projected vertices
  float *colour_lut = nullptr;           // [2][3][256]: trunc / no-trunc, RGB - mean
  double lut_means[3] = {-1, -1, -1};
  // zoom scratch
  int *bbox8 = nullptr;      // [max_batch, 8]
  int *status = nullptr;     // [max_batch]
  int *cls_flag = nullptr;   // [max_batch] rasteriser: 2 = class index out of range / mesh missing (raster.cu mesh_for)
  int *status_hist = nullptr;  // [8, max_batch] per-iteration status of the last fused refinement (dim_refine_status)
  float *zoom_factor = nullptr;  // [max_batch,4]
  // refine-loop state
  float *image_rendered = nullptr, *depth_rendered = nullptr, *mask_rendered = nullptr;
  int *bbox_ren = nullptr;   // [max_batch,4]
  double *pose_cur = nullptr;  // [max_batch,3,4]
  float *pose_cur_f32 = nullptr;
  float *se3_cur = nullptr;  // [max_batch,7]
  float4 *ren4 = nullptr, *obs4 = nullptr;  // [max_batch,H,W] pixel-interleaved images of the fused loop
  uint8_t *image_observed_u8 = nullptr;
  int *cls_dev = nullptr;
  double *poses_dev = nullptr;  // [8, max_batch, 12]
  float *se3_hist_dev = nullptr;
  dim::NetState *net = nullptr;
  // CUDA graphs of the fused refinement chain (capi.cu refine_graphed): one executable graph per distinct argument set
  struct RefineGraph {
    std::vector<unsigned char> key;  // every launch argument of refine_core, byte for byte
    cudaGraphExec_t exec = nullptr;  // nullptr: seen once (eager warm-up run), captured on the next call
    long long kernels = 0;           // kernel nodes (dim_launch_count bookkeeping)
  };
  std::vector<RefineGraph> graphs;
  bool use_graph = true;
  // loss weights / normalisers / pose parameterisation (dim_train_set_config); defaults = the shipped LM6d config
  dim_train_config cfg = {0.25f, 0.03f, 0.1f, 3000.f, 0.1f, 20.f, {0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}, 1};
  // stage profiling (dim_profile_enable)
  bool prof = false;
  std::vector<cudaEvent_t> prof_events;  // 5 per recorded iteration
  size_t prof_used = 0;
};
