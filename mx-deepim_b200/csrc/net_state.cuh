// net_state.cuh -- state of the FlowNetS network shared by net.cu (inference) and train.cu (training step).
#pragma once
#include <map>

#include "conv_igemm.cuh"

namespace dim {

// FlowNetS tower: name, Cout, Cin, k, stride, pad (deepIM_flownet.py:63-107)
struct LayerSpec {
  const char *name;
  int Cout, Cin, k, stride, pad;
};
static const LayerSpec kLayers[10] = {
    {"flow_conv1", 64, 8, 7, 2, 3},  {"conv2", 128, 64, 5, 2, 2},   {"conv3", 256, 128, 5, 2, 2},
    {"conv3_1", 256, 256, 3, 1, 1},  {"conv4", 512, 256, 3, 2, 1},  {"conv4_1", 512, 512, 3, 1, 1},
    {"conv5", 512, 512, 3, 2, 1},    {"conv5_1", 512, 512, 3, 1, 1}, {"conv6", 1024, 512, 3, 2, 1},
    {"conv6_1", 1024, 1024, 3, 1, 1}};

struct LayerGeom {
  // logical
  int Cin, Cout, k, stride, pad, Hin, Win, Ho, Wo;
  // input buffer: [B, rows, cols, Cbuf] bf16 (conv1: space-to-depth, Cbuf = 32)
  int rows, cols, Cbuf, py, px;  // py/px: where the producer writes pixel (0,0) (pre-s2d for conv1)
  // implicit GEMM view
  int KH, KW, stride_eff, Ceff, Hq;
  int BLOCK_N, BLOCK_K, BW, BH, n_col_tiles, kblocks;
  int pair;  // 1: CTA-pair kernel (cta_group::2, 256 x BLOCK_N tiles)
  int occ;  // resident CTAs per SM of the persistent kernel variant used for this layer (bf16 mode)
};

struct TensorMaps {
  ConvKParams kp[10];
  int ksplit[10];
  LayerGeom g[10];  // per-batch-size effective geometry (tile width may depend on the batch)
};

// conv1 operand K order inside one (dh, dw) tap of the space-to-depth form: the four 8-channel chunks (ph, pw) sit at
// ph*16 + pw*8 -- the order of the input strip's chunk planes -- except for dw = 3, where the two pw = 0 chunks come first
// (the pw = 1 half is kw = 7, outside the filter): the kernels feed that K step from chunk planes 0 and 2 and drop the other.
__host__ __device__ inline int conv1_kslot(int dw, int ph, int pw) { return dw == 3 ? pw * 16 + ph * 8 : ph * 16 + pw * 8; }

struct NetState {
  LayerGeom g[10];
  __nv_bfloat16 *w_hi[10] = {}, *w_lo[10] = {};
  __nv_bfloat16 *w_f16[10] = {};  // the same packs as IEEE half (DIM_PREC_FP16; 16-bit payload, typed like the others)
  float *bias[10] = {};
  __nv_bfloat16 *act_hi[11] = {}, *act_lo[11] = {};  // act[i] = input of layer i, act[10] = fc6 input
  size_t act_elems_per_image[11] = {};
  // fc
  __nv_bfloat16 *fc6_w_hi = nullptr, *fc6_w_lo = nullptr;  // [256][81920] in (h,w,c) order
  __nv_bfloat16 *fc6_w_f16 = nullptr;
  bool train_aliased = false;  // dim_train_load_params made biases / head parameters alias the fp32 master vector
  bool f16_stale = false;  // training updated the weights: the fp16 packs are re-derived lazily (net_refresh_f16)
  float *fc6_b = nullptr, *fc7_wT = nullptr, *fc7_b = nullptr, *rot_w = nullptr, *rot_b = nullptr,
        *trans_w = nullptr, *trans_b = nullptr;
  float *fc6_partial = nullptr;  // [FC6_SPLITS][max_batch][256]
  float *conv_partial = nullptr;
  size_t conv_partial_elems = 0;
  // kernel variants; defaults = the measured-best set (tools/conv_lab.py, profiles/r02_conv_lab.json), switchable at run
  // time through dim_debug_set_option for A/B measurements
  bool conv1_stack = true;   // conv1 (single-pass precisions) on conv1_stack_kernel; false: conv1_roll_kernel (always used by bf16x3)
  int pair_mask = 1 << 1;    // bit i: conv layer i runs on the CTA-pair (cta_group::2) kernel; default: conv2 (N = 128)
  cudaEvent_t *layer_events = nullptr;  // tuning hook: 11 events around the conv layers of the last forward
  bool loaded = false, net_ok = false;
  float *save_h6 = nullptr, *save_h7 = nullptr;
  cudaEvent_t repack_done = nullptr;  // training: the operand packs are refreshed on an internal stream after an update;
                                      // every consumer (net_forward) orders itself behind this event
  bool lo_stale = false;  // training updated the weights without refreshing the bf16 'lo' halves (bf16x3 mode refreshes lazily)  // training: fc6 / fc7 activations kept for the backward pass ([B][256])
  std::map<int, TensorMaps> maps;  // per batch size (+ kF16MapKey for the fp16 operand maps)
  int max_batch = 0, num_sms = 148;
};

static constexpr int FC6_K = 1024 * 8 * 10;
static constexpr int FC6_KC = 256;
static constexpr int FC6_SPLITS = FC6_K / FC6_KC;  // 320


// helpers implemented in net.cu
int encode_map(CUtensorMap *m, void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
               const uint32_t *box, int block_k /*64: SW128, 32: SW64, 0: no swizzle*/);
uint32_t make_idesc(int M, int N, bool f16 = false);
static constexpr int kF16MapKey = 1 << 20;
int train_refresh_lo(dim_ctx *ctx, cudaStream_t st);  // train.cu
int net_forward(dim_ctx *ctx, int B, int precision, const float *zoom_factor, float *rot_out, float *trans_out,
                float *se3_out, cudaStream_t st, cudaEvent_t after_conv);

template <typename T>
static int dev_alloc(dim_ctx *ctx, T **p, size_t n, bool zero) {
  void *q = nullptr;
  DIM_CHECK(cudaMalloc(&q, n * sizeof(T)));
  if (zero) DIM_CHECK(cudaMemset(q, 0, n * sizeof(T)));
  ctx->owned.push_back(q);
  *p = reinterpret_cast<T *>(q);
  return 0;
}

}  // namespace dim
