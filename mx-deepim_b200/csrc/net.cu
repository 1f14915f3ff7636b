// net.cu -- FlowNetS encoder + fc + heads on the device (deepim/symbols/deepIM_flownet.py:53-116 and
// 716-726), weight repacking, tensor-map construction and layer scheduling.
//
//   conv tower : conv1_stack_kernel / conv1_roll_kernel + 9 x conv_igemm_persistent_kernel / conv_igemm_pair_kernel (tcgen05 + TMA,
//                conv_igemm.cuh), split-K + finalize for the layers whose tile count cannot fill 148 SMs
//   fc6        : 81920 -> 256, a pure weight stream (HBM-bound): split-K mma.sync kernel (batch = M = 16),
//                deterministic two-pass reduction (partials reduced in fixed order by the head kernel)
//   head       : fc6 reduce + bias + LeakyReLU -> fc7 -> LeakyReLU -> rot(4), trans(3) ->
//                ZoomTrans^-1 (zoom_trans.py:30-31) -> se3 (B,7)
#include <cuda_fp16.h>
#include <string.h>

#include <mutex>

#include "net_state.cuh"

namespace dim {

// --------------------------------------------------------------------------------- geometry
static void build_geometry(NetState *ns, int H, int W) {
  int h = H, w = W;
  for (int i = 0; i < 10; ++i) {
    const LayerSpec &s = kLayers[i];
    LayerGeom &g = ns->g[i];
    g.Cin = s.Cin; g.Cout = s.Cout; g.k = s.k; g.stride = s.stride; g.pad = s.pad;
    g.Hin = h; g.Win = w;
    g.Ho = (h + 2 * s.pad - s.k) / s.stride + 1;
    g.Wo = (w + 2 * s.pad - s.k) / s.stride + 1;
    if (g.Ho < 1) g.Ho = 1;
    if (g.Wo < 1) g.Wo = 1;
    int Hp = h + 2 * s.pad, Wp = w + 2 * s.pad;
    if (s.stride == 2) { Hp += Hp & 1; Wp += Wp & 1; }
    g.py = g.px = s.pad;
    if (i == 0) {
      // conv1 (Cin = 8, 7x7 s2): space-to-depth -> 4x4 stride-1 taps over 32 channels
      g.rows = Hp / 2; g.cols = Wp / 2; g.Cbuf = 32;
      g.KH = g.KW = 4; g.stride_eff = 1; g.Ceff = 32; g.Hq = g.rows;
      g.BLOCK_K = 32; g.BLOCK_N = 64;
    } else {
      g.rows = Hp; g.cols = Wp; g.Cbuf = s.Cin;
      g.KH = g.KW = s.k; g.stride_eff = s.stride; g.Ceff = s.Cin;
      g.Hq = (s.stride == 2) ? Hp / 2 : Hp;
      g.BLOCK_K = 64; g.BLOCK_N = s.Cout >= 256 ? 256 : 128;
    }
    // M tile: BW x BH rectangle of output pixels with BW | Wo and BW*BH <= 128.  Maximise the rows
    // used; keep boxes at least 8 pixels wide (>= 1 KB contiguous per TMA row) when possible.
    int best_bw = 0, best_score = -1;
    for (int pass = 0; pass < 2 && best_bw == 0; ++pass)
      for (int d = 1; d <= g.Wo && d <= 128; ++d) {
        if (g.Wo % d) continue;
        if (pass == 0 && d < 8) continue;
        const int score = d * (128 / d) * 1000 + d;
        if (score > best_score) { best_score = score; best_bw = d; }
      }
    if (best_bw == 0) best_bw = 1;  // degenerate geometry (tiny geometry-only contexts)
    g.BW = best_bw; g.BH = 128 / best_bw;
    g.n_col_tiles = g.Wo > 0 ? g.Wo / g.BW : 0;
    if (i == 0) {  // conv1 strip kernel: one output row x BW columns per tile, BW = ceil(Wo / ceil(Wo/128))
      const int nt = cdiv(g.Wo, 128);
      g.BW = cdiv(g.Wo, nt); g.BH = 1; g.n_col_tiles = nt;
    }
    g.kblocks = g.KH * g.KW * (g.Ceff / g.BLOCK_K);
    g.occ = (i == 0) ? 2 : 1;
    g.pair = 0;  // per-layer kernel choice is applied per batch size in effective_geom() (NetState::pair_mask)
    h = g.Ho; w = g.Wo;
  }
}

// Kernel variant per layer.  bit i of NetState::pair_mask puts conv layer i (1..9) on the CTA-pair kernel
// (cta_group::2, 256 x BLOCK_N tiles: each CTA stages its 128 activation rows and HALF of the weight tile, so the
// shared-memory traffic per MMA drops from A + B to A + B/2 -- what bounds the N = 128 layer conv2).
static LayerGeom effective_geom(const NetState *ns, int i, int B) {
  LayerGeom g = ns->g[i];
  (void)B;
  if (i >= 1 && ((ns->pair_mask >> i) & 1)) g.pair = (g.BLOCK_N == 128) ? 2 : 1;  // 2: 3-stage ring, two pairs per SM pair
  if (i >= 1 && g.BLOCK_N == 128 && !g.pair) g.occ = 2;  // conv2 on the 1-CTA kernel: two CTAs per SM
  return g;
}

// split-K factor of the CTA-pair kernel (the 1-CTA persistent kernel never splits: see conv_igemm.cuh)
static int choose_ksplit(const NetState *ns, const LayerGeom &g, int B) {
  if (!g.pair) return 1;
  const int tiles = cdiv(cdiv(B * g.Hq, g.BH) * g.n_col_tiles, 2) * (g.Cout / g.BLOCK_N);
  const int cap = ns->num_sms / 2;
  if (tiles >= 2 * cap || g.kblocks < 16) return 1;
  int best = 1;
  double best_score = -1.0;
  for (int ks = 1; ks <= 8; ++ks) {
    if (ks > 1 && g.kblocks / ks < 8) break;
    if (ks > 1 && (ks - 1) * cdiv(g.kblocks, ks) >= g.kblocks) continue;  // no empty K slice
    const int t = tiles * ks;
    const double util = (double)t / (double)(cdiv(t, cap) * cap);
    const double score = util - 0.06 * (ks - 1);
    if (score > best_score + 1e-9) { best_score = score; best = ks; }
  }
  return best;
}

// --------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_map(CUtensorMap *m, void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                      const uint32_t *box, int block_k /*64: SW128, 32: SW64, 0: no swizzle*/) {
  EncodeTiledFn fn = get_encode();
  DIM_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t gd[5]; cuuint64_t gs[5]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu box %u %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0);
    return 3;
  }
  return 0;
}

uint32_t make_idesc(int M, int N, bool f16) {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format @7/@10 (0 = F16, 1 = BF16), K-major both,
  // n_dim = N>>3 @17, m_dim = M>>4 @24
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

static int build_maps(NetState *ns, int B, bool f16, TensorMaps &tm) {
  for (int i = 0; i < 10; ++i) {
    tm.g[i] = effective_geom(ns, i, B);
    const LayerGeom &g = tm.g[i];
    ConvKParams &kp = tm.kp[i];
    memset(&kp, 0, sizeof(kp));
    for (int lo = 0; lo < 2; ++lo) {
      __nv_bfloat16 *base = lo ? ns->act_lo[i] : ns->act_hi[i];
      CUtensorMap *maps = lo ? kp.a_lo_map : kp.a_map;
      const uint32_t box[3] = {(uint32_t)g.BLOCK_K, (uint32_t)g.BW, (uint32_t)g.BH};
      if (i == 0) {
        // conv1 strip layout [B*rows][4 chunks][cols][8 ch]: box = 8 ch x (BW+3) cols x 4 chunks x 1 row
        const uint64_t dims[4] = {8, (uint64_t)g.cols, 4, (uint64_t)B * g.rows};
        const uint64_t str[3] = {16, (uint64_t)g.cols * 16, (uint64_t)g.cols * 64};
        const uint32_t box4[4] = {8, (uint32_t)(g.BW + 3), 4, 1};
        if (int rc = encode_map(&maps[0], base, 4, dims, str, box4, 0)) return rc;
        maps[1] = maps[2] = maps[3] = maps[0];
      } else if (g.stride_eff == 1) {
        const uint64_t dims[3] = {(uint64_t)g.Cbuf, (uint64_t)g.cols, (uint64_t)B * g.rows};
        const uint64_t str[2] = {(uint64_t)g.Cbuf * 2, (uint64_t)g.cols * g.Cbuf * 2};
        if (int rc = encode_map(&maps[0], base, 3, dims, str, box, g.BLOCK_K)) return rc;
        maps[1] = maps[2] = maps[3] = maps[0];
      } else {
        for (int ph = 0; ph < 2; ++ph)
          for (int pw = 0; pw < 2; ++pw) {
            const uint64_t dims[3] = {(uint64_t)g.Cbuf, (uint64_t)g.cols / 2, (uint64_t)B * g.rows / 2};
            const uint64_t str[2] = {(uint64_t)2 * g.Cbuf * 2, (uint64_t)2 * g.cols * g.Cbuf * 2};
            __nv_bfloat16 *vb = base + ((size_t)ph * g.cols + pw) * g.Cbuf;
            if (int rc = encode_map(&maps[(ph << 1) | pw], vb, 3, dims, str, box, g.BLOCK_K)) return rc;
          }
      }
    }
    {
      const uint64_t Ktot = (uint64_t)g.KH * g.KW * g.Ceff;
      const uint64_t dims[2] = {Ktot, (uint64_t)g.Cout};
      const uint64_t str[1] = {Ktot * 2};
      const uint32_t box[2] = {(uint32_t)g.BLOCK_K, (uint32_t)g.BLOCK_N};
      __nv_bfloat16 *wop = f16 ? ns->w_f16[i] : ns->w_hi[i];
      if (int rc = encode_map(&kp.b_map, wop, 2, dims, str, box, g.BLOCK_K)) return rc;
      if (int rc = encode_map(&kp.b_lo_map, ns->w_lo[i], 2, dims, str, box, g.BLOCK_K)) return rc;
      const uint32_t box2[2] = {(uint32_t)g.BLOCK_K, (uint32_t)(g.BLOCK_N / 2)};
      if (int rc = encode_map(&kp.b2_map, wop, 2, dims, str, box2, g.BLOCK_K)) return rc;
      if (int rc = encode_map(&kp.b2_lo_map, ns->w_lo[i], 2, dims, str, box2, g.BLOCK_K)) return rc;
    }
    kp.KH = g.KH; kp.KW = g.KW; kp.stride = g.stride_eff; kp.cchunks = g.Ceff / g.BLOCK_K;
    kp.BW = g.BW; kp.BH = g.BH; kp.n_col_tiles = g.n_col_tiles;
    kp.Hq = g.Hq; kp.Ho = g.Ho; kp.Wo = g.Wo; kp.Bn = B;
    if (i < 9) {
      const LayerGeom &nx = ns->g[i + 1];
      kp.out_Hp = nx.rows; kp.out_Wp = nx.cols; kp.out_py = nx.py; kp.out_px = nx.px;
    } else {
      kp.out_Hp = g.Ho; kp.out_Wp = g.Wo; kp.out_py = 0; kp.out_px = 0;
    }
    kp.Cout = g.Cout;
    kp.kblocks = g.kblocks;
    kp.ksplit = tm.ksplit[i] = choose_ksplit(ns, g, B);
    kp.idesc = make_idesc(g.pair ? 256 : 128, g.BLOCK_N, f16);
    kp.f16 = f16 ? 1 : 0;
    kp.slope = 0.1f;
    kp.bias = ns->bias[i];
    kp.out_hi = ns->act_hi[i + 1];
    kp.out_lo = ns->act_lo[i + 1];
    kp.partial = ns->conv_partial;
  }
  return 0;
}

// --------------------------------------------------------------------------------- fc6 + head
// fc6 split-K on the legacy tensor path: the batch (<= 16 instances) is exactly the M = 16 of
// mma.sync.m16n8k16, so each warp streams its weight rows once (16 B per lane, K-permuted so that the
// 8 consecutive bf16 a lane loads feed two MMA steps) and multiplies them against the activation
// chunk staged in shared memory.  HBM-bound by design: 42 MB of bf16 weights per call (84 MB with the
// lo halves in bf16x3 mode).  CTA s owns k in [s*KC, (s+1)*KC) for all 256 outputs; partials are
// reduced in fixed order by the head kernel (deterministic, no atomics).
template <bool F16>
__device__ __forceinline__ void mma_bf16_16816(float *c, const uint32_t *a, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// S3: bf16 hi/lo operands (3 MMAs per step); F16: IEEE half operands (one MMA per step, like plain bf16)
template <bool S3, bool F16 = false>
__global__ void __launch_bounds__(256) fc6_mma_kernel(const __nv_bfloat16 *__restrict__ act_hi,
                                                      const __nv_bfloat16 *__restrict__ act_lo,
                                                      const __nv_bfloat16 *__restrict__ w_hi,
                                                      const __nv_bfloat16 *__restrict__ w_lo, int B, int max_batch,
                                                      float *__restrict__ partial) {
  constexpr int PITCH = FC6_KC * 2 + 64;  // bytes; +64 keeps the 16-B fragment loads conflict-free
  __shared__ __align__(16) uint8_t a_hi_s[16 * PITCH];
  __shared__ __align__(16) uint8_t a_lo_s[S3 ? 16 * PITCH : 16];
  const int s = blockIdx.x, k0 = s * FC6_KC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  for (int b0 = 0; b0 < B; b0 += 16) {
    const int nb = min(16, B - b0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < 16 * (FC6_KC / 8); idx += blockDim.x) {
      const int r = idx / (FC6_KC / 8), c8 = idx - r * (FC6_KC / 8);
      uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
      if (r < nb) {
        const size_t o = (size_t)(b0 + r) * FC6_K + k0 + c8 * 8;
        vh = *reinterpret_cast<const uint4 *>(act_hi + o);
        if (S3) vl = *reinterpret_cast<const uint4 *>(act_lo + o);
      }
      *reinterpret_cast<uint4 *>(a_hi_s + r * PITCH + c8 * 16) = vh;
      if (S3) *reinterpret_cast<uint4 *>(a_lo_s + r * PITCH + c8 * 16) = vl;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
#pragma unroll 2
    for (int kc = 0; kc < FC6_KC; kc += 32) {
      // activation fragments for both MMA steps of this 32-k chunk: rows g and g+8, 8 bf16 each
      const uint4 ah0 = *reinterpret_cast<const uint4 *>(a_hi_s + g * PITCH + (kc + q * 8) * 2);
      const uint4 ah1 = *reinterpret_cast<const uint4 *>(a_hi_s + (g + 8) * PITCH + (kc + q * 8) * 2);
      const uint32_t A0[4] = {ah0.x, ah1.x, ah0.y, ah1.y}, A1[4] = {ah0.z, ah1.z, ah0.w, ah1.w};
      uint32_t L0[4] = {0, 0, 0, 0}, L1[4] = {0, 0, 0, 0};
      if (S3) {
        const uint4 al0 = *reinterpret_cast<const uint4 *>(a_lo_s + g * PITCH + (kc + q * 8) * 2);
        const uint4 al1 = *reinterpret_cast<const uint4 *>(a_lo_s + (g + 8) * PITCH + (kc + q * 8) * 2);
        L0[0] = al0.x; L0[1] = al1.x; L0[2] = al0.y; L0[3] = al1.y;
        L1[0] = al0.z; L1[1] = al1.z; L1[2] = al0.w; L1[3] = al1.w;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = warp * 32 + t * 8 + g;  // output row whose weights this lane streams
        const size_t wo = (size_t)j * FC6_K + k0 + kc + q * 8;
        const uint4 wh = __ldg(reinterpret_cast<const uint4 *>(w_hi + wo));
        mma_bf16_16816<F16>(acc[t], A0, wh.x, wh.y);
        mma_bf16_16816<F16>(acc[t], A1, wh.z, wh.w);
        if (S3) {
          const uint4 wl = __ldg(reinterpret_cast<const uint4 *>(w_lo + wo));
          mma_bf16_16816<false>(acc[t], L0, wh.x, wh.y);
          mma_bf16_16816<false>(acc[t], L1, wh.z, wh.w);
          mma_bf16_16816<false>(acc[t], A0, wl.x, wl.y);
          mma_bf16_16816<false>(acc[t], A1, wl.z, wl.w);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = warp * 32 + t * 8 + q * 2;
      if (g < nb)
        *reinterpret_cast<float2 *>(partial + ((size_t)s * max_batch + b0 + g) * 256 + j) = make_float2(acc[t][0], acc[t][1]);
      if (g + 8 < nb)
        *reinterpret_cast<float2 *>(partial + ((size_t)s * max_batch + b0 + g + 8) * 256 + j) = make_float2(acc[t][2], acc[t][3]);
    }
  }
}

// one CTA of 1024 threads per instance: 4 thread groups share the fc6 split reduction and the fc7
// reduction (fixed summation order -> deterministic)
__global__ void __launch_bounds__(1024) head_kernel(const float *__restrict__ partial, int max_batch,
                                                    const float *__restrict__ fc6_b, const float *__restrict__ fc7_wT,
                                                    const float *__restrict__ fc7_b, const float *__restrict__ rot_w,
                                                    const float *__restrict__ rot_b, const float *__restrict__ trans_w,
                                                    const float *__restrict__ trans_b,
                                                    const float *__restrict__ zoom_factor /*nullable*/,
                                                    float *__restrict__ rot_out, float *__restrict__ trans_out,
                                                    float *__restrict__ se3_out, float *__restrict__ h6_out,
                                                    float *__restrict__ h7_out) {
  __shared__ float red[4][256];
  __shared__ float h6[256], h7[256], outv[8];
  const int b = blockIdx.x, j = threadIdx.x & 255, grp = threadIdx.x >> 8;
  float a = 0.f;
#pragma unroll 8
  for (int s = grp; s < FC6_SPLITS; s += 4) a += partial[((size_t)s * max_batch + b) * 256 + j];
  red[grp][j] = a;
  __syncthreads();
  if (grp == 0) {
    const float v = (((red[0][j] + red[1][j]) + red[2][j]) + red[3][j]) + fc6_b[j];
    h6[j] = v > 0.f ? v : 0.1f * v;
  }
  __syncthreads();
  float c = 0.f;
#pragma unroll 8
  for (int k = grp * 64; k < grp * 64 + 64; ++k) c = fmaf(h6[k], fc7_wT[k * 256 + j], c);
  red[grp][j] = c;
  __syncthreads();
  if (grp == 0) {
    const float v = (((red[0][j] + red[1][j]) + red[2][j]) + red[3][j]) + fc7_b[j];
    h7[j] = v > 0.f ? v : 0.1f * v;
    if (h6_out) { h6_out[b * 256 + j] = h6[j]; h7_out[b * 256 + j] = h7[j]; }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 7) {
    const float *wrow = warp < 4 ? rot_w + warp * 256 : trans_w + (warp - 4) * 256;
    float v = 0.f;
    for (int k = lane; k < 256; k += 32) v = fmaf(h7[k], wrow[k], v);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) outv[warp] = v + (warp < 4 ? rot_b[warp] : trans_b[warp - 4]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (rot_out)
      for (int k = 0; k < 4; ++k) rot_out[4 * b + k] = outv[k];
    if (trans_out)
      for (int k = 0; k < 3; ++k) trans_out[3 * b + k] = outv[4 + k];
    if (se3_out) {
      // invZoomTrans (zoom_trans.py:30-31, b_inv_zoom=True): wx for both axes
      const float w = zoom_factor[4 * b];
      for (int k = 0; k < 4; ++k) se3_out[7 * b + k] = outv[k];
      se3_out[7 * b + 4] = outv[4] * w;
      se3_out[7 * b + 5] = outv[5] * w;
      se3_out[7 * b + 6] = outv[6];
    }
  }
}

// --------------------------------------------------------------------------------- host API
int net_create(dim_ctx *ctx) {
  NetState *ns = new NetState();
  ctx->net = ns;
  ns->max_batch = ctx->max_batch;
  ns->num_sms = ctx->num_sms;
  build_geometry(ns, ctx->H, ctx->W);
  if ((size_t)ns->g[9].Ho * ns->g[9].Wo * ns->g[9].Cout != (size_t)FC6_K) return 0;  // geometry-only context
  for (int i = 0; i <= 10; ++i) {
    size_t per;
    if (i < 10) per = (size_t)ns->g[i].rows * ns->g[i].cols * ns->g[i].Cbuf;
    else per = (size_t)ns->g[9].Ho * ns->g[9].Wo * ns->g[9].Cout;
    ns->act_elems_per_image[i] = per;
    // borders must be (and stay) zero: the epilogues only ever write interior pixels
    if (int rc = dev_alloc(ctx, &ns->act_hi[i], per * ctx->max_batch, true)) return rc;
    if (int rc = dev_alloc(ctx, &ns->act_lo[i], per * ctx->max_batch, true)) return rc;
  }
  ns->net_ok = ns->act_elems_per_image[10] == (size_t)FC6_K;  // fc6 is 81920 -> 256: needs 480x640
  size_t pmax = 0;  // split-K partials of the CTA-pair kernel: sized for every layer on it (the choice may change at run time)
  const int mask_now = ns->pair_mask;
  ns->pair_mask = 0x3FE;
  for (int B = 1; B <= ctx->max_batch; ++B)
    for (int i = 0; i < 10; ++i) {
      int ks = choose_ksplit(ns, effective_geom(ns, i, B), B);
      if (ks > 1) {
        size_t e = (size_t)ks * B * ns->g[i].Ho * ns->g[i].Wo * ns->g[i].Cout;
        pmax = e > pmax ? e : pmax;
      }
    }
  ns->pair_mask = mask_now;
  ns->conv_partial_elems = pmax;
  if (pmax)
    if (int rc = dev_alloc(ctx, &ns->conv_partial, pmax, false)) return rc;
  if (int rc = dev_alloc(ctx, &ns->fc6_partial, (size_t)FC6_SPLITS * ctx->max_batch * 256, true)) return rc;
  return 0;
}

void net_destroy(dim_ctx *ctx) {
  if (ctx->net && ctx->net->layer_events) {
    for (int i = 0; i < 11; ++i) cudaEventDestroy(ctx->net->layer_events[i]);
    delete[] ctx->net->layer_events;
  }
  delete ctx->net;
  ctx->net = nullptr;
}

static void split_bf16(const float *src, size_t n, std::vector<__nv_bfloat16> &hi, std::vector<__nv_bfloat16> &lo,
                       std::vector<__nv_bfloat16> &f16) {
  hi.resize(n);
  lo.resize(n);
  f16.resize(n);
  for (size_t i = 0; i < n; ++i) {
    hi[i] = __float2bfloat16_rn(src[i]);
    lo[i] = __float2bfloat16_rn(src[i] - __bfloat162float(hi[i]));
    const __half h = __float2half_rn(src[i]);  // He-scale weights are far inside the fp16 range; inf only for |w| > 65504
    memcpy(&f16[i], &h, 2);
  }
}

// first call allocates, later calls (dim_net_load on a loaded context) overwrite in place: the tensor maps
// cached in NetState::maps keep pointing at valid storage
template <typename T>
static int upload(dim_ctx *ctx, T **dst, const std::vector<T> &v) {
  if (*dst == nullptr)
    if (int rc = dev_alloc(ctx, dst, v.size(), false)) return rc;
  DIM_CHECK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// fp16 operand packs re-derived on the device from the bf16 hi/lo pair (hi + lo carries 16 significant bits): used
// after a training update refreshed hi/lo from the fp32 master weights (train.cu repack_all)
__global__ void __launch_bounds__(256) f16_from_hilo_kernel(const __nv_bfloat16 *hi, const __nv_bfloat16 *lo, size_t n,
                                                            __nv_bfloat16 *out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const __half h = __float2half_rn(__bfloat162float(hi[i]) + __bfloat162float(lo[i]));
  reinterpret_cast<__half *>(out)[i] = h;
}

static int net_refresh_f16(dim_ctx *ctx, cudaStream_t st) {
  NetState *ns = ctx->net;
  if (ns->lo_stale)
    if (int rc = train_refresh_lo(ctx, st)) return rc;
  for (int i = 0; i < 10; ++i) {
    const LayerGeom &g = ns->g[i];
    const size_t n = (size_t)g.Cout * g.KH * g.KW * g.Ceff;
    f16_from_hilo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ns->w_hi[i], ns->w_lo[i], n, ns->w_f16[i]);
    DIM_LAUNCH_CHECK();
  }
  const size_t n6 = (size_t)256 * FC6_K;
  f16_from_hilo_kernel<<<(unsigned)((n6 + 255) / 256), 256, 0, st>>>(ns->fc6_w_hi, ns->fc6_w_lo, n6, ns->fc6_w_f16);
  DIM_LAUNCH_CHECK();
  ns->f16_stale = false;
  return 0;
}

int net_load(dim_ctx *ctx, const float *const *W, const float *const *Bv) {
  NetState *ns = ctx->net;
  DIM_REQUIRE(ns != nullptr, "net not created");
  DIM_REQUIRE(ns->net_ok, "dim_net_load: FlowNetS + fc6 (81920 inputs) needs a 480x640 context");
  DIM_REQUIRE(!ns->train_aliased, "dim_net_load: this context trains; use dim_train_load_params (it owns the weights)");
  if (ns->loaded) DIM_CHECK(cudaDeviceSynchronize());  // a reload must not race kernels still reading the old weights
  for (int i = 0; i < 10; ++i) {
    const LayerGeom &g = ns->g[i];
    const size_t Ktot = (size_t)g.KH * g.KW * g.Ceff;
    std::vector<float> packed((size_t)g.Cout * Ktot, 0.f);
    const float *w = W[i];  // [Cout][Cin][k][k]
    if (i == 0) {
      // space-to-depth repack: W'[co][dh][dw][conv1_kslot(dw,ph,pw)+c] = W[co][c][2dh+ph][2dw+pw] (0 beyond 7x7)
      for (int co = 0; co < g.Cout; ++co)
        for (int dh = 0; dh < 4; ++dh)
          for (int dw = 0; dw < 4; ++dw)
            for (int ph = 0; ph < 2; ++ph)
              for (int pw = 0; pw < 2; ++pw)
                for (int c = 0; c < 8; ++c) {
                  const int kh = 2 * dh + ph, kw = 2 * dw + pw;
                  if (kh >= 7 || kw >= 7) continue;
                  packed[(size_t)co * Ktot + (size_t)(dh * 4 + dw) * 32 + conv1_kslot(dw, ph, pw) + c] =
                      w[(((size_t)co * 8 + c) * 7 + kh) * 7 + kw];
                }
    } else {
      for (int co = 0; co < g.Cout; ++co)
        for (int c = 0; c < g.Cin; ++c)
          for (int kh = 0; kh < g.k; ++kh)
            for (int kw = 0; kw < g.k; ++kw)
              packed[(size_t)co * Ktot + (size_t)(kh * g.k + kw) * g.Cin + c] =
                  w[(((size_t)co * g.Cin + c) * g.k + kh) * g.k + kw];
    }
    std::vector<__nv_bfloat16> hi, lo, hf;
    split_bf16(packed.data(), packed.size(), hi, lo, hf);
    if (int rc = upload(ctx, &ns->w_hi[i], hi)) return rc;
    if (int rc = upload(ctx, &ns->w_lo[i], lo)) return rc;
    if (int rc = upload(ctx, &ns->w_f16[i], hf)) return rc;
    std::vector<float> bv(Bv[i], Bv[i] + g.Cout);
    if (int rc = upload(ctx, &ns->bias[i], bv)) return rc;
  }
  {  // fc6: (out, c*80 + h*10 + w) -> (out, (h*10+w)*1024 + c)   (NCHW flatten, deepIM_flownet.py:110)
    std::vector<float> p((size_t)256 * FC6_K);
    for (int o = 0; o < 256; ++o)
      for (int c = 0; c < 1024; ++c)
        for (int hw = 0; hw < 80; ++hw) p[(size_t)o * FC6_K + (size_t)hw * 1024 + c] = W[10][(size_t)o * FC6_K + (size_t)c * 80 + hw];
    std::vector<__nv_bfloat16> hb, lb, fb;
    split_bf16(p.data(), p.size(), hb, lb, fb);
    if (int rc = upload(ctx, &ns->fc6_w_hi, hb)) return rc;
    if (int rc = upload(ctx, &ns->fc6_w_lo, lb)) return rc;
    if (int rc = upload(ctx, &ns->fc6_w_f16, fb)) return rc;
    if (int rc = upload(ctx, &ns->fc6_b, std::vector<float>(Bv[10], Bv[10] + 256))) return rc;
    std::vector<float> t((size_t)256 * 256);
    for (int o = 0; o < 256; ++o)
      for (int k = 0; k < 256; ++k) t[(size_t)k * 256 + o] = W[11][(size_t)o * 256 + k];
    if (int rc = upload(ctx, &ns->fc7_wT, t)) return rc;
    if (int rc = upload(ctx, &ns->fc7_b, std::vector<float>(Bv[11], Bv[11] + 256))) return rc;
    if (int rc = upload(ctx, &ns->rot_w, std::vector<float>(W[12], W[12] + 4 * 256))) return rc;
    if (int rc = upload(ctx, &ns->rot_b, std::vector<float>(Bv[12], Bv[12] + 4))) return rc;
    if (int rc = upload(ctx, &ns->trans_w, std::vector<float>(W[13], W[13] + 3 * 256))) return rc;
    if (int rc = upload(ctx, &ns->trans_b, std::vector<float>(Bv[13], Bv[13] + 3))) return rc;
  }
  ns->loaded = true;
  return 0;
}

template <int BN, int BK, int ST, bool S3, bool RES, int KRES>
static int launch_conv2(NetState *ns, const ConvKParams &kp, int total_tiles, int n_tiles, int cap, cudaStream_t st) {
  using S = ConvSmem2<BN, BK, ST, S3, RES, KRES>;
  static bool attr_set = false;
  if (!attr_set) {
    DIM_CHECK(cudaFuncSetAttribute(conv_igemm_persistent_kernel<BN, BK, ST, S3, RES, KRES>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  const int grid = total_tiles < cap ? total_tiles : cap;
  conv_igemm_persistent_kernel<BN, BK, ST, S3, RES, KRES><<<grid, 192, S::TOTAL, st>>>(kp, total_tiles, n_tiles);
  DIM_LAUNCH_CHECK();
  return 0;
}

template <int BN, int ST, bool S3>
static int launch_pair(const ConvKParams &kp, int pair_tiles, int n_tiles, int sms, cudaStream_t st) {
  using S = ConvSmemPair<BN, ST, S3>;
  static bool attr_set = false;
  if (!attr_set) {
    DIM_CHECK(cudaFuncSetAttribute(conv_igemm_pair_kernel<BN, ST, S3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   S::TOTAL));
    attr_set = true;
  }
  const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
  conv_igemm_pair_kernel<BN, ST, S3><<<2 * pairs, 192, S::TOTAL, st>>>(kp, pair_tiles, n_tiles);  // __cluster_dims__(2,1,1)
  DIM_LAUNCH_CHECK();
  return 0;
}

// where conv1's input buffer expects pixel (i,j) of the 8-channel blob (space-to-depth, pad 3)
void net_input_geometry(dim_ctx *ctx, int *rows, int *cols, int *pad, __nv_bfloat16 **hi, __nv_bfloat16 **lo) {
  NetState *ns = ctx->net;
  *rows = ns->g[0].rows; *cols = ns->g[0].cols; *pad = ns->g[0].py;
  *hi = ns->act_hi[0]; *lo = ns->act_lo[0];
}

// runs conv tower + fc6 + head on the already-filled conv1 input buffer
int net_forward(dim_ctx *ctx, int B, int precision, const float *zoom_factor, float *rot_out, float *trans_out,
                float *se3_out, cudaStream_t st, cudaEvent_t after_conv) {
  NetState *ns = ctx->net;
  DIM_REQUIRE(ns && ns->loaded, "dim_net_load has not been called");
  DIM_REQUIRE(B >= 1 && B <= ctx->max_batch, "batch exceeds max_batch");
  DIM_REQUIRE(precision == DIM_PREC_BF16 || precision == DIM_PREC_BF16X3 || precision == DIM_PREC_FP16,
              "unknown precision (DIM_PREC_BF16 / DIM_PREC_BF16X3 / DIM_PREC_FP16)");
  const bool s3 = precision == DIM_PREC_BF16X3, f16 = precision == DIM_PREC_FP16;
  const int key = B + (f16 ? kF16MapKey : 0);
  auto it = ns->maps.find(key);
  if (it == ns->maps.end()) {
    TensorMaps tm;
    if (int rc = build_maps(ns, B, f16, tm)) return rc;
    it = ns->maps.emplace(key, tm).first;
  }
  const TensorMaps &tm = it->second;
  if (ns->repack_done) DIM_CHECK(cudaStreamWaitEvent(st, ns->repack_done, 0));
  if (s3 && ns->lo_stale)
    if (int rc = train_refresh_lo(ctx, st)) return rc;
  if (f16 && ns->f16_stale)
    if (int rc = net_refresh_f16(ctx, st)) return rc;
  if (ns->layer_events) DIM_CHECK(cudaEventRecord(ns->layer_events[0], st));
  for (int i = 0; i < 10; ++i) {
    const LayerGeom &g = tm.g[i];
    const ConvKParams &kp = tm.kp[i];
    const int n_tiles = g.Cout / g.BLOCK_N;
    const int total_tiles = cdiv(B * g.Hq, g.BH) * g.n_col_tiles * n_tiles * kp.ksplit;
    const int sms = ns->num_sms;
    int rc;
    if (i == 0) {
      // conv1: a CTA walks down a run of output rows of one column tile, one new input strip per row
      const int rows_total = B * g.Hq;
      int chunks = sms / g.n_col_tiles;
      if (chunks > cdiv(rows_total, 16)) chunks = cdiv(rows_total, 16);  // keep the 3-row halo below ~20 %
      if (chunks < 1) chunks = 1;
      const int rpc = cdiv(rows_total, chunks);
      chunks = cdiv(rows_total, rpc);
      const int strip_bytes = cdiv((g.BW + 3) * 64, 128) * 128;
      const int grid = g.n_col_tiles * chunks;
      if (s3) {  // hi/lo operands: rolling strips, three MMA passes per step
        constexpr int ST = 5;
        const int smem_bytes = 2 * 16 * 4096 + ST * 2 * strip_bytes + (8 * 2048 + 256) + 1024 + 512;
        DIM_REQUIRE(smem_bytes <= 227 * 1024, "conv1 (bf16x3): image too wide for the rolling-strip ring");
        static int set1 = 0;
        if (set1 < smem_bytes) { DIM_CHECK(cudaFuncSetAttribute(conv1_roll_kernel<ST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); set1 = smem_bytes; }
        conv1_roll_kernel<ST, true><<<grid, 320, smem_bytes, st>>>(kp, rows_total, rpc, chunks, strip_bytes);
      } else if (ns->conv1_stack) {
        // stacked filter rows: one 128 x 256 MMA per (dw, k) step feeds four output rows (A read once instead of four times)
        constexpr int ST = 6;
        const int smem_bytes = 16 * 4096 + ST * strip_bytes + (8 * 2048 + 256) + 512 + 512;
        static int set3 = 0;
        if (set3 < smem_bytes) { DIM_CHECK(cudaFuncSetAttribute(conv1_stack_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); set3 = smem_bytes; }
        conv1_stack_kernel<ST><<<grid, 320, smem_bytes, st>>>(kp, rows_total, rpc, chunks, strip_bytes);
      } else {
        constexpr int ST = 8;
        const int smem_bytes = 16 * 4096 + ST * strip_bytes + (8 * 2048 + 256) + 1024 + 512;
        static int set0 = 0;
        if (set0 < smem_bytes) { DIM_CHECK(cudaFuncSetAttribute(conv1_roll_kernel<ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); set0 = smem_bytes; }
        conv1_roll_kernel<ST, false><<<grid, 320, smem_bytes, st>>>(kp, rows_total, rpc, chunks, strip_bytes);
      }
      DIM_LAUNCH_CHECK();
      rc = 0;
    }
    else if (g.pair) {
      const int m_tiles = cdiv(B * g.Hq, g.BH) * g.n_col_tiles;
      const int pair_tiles = cdiv(m_tiles, 2) * n_tiles * kp.ksplit;
      if (g.BLOCK_N == 128 && g.pair == 2 && !s3)
        rc = launch_pair<128, 3, false>(kp, pair_tiles, n_tiles, 2 * sms, st);
      else if (g.BLOCK_N == 128)
        rc = s3 ? launch_pair<128, 4, true>(kp, pair_tiles, n_tiles, sms, st) : launch_pair<128, 8, false>(kp, pair_tiles, n_tiles, sms, st);
      else
        rc = s3 ? launch_pair<256, 3, true>(kp, pair_tiles, n_tiles, sms, st) : launch_pair<256, 6, false>(kp, pair_tiles, n_tiles, sms, st);
    } else if (g.BLOCK_N <= 128) {
      rc = s3 ? launch_conv2<128, 64, 3, true, false, 0>(ns, kp, total_tiles, n_tiles, sms, st)
              : (g.occ == 2 ? launch_conv2<128, 64, 2, false, false, 0>(ns, kp, total_tiles, n_tiles, 2 * sms, st)
                            : launch_conv2<128, 64, 5, false, false, 0>(ns, kp, total_tiles, n_tiles, sms, st));
    }
    else
      rc = s3 ? launch_conv2<256, 64, 2, true, false, 0>(ns, kp, total_tiles, n_tiles, sms, st)
              : launch_conv2<256, 64, 4, false, false, 0>(ns, kp, total_tiles, n_tiles, sms, st);
    if (rc) return rc;
    if (ns->layer_events && i < 9) DIM_CHECK(cudaEventRecord(ns->layer_events[i + 1], st));
    if (kp.ksplit > 1) {
      const int npix = B * g.Ho * g.Wo;
      const size_t n4 = (size_t)npix * g.Cout / 4;
      conv_splitk_finalize_kernel<<<(unsigned)cdiv((int)n4, 256), 256, 0, st>>>(
          ns->conv_partial, kp.ksplit, npix, g.Cout, g.Ho, g.Wo, kp.out_Hp, kp.out_Wp, kp.out_py, kp.out_px,
          ns->bias[i], 0.1f, ns->act_hi[i + 1], s3 ? ns->act_lo[i + 1] : nullptr, f16 ? 1 : 0);
      DIM_LAUNCH_CHECK();
    }
  }
  if (ns->layer_events) DIM_CHECK(cudaEventRecord(ns->layer_events[10], st));
  if (after_conv) DIM_CHECK(cudaEventRecord(after_conv, st));
  if (s3)
    fc6_mma_kernel<true><<<FC6_SPLITS, 256, 0, st>>>(ns->act_hi[10], ns->act_lo[10], ns->fc6_w_hi, ns->fc6_w_lo, B,
                                                     ctx->max_batch, ns->fc6_partial);
  else if (f16)
    fc6_mma_kernel<false, true><<<FC6_SPLITS, 256, 0, st>>>(ns->act_hi[10], nullptr, ns->fc6_w_f16, nullptr, B,
                                                            ctx->max_batch, ns->fc6_partial);
  else
    fc6_mma_kernel<false><<<FC6_SPLITS, 256, 0, st>>>(ns->act_hi[10], nullptr, ns->fc6_w_hi, nullptr, B,
                                                      ctx->max_batch, ns->fc6_partial);
  DIM_LAUNCH_CHECK();
  head_kernel<<<B, 1024, 0, st>>>(ns->fc6_partial, ctx->max_batch, ns->fc6_b, ns->fc7_wT, ns->fc7_b, ns->rot_w,
                                 ns->rot_b, ns->trans_w, ns->trans_b, zoom_factor, rot_out, trans_out, se3_out, ns->save_h6,
                                 ns->save_h7);
  DIM_LAUNCH_CHECK();
  return 0;
}

// the refinement chain may be captured into a CUDA graph only when nothing host-dependent hangs on the forward pass
bool net_graph_safe(dim_ctx *ctx) {
  NetState *ns = ctx->net;
  return ns && ns->loaded && !ns->layer_events && !ns->train_aliased;
}

// tuning hook (tools/conv_lab.py): kernel-variant switches at run time; cached launch descriptors are rebuilt
int net_set_option(dim_ctx *ctx, const char *key, int value) {
  NetState *ns = ctx->net;
  DIM_REQUIRE(ns != nullptr, "net not created");
  if (!strcmp(key, "pair_mask")) ns->pair_mask = value & 0x3FE;
  else if (!strcmp(key, "conv1_stack")) ns->conv1_stack = value != 0;
  else { set_error("dim_debug_set_option: unknown key '%s'", key); return 2; }
  DIM_CHECK(cudaDeviceSynchronize());
  ns->maps.clear();
  return 0;
}

// tuning hook: per-layer device times of the LAST net_forward (11 events: before conv1, after each of the 10 layers)
int net_layer_profile(dim_ctx *ctx, int enable, float *ms10) {
  NetState *ns = ctx->net;
  DIM_REQUIRE(ns != nullptr, "net not created");
  if (enable && !ns->layer_events) {
    ns->layer_events = new cudaEvent_t[11];
    for (int i = 0; i < 11; ++i) DIM_CHECK(cudaEventCreate(&ns->layer_events[i]));
  }
  if (ms10 && ns->layer_events) {
    DIM_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < 10; ++i) DIM_CHECK(cudaEventElapsedTime(&ms10[i], ns->layer_events[i], ns->layer_events[i + 1]));
  }
  if (!enable && ns->layer_events) {
    for (int i = 0; i < 11; ++i) cudaEventDestroy(ns->layer_events[i]);
    delete[] ns->layer_events;
    ns->layer_events = nullptr;
  }
  return 0;
}

// debugging / test hook: copy the bf16 activation buffer (input of layer `idx`, 10 = fc6 input) to host
int net_debug_activation(dim_ctx *ctx, int idx, int lo, void *host_dst, size_t bytes) {
  NetState *ns = ctx->net;
  DIM_REQUIRE(ns && idx >= 0 && idx <= 10, "bad activation index");
  const size_t have = ns->act_elems_per_image[idx] * ctx->max_batch * 2;
  DIM_REQUIRE(bytes <= have, "activation copy larger than buffer");
  DIM_CHECK(cudaMemcpy(host_dst, lo ? ns->act_lo[idx] : ns->act_hi[idx], bytes, cudaMemcpyDeviceToHost));
  return 0;
}

void net_layer_geometry(dim_ctx *ctx, int idx, int *out /*rows, cols, Cbuf, py, px, Ho, Wo, Cout*/) {
  NetState *ns = ctx->net;
  if (idx < 10) {
    const LayerGeom &g = ns->g[idx];
    out[0] = g.rows; out[1] = g.cols; out[2] = g.Cbuf; out[3] = g.py; out[4] = g.px; out[5] = g.Ho; out[6] = g.Wo;
    out[7] = g.Cout;
  } else {
    const LayerGeom &g = ns->g[9];
    out[0] = g.Ho; out[1] = g.Wo; out[2] = g.Cout; out[3] = 0; out[4] = 0; out[5] = 0; out[6] = 0; out[7] = 256;
  }
}

}  // namespace dim
