// conv_wgrad.cuh -- weight gradients of the convolutions / deconvolutions on tcgen05 (training step,
// SURVEY 8 row a10; replaces cuDNN's backward-filter behind mx.symbol.Convolution/Deconvolution).
//
//   dW[m][tap][n] = sum over pixels (b, y, x) of  Z[b, y, x, m] * A[b, y*s + kh, x*s + kw, n]
//
// i.e. per filter tap one GEMM with M = channels of Z (the output gradient dZ for a convolution, the input
// activation for a deconvolution), N = channels of A and the pixel index as the contraction dimension.
// Both operands are NHWC, so the contraction index is the STRIDED one: the tiles are MN-major UMMA
// operands (instruction-descriptor bits 15/16), which is exactly what a TMA box {64 channels, BW, BH}
// with the 128-byte swizzle produces (row = pixel, 128 B = 64 channels):
//     canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units  (cute mma_traits_sm100 make_umma_desc<MN>)
//     -> LBO = bytes between 64-channel groups (one TMA box = 8 KB), SBO = 1024 (8 pixel rows of 128 B).
// A K block is a BW x BH = 64 pixel rectangle of ONE image (4-D tensor maps carry the batch index, so a
// block never straddles images); rows/cols beyond the buffer are zero-filled by TMA, the zero border of the
// Z buffer makes overhanging pixels contribute nothing.
// Work item = (K slice, tap, M tile, N tile); fp32 partial tiles are reduced (fixed order -> deterministic)
// and re-laid out to the MXNet parameter layout by wgrad_reduce_kernel.
#pragma once
#include "conv_igemm.cuh"

namespace dim {

struct WgradParams {
  CUtensorMap z_map;     // (C, cols, rows, B), box {64, BW, BH, 1}, SWIZZLE_128B
  CUtensorMap a_map[4];  // (C, cols, rows, B) views ((row parity, col parity) for stride 2), box {min(BN,64), BW, BH, 1}
  int KH, KW, stride;
  int BW, BH, rects_x, rects_y, Bn;
  int z_off_r, z_off_c, a_off_r, a_off_c;
  int m_tiles, n_tiles, kslices, kb_per_slice, kb_total;
  uint32_t idesc;
  float *partial;  // [kslices][taps][m_tiles*128][n_tiles*BN]
};

namespace ptx {
// MN-major swizzled operand descriptor: LBO = stride between 64-element (SW128) / 32-element (SW64) groups along
// M/N, SBO = stride between groups of 8 K rows.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
}  // namespace ptx

template <int BN, int STAGES>
struct WgradSmem {
  static constexpr int A_BYTES = 2 * 8192;                     // 128 channels x 64 pixels
  static constexpr int B_BYTES = BN >= 64 ? (BN / 64) * 8192 : 4096;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using S = WgradSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * S::STAGE_BYTES);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *done_bar = empty_bar + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.KH * p.KW;
  int w = blockIdx.x;
  const int nt = w % p.n_tiles; w /= p.n_tiles;
  const int mt = w % p.m_tiles; w /= p.m_tiles;
  const int tap = w % taps;
  const int slice = w / taps;
  const int kb0 = slice * p.kb_per_slice;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&p.z_map);
    ptx::prefetch_tmap(&p.a_map[0]);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {  // whole warp; the elected lane issues (conv_igemm.cuh ptx::elect_one)
      const int kh = tap / p.KW, kw = tap - kh * p.KW;
      int view = 0, dr = kh, dc = kw;
      if (p.stride == 2) {
        view = ((kh & 1) << 1) | (kw & 1);
        dr = kh >> 1;
        dc = kw >> 1;
      }
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        const int rx = kb % p.rects_x;
        const int t = kb / p.rects_x;
        const int ry = t % p.rects_y, b = t / p.rects_y;
        const int oy0 = ry * p.BH, ox0 = rx * p.BW;
        uint8_t *st = smem + s * S::STAGE_BYTES;
        ptx::mbar_expect_tx(&full_bar[s], (uint32_t)S::STAGE_BYTES);
        tma_load_4d(st, &p.z_map, &full_bar[s], mt * 128, ox0 + p.z_off_c, oy0 + p.z_off_r, b);
        tma_load_4d(st + 8192, &p.z_map, &full_bar[s], mt * 128 + 64, ox0 + p.z_off_c, oy0 + p.z_off_r, b);
        uint8_t *bs = st + S::A_BYTES;
        if (BN >= 64) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(bs + j * 8192, &p.a_map[view], &full_bar[s], nt * BN + j * 64, ox0 + dc + p.a_off_c,
                        oy0 + dr + p.a_off_r, b);
        } else {
          tma_load_4d(bs, &p.a_map[view], &full_bar[s], nt * BN, ox0 + dc + p.a_off_c, oy0 + dr + p.a_off_r, b);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {  // whole warp; the elected lane issues
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b0 = a0 + S::A_BYTES;
        if (ptx::elect_one()) {  // one election per K-block; descriptors advance by (bytes >> 4)
          const uint64_t da0 = ptx::umma_desc_mn(a0, 8192, 1024, 2u);
          const uint64_t db0 = BN >= 64 ? ptx::umma_desc_mn(b0, 8192, 1024, 2u) : ptx::umma_desc_mn(b0, 4096, 512, 4u);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 64 pixels = 4 x UMMA_K 16
            ptx::umma_f16_raw(tmem_base, da0 + (uint64_t)(k * 128), db0 + (uint64_t)(k * (BN >= 64 ? 128 : 64)), p.idesc,
                              (kb > kb0 || k > 0) ? 1u : 0u);
          ptx::umma_commit_raw(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      ptx::umma_commit(done_bar);
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    if (kb1 > kb0) {
      ptx::mbar_wait(done_bar, 0);
      ptx::tc_fence_after();
    }
    const size_t Mp = (size_t)p.m_tiles * 128, Np = (size_t)p.n_tiles * BN;
    float *dst = p.partial + (((size_t)slice * taps + tap) * Mp + (size_t)mt * 128 + m) * Np + (size_t)nt * BN;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      if (kb1 > kb0) {
        ptx::tmem_ld_32x32(trow + c, r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4 *>(dst + c + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// kinds of parameter tensors the reduction writes (MXNet layouts)
enum { WG_CONV = 0, WG_CONV1_S2D = 1, WG_DECONV = 2, WG_CONV1_ROW = 3 };

// grad[dst] = sum_slices partial[slice][tap][m][n]; one thread per SOURCE element (tap, m, n) with n fastest, so the
// partial tiles are read coalesced; the (Cout,Cin,kh,kw) destination is written with a k*k-element stride.
//   WG_CONV      : m = co, n = ci            -> (Cout, Cin, k, k)
//   WG_CONV1_S2D : m = co, n = ph*16+pw*8+c  -> (64, 8, 7, 7), kh = 2dh+ph, kw = 2dw+pw      (taps = 16)
//   WG_CONV1_ROW : m = dw*32+ph*16+pw*8+c, n = co, tap = dh -> (64, 8, 7, 7)                  (conv1_wgrad_kernel, taps = 4)
//   WG_DECONV    : m = ci, n = co            -> (Cin, Cout, k, k)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, int kslices, int taps,
                                                           int Mp, int Np, int kind, int D0, int D1, int k,
                                                           float *__restrict__ grad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)taps * Mp * Np) return;
  const int n = (int)(idx % Np), m = (int)((idx / Np) % Mp), tap = (int)(idx / ((size_t)Np * Mp));
  int d0, d1, kh, kw;
  if (kind == WG_CONV1_S2D) {
    d0 = m; d1 = n & 7;
    kh = 2 * (tap >> 2) + ((n >> 4) & 1); kw = 2 * (tap & 3) + ((n >> 3) & 1);
    if (n >= 32) return;
  } else if (kind == WG_CONV1_ROW) {
    d0 = n; d1 = m & 7;
    kh = 2 * tap + ((m >> 4) & 1); kw = 2 * (m >> 5) + ((m >> 3) & 1);
  } else {
    d0 = m; d1 = n; kh = tap / k; kw = tap - kh * k;
  }
  if (d0 >= D0 || d1 >= D1 || kh >= k || kw >= k) return;
  float acc = 0.f;
  for (int s = 0; s < kslices; ++s) acc += partial[(size_t)s * taps * Mp * Np + idx];
  grad[(((size_t)d0 * D1 + d1) * k + kh) * k + kw] = acc;
}

// ---------------------------------------------------------------------------------------------
// conv1 weight gradient (64 x 8 x 7 x 7 from 240 x 320 output pixels): the generic kernel would run 16 taps of
// 128(64 used) x 32 tiles and fetch dZ 16 times.  Here one GEMM per filter ROW dh: M = (dw, 32 space-to-depth channels) =
// 128 rows assembled from FOUR column-shifted TMA boxes of the NHWC-32 input copy (64-byte rows, SWIZZLE_64B, MN-major
// groups LBO = 4 KB apart), N = 64 output channels of dZ (one SWIZZLE_128B box), K = 64 pixels: dZ is fetched 4x instead of
// 16x and no MMA row is padding.
template <int STAGES>
struct Conv1WgradSmem {
  static constexpr int A_BYTES = 4 * 4096, B_BYTES = 8192, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int STAGES>
__global__ void __launch_bounds__(192) conv1_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using S = Conv1WgradSmem<STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * S::STAGE_BYTES);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *done_bar = empty_bar + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dh = blockIdx.x % 4, slice = blockIdx.x / 4;
  const int kb0 = slice * p.kb_per_slice;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_slice);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&p.z_map);
    ptx::prefetch_tmap(&p.a_map[0]);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 64);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    {  // whole warp; the elected lane issues (conv_igemm.cuh ptx::elect_one)
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        const int rx = kb % p.rects_x;
        const int t = kb / p.rects_x;
        const int ry = t % p.rects_y, b = t / p.rects_y;
        const int oy0 = ry * p.BH, ox0 = rx * p.BW;
        uint8_t *st = smem + s * S::STAGE_BYTES;
        ptx::mbar_expect_tx(&full_bar[s], (uint32_t)S::STAGE_BYTES);
#pragma unroll
        for (int dw = 0; dw < 4; ++dw) tma_load_4d(st + dw * 4096, &p.a_map[0], &full_bar[s], 0, ox0 + dw, oy0 + dh, b);
        tma_load_4d(st + S::A_BYTES, &p.z_map, &full_bar[s], 0, ox0 + p.z_off_c, oy0 + p.z_off_r, b);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {  // whole warp; the elected lane issues
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b0 = a0 + S::A_BYTES;
        if (ptx::elect_one()) {
          const uint64_t da0 = ptx::umma_desc_mn(a0, 4096, 512, 4u), db0 = ptx::umma_desc_mn(b0, 8192, 1024, 2u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16_raw(tmem_base, da0 + (uint64_t)(k * 64), db0 + (uint64_t)(k * 128), p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          ptx::umma_commit_raw(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      ptx::umma_commit(done_bar);
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    if (kb1 > kb0) {
      ptx::mbar_wait(done_bar, 0);
      ptx::tc_fence_after();
    }
    float *dst = p.partial + (((size_t)slice * 4 + dh) * 128 + m) * 64;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t r[32];
      if (kb1 > kb0) {
        ptx::tmem_ld_32x32(trow + c, r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4 *>(dst + c + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace dim
