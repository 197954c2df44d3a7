// Folded temporal attention block of the inference engines (C = 32), shared-weight GEMMs on tcgen05:
//   out = x + sum_h softmax_j((x A_h + u_h) x^T)_h x M_h + b'        (algebra: mha_temporal_folded_fwd_kernel, mha_mma.cu;
//   reference: modules.py:285-326 + the residual of unet3d.py:86-96, sequences = the F frames of one pixel)
//
// The warp-MMA version spends 16 of its 24 mma.sync per (pixel, head) on the two products with weights that are the
// same for every pixel (y = x A_h: 8, Z M_h: 8) and runs at the legacy-MMA issue rate (~27 cycles per m16n8k16 per SM
// sub-partition, 287 us at 16 samples). Here a CTA owns tiles of 8 pixels x 16 token rows (row = px * 16 + f; rows
// f >= F stay zero), 128 rows = one UMMA M tile:
//   1. TMA: the F frames of each pixel (box 32 ch x F of the (C, F, HW, B) view) -> 16-row slots of a K-major SW64 tile
//   2. tcgen05: Y[128 x 256] = X A (all 8 heads: two M128 x N256 x K16 instructions), accumulator in TMEM
//   3. TMEM -> registers -> (+u) bf16 -> smem in the K-major SW128 chunk layout GEMM 2 will read
//   4. per (pixel, head), one warp per head: S = y x^T, softmax, Z = P x as 8 mma.sync with ldmatrix-fed operands
//      (y rows, x rows plain and transposed); Z overwrites y in place through stmatrix (same rows, same 64 bytes);
//      every ldmatrix / stmatrix address is a per-lane constant + px * 16 rows
//   5. tcgen05: O[128 x 32] = Z M (K = 256: 16 instructions M128 x N32 x K16) into the same TMEM columns
//   6. TMEM -> + b' + x -> bf16 -> global
// Two CTAs per SM (105 KB of shared memory, 256 TMEM columns each) overlap each other's phases; the next tile's x
// load is in flight during phases 5-6.
#include <algorithm>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {
namespace {

constexpr int kFtThreads = 256;
constexpr int kXBytes = 128 * 64;          // x tile: 128 rows x 32 ch bf16, SW64
constexpr int kWBytes = 256 * 64;          // A (and M) stacked over heads: 256 rows x 32 bf16, SW64
constexpr int kYZBytes = 4 * 128 * 128;    // y / Z: 4 K-chunks (2 heads each) of 128 rows x 128 B, SW128
constexpr int kFtSmem = 1024 + kXBytes + 2 * kWBytes + kYZBytes;

struct FtMaps {
  CUtensorMap x, a, m;
};

__device__ __forceinline__ void ft_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ft_ldsm(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ft_stsm(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}
__device__ __forceinline__ void ft_ldsm_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__global__ void __launch_bounds__(kFtThreads, 2) mha_folded_tc_kernel(const __grid_constant__ FtMaps maps,
                                                                     const bf16* __restrict__ x, const float* __restrict__ fu,
                                                                     const float* __restrict__ fb, bf16* __restrict__ out,
                                                                     int B, int F, int HW, int PXT, int tiles_per_img,
                                                                     int n_tiles, long long* trace) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_bar, w_bar, mma_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) uint4 s_zero;
  __shared__ __align__(16) float s_fu[256];
  __shared__ __align__(16) float s_fb[32];

  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* sX = smem;
  uint8_t* sA = sX + kXBytes;
  uint8_t* sM = sA + kWBytes;
  uint8_t* sYZ = sM + kWBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.m);
    mbar_init(&x_bar, 1);
    mbar_init(&w_bar, 1);
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
    s_zero = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 256u);
    tmem_relinquish();
  }
  // rows R..127 of the x tile are never written by TMA: keep them finite (they only feed discarded accumulator rows)
  for (int i = threadIdx.x; i < kXBytes / 16; i += kFtThreads) reinterpret_cast<uint4*>(sX)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < kYZBytes / 16; i += kFtThreads) reinterpret_cast<uint4*>(sYZ)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  for (int i = threadIdx.x; i < 256; i += kFtThreads) s_fu[i] = fu[i];
  if (threadIdx.x < 32) s_fb[threadIdx.x] = fb[threadIdx.x];

  int tile = blockIdx.x;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&w_bar, 2u * kWBytes);
    tma_load_2d(sA, &maps.a, &w_bar, 0, 0);
    tma_load_2d(sM, &maps.m, &w_bar, 0, 0);
    if (tile < n_tiles) {
      const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * PXT;
      mbar_expect_tx(&x_bar, (uint32_t)(PXT * F * 64));
      for (int px = 0; px < PXT; ++px) tma_load_4d(sX + px * 1024, &maps.x, &x_bar, 0, 0, p0 + px, b);
    }
  }
  mbar_wait(&w_bar, 0);
  __syncthreads();  // s_fu / s_fb visible

  const uint32_t desc_hi64 = ((8u * 64u) >> 4) | (1u << 14) | (umma_layout_type(64) << 29);
  const uint32_t desc_hi128 = ((8u * 128u) >> 4) | (1u << 14) | (umma_layout_type(128) << 29);
  const uint32_t sx16 = smem_u32(sX) >> 4, sa16 = smem_u32(sA) >> 4, sm16 = smem_u32(sM) >> 4, syz16 = smem_u32(sYZ) >> 4;
  const uint32_t sX_u = smem_u32(sX), sYZ_u = smem_u32(sYZ), zero_u = smem_u32(&s_zero);
  uint32_t x_ph = 0u, mma_ph = 0u;
  int tile_i = 0;
  const int t = lane & 3;
  const int h = warp;  // phase 4: one warp per head
  // per-lane ldmatrix / stmatrix addresses for pixel 0 of a tile (a pixel = 16 rows; add px * 16 rows)
  uint32_t y_addr0, y_addr1, xs_addr0, xs_addr1, xt_addr0, xt_addr1;
  bool cv[2][2];
  {
    const uint32_t row16 = (uint32_t)((lane & 7) + 8 * ((lane >> 3) & 1));  // A operand / transposed B / stmatrix: token
    const uint32_t yz_h = sYZ_u + (uint32_t)((h >> 1) * (128 * 128)) + row16 * 128u;
    const uint32_t hc = (uint32_t)((h & 1) * 4), sw128 = row16 & 7u;
    y_addr0 = yz_h + (((hc + (uint32_t)(lane >> 4)) ^ sw128) << 4);        // channel chunks 0, 1 of the head
    y_addr1 = yz_h + (((hc + 2u + (uint32_t)(lane >> 4)) ^ sw128) << 4);   // channel chunks 2, 3
    const uint32_t row8 = (uint32_t)(lane & 7);                            // plain B operand: key token within the n-tile
    const uint32_t c4 = (uint32_t)(lane >> 3);
    xs_addr0 = sX_u + row8 * 64u + ((c4 ^ ((row8 >> 1) & 3u)) << 4);
    xs_addr1 = sX_u + (row8 + 8u) * 64u + ((c4 ^ (((row8 + 8u) >> 1) & 3u)) << 4);
    const uint32_t sw64 = (row16 >> 1) & 3u;
    xt_addr0 = sX_u + row16 * 64u + ((((uint32_t)(lane >> 4)) ^ sw64) << 4);
    xt_addr1 = sX_u + row16 * 64u + (((2u + (uint32_t)(lane >> 4)) ^ sw64) << 4);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i) cv[nt][i] = 8 * nt + 2 * t + i < F;
  }

  for (; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * PXT;
    const int n_px = min(PXT, HW - p0);
    long long* tr = (trace && blockIdx.x == 0 && threadIdx.x == 0 && tile_i < 120) ? trace + 8 * tile_i : nullptr;
    ++tile_i;
    if (tr) tr[0] = clock64();
    // ---- 1 + 2: x tile landed -> Y = X A on the tensor cores ----
    mbar_wait(&x_bar, x_ph);
    x_ph ^= 1u;
    if (tr) tr[1] = clock64();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi64) << 32) | ((sx16 + 2u * k) | (1u << 16)),
                  (static_cast<uint64_t>(desc_hi64) << 32) | ((sa16 + 2u * k) | (1u << 16)), idesc, k);
      tc_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    if (tr) tr[2] = clock64();
    // ---- 3: y (+ u) -> bf16 -> shared memory, K-major SW128 chunks ----
    {
      const int quarter = warp & 3, half = warp >> 2;
      const int r = quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      auto stage_head = [&](const uint32_t (&raw)[32], int hh) {
        uint8_t* rowp = sYZ + (hh >> 1) * (128 * 128) + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 u0 = *reinterpret_cast<const float4*>(&s_fu[hh * 32 + c * 8]);
          const float4 u1 = *reinterpret_cast<const float4*>(&s_fu[hh * 32 + c * 8 + 4]);
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(raw[8 * c + 0]) + u0.x, __uint_as_float(raw[8 * c + 1]) + u0.y);
          q.y = pack_bf16x2(__uint_as_float(raw[8 * c + 2]) + u0.z, __uint_as_float(raw[8 * c + 3]) + u0.w);
          q.z = pack_bf16x2(__uint_as_float(raw[8 * c + 4]) + u1.x, __uint_as_float(raw[8 * c + 5]) + u1.y);
          q.w = pack_bf16x2(__uint_as_float(raw[8 * c + 6]) + u1.z, __uint_as_float(raw[8 * c + 7]) + u1.w);
          *reinterpret_cast<uint4*>(rowp + (((uint32_t)((hh & 1) * 4 + c) ^ (uint32_t)(r & 7)) << 4)) = q;
        }
      };
      // four heads per warp, the next head's TMEM load in flight while the current one is converted
      uint32_t raw_a[32], raw_b[32];
      const int h0 = half * 4;
      tmem_ld_32x32(taddr + (uint32_t)(h0 * 32), raw_a);
      tmem_ld_wait();
      tmem_ld_32x32(taddr + (uint32_t)((h0 + 1) * 32), raw_b);
      stage_head(raw_a, h0);
      tmem_ld_wait();
      tmem_ld_32x32(taddr + (uint32_t)((h0 + 2) * 32), raw_a);
      stage_head(raw_b, h0 + 1);
      tmem_ld_wait();
      tmem_ld_32x32(taddr + (uint32_t)((h0 + 3) * 32), raw_b);
      stage_head(raw_a, h0 + 2);
      tmem_ld_wait();
      stage_head(raw_b, h0 + 3);
    }
    tc_fence_before();
    __syncthreads();
    if (tr) tr[3] = clock64();
    // ---- 4: per (pixel, head): S = y x^T, softmax, Z = P x (Z replaces y) ----
    // A pixel owns 16 consecutive rows (frames F..15 are zero rows of x), so every ldmatrix / stmatrix address of a
    // lane is a per-lane constant plus px * (16 rows): no predicates, no per-pixel swizzle arithmetic.
    // Two pixels per iteration with their instruction streams interleaved in source order (the asm statements are
    // volatile, so the compiler keeps this order): one chain's ldmatrix -> mma -> shuffle latencies hide behind the other's.
    {
      struct Px {
        uint32_t ya[2][4], pa[4];
        float S[2][4], Z[4][4], inv_lo, inv_hi;
      };
      auto scores = [&](Px& p, uint32_t yo, uint32_t xo) {
        ft_ldsm(p.ya[0], y_addr0 + yo);
        ft_ldsm(p.ya[1], y_addr1 + yo);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          uint32_t xb[4];
          ft_ldsm(xb, (nt ? xs_addr1 : xs_addr0) + xo);
          p.S[nt][0] = p.S[nt][1] = p.S[nt][2] = p.S[nt][3] = 0.f;
          ft_mma(p.S[nt], p.ya[0], xb[0], xb[1]);
          ft_mma(p.S[nt], p.ya[1], xb[2], xb[3]);
        }
      };
      auto softmax = [&](Px& p) {
        float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            p.S[nt][i] = cv[nt][i] ? p.S[nt][i] : -INFINITY;
            p.S[nt][2 + i] = cv[nt][i] ? p.S[nt][2 + i] : -INFINITY;
            m_lo = fmaxf(m_lo, p.S[nt][i]);
            m_hi = fmaxf(m_hi, p.S[nt][2 + i]);
          }
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
        m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
        m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
        float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            p.S[nt][i] = __expf(p.S[nt][i] - m_lo);
            p.S[nt][2 + i] = __expf(p.S[nt][2 + i] - m_hi);
            l_lo += p.S[nt][i];
            l_hi += p.S[nt][2 + i];
          }
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
        p.inv_lo = __fdividef(1.f, l_lo);
        p.inv_hi = __fdividef(1.f, l_hi);
        p.pa[0] = pack_bf16x2(p.S[0][0], p.S[0][1]);
        p.pa[1] = pack_bf16x2(p.S[0][2], p.S[0][3]);
        p.pa[2] = pack_bf16x2(p.S[1][0], p.S[1][1]);
        p.pa[3] = pack_bf16x2(p.S[1][2], p.S[1][3]);
      };
      auto mix = [&](Px& p, uint32_t xo) {  // Z = P x
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t xb[4];
          ft_ldsm_t(xb, (np ? xt_addr1 : xt_addr0) + xo);
          p.Z[2 * np][0] = p.Z[2 * np][1] = p.Z[2 * np][2] = p.Z[2 * np][3] = 0.f;
          p.Z[2 * np + 1][0] = p.Z[2 * np + 1][1] = p.Z[2 * np + 1][2] = p.Z[2 * np + 1][3] = 0.f;
          ft_mma(p.Z[2 * np], p.pa, xb[0], xb[1]);
          ft_mma(p.Z[2 * np + 1], p.pa, xb[2], xb[3]);
        }
      };
      auto put = [&](const Px& p, uint32_t yo) {  // Z (bf16) over y: (rows 0-7 | 8-15) x (channel chunks 0, 1 | 2, 3)
#pragma unroll
        for (int np = 0; np < 2; ++np)
          ft_stsm((np ? y_addr1 : y_addr0) + yo, pack_bf16x2(p.Z[2 * np][0] * p.inv_lo, p.Z[2 * np][1] * p.inv_lo),
                  pack_bf16x2(p.Z[2 * np][2] * p.inv_hi, p.Z[2 * np][3] * p.inv_hi),
                  pack_bf16x2(p.Z[2 * np + 1][0] * p.inv_lo, p.Z[2 * np + 1][1] * p.inv_lo),
                  pack_bf16x2(p.Z[2 * np + 1][2] * p.inv_hi, p.Z[2 * np + 1][3] * p.inv_hi));
      };
      int px = 0;
      for (; px + 1 < n_px; px += 2) {
        const uint32_t yo = (uint32_t)px * (16u * 128u), xo = (uint32_t)px * (16u * 64u);
        Px a, c;
        scores(a, yo, xo);
        scores(c, yo + 16u * 128u, xo + 16u * 64u);
        softmax(a);
        softmax(c);
        mix(a, xo);
        mix(c, xo + 16u * 64u);
        put(a, yo);
        put(c, yo + 16u * 128u);
      }
      if (px < n_px) {
        const uint32_t yo = (uint32_t)px * (16u * 128u), xo = (uint32_t)px * (16u * 64u);
        Px a;
        scores(a, yo, xo);
        softmax(a);
        mix(a, xo);
        put(a, yo);
      }
    }
    fence_proxy_async_smem();  // Z was written through the generic proxy, the tensor cores read it through the async one
    __syncthreads();
    if (tr) tr[4] = clock64();
    // ---- 5: O = Z M on the tensor cores; the next tile's x load goes out now (sX is no longer read) ----
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
#pragma unroll
      for (int hh = 0; hh < 8; ++hh) {
        const uint32_t a0 = syz16 + (uint32_t)(((hh >> 1) * (128 * 128) + (hh & 1) * 64) >> 4);
        const uint32_t b0 = sm16 + (uint32_t)((hh * 32 * 64) >> 4);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi128) << 32) | ((a0 + 2u * k) | (1u << 16)),
                    (static_cast<uint64_t>(desc_hi64) << 32) | ((b0 + 2u * k) | (1u << 16)), idesc, (hh | k) != 0 ? 1u : 0u);
      }
      tc_commit(&mma_bar);
      const int nt_ = tile + gridDim.x;
      if (nt_ < n_tiles) {
        const int nb = nt_ / tiles_per_img, np0 = (nt_ - nb * tiles_per_img) * PXT;
        mbar_expect_tx(&x_bar, (uint32_t)(PXT * F * 64));
        for (int px = 0; px < PXT; ++px) tma_load_4d(sX + px * 1024, &maps.x, &x_bar, 0, 0, np0 + px, nb);
      }
    }
    // the residual rows of x (L2 hits) are requested before the wait for O: warps 0-7 = 4 lane quarters x 2 column halves
    const int o_r = (warp & 3) * 32 + lane, o_px = o_r >> 4, o_f = o_r & 15, o_half = warp >> 2;
    const bool o_valid = o_f < F && o_px < n_px;
    const long o_row = ((long)b * F + o_f) * HW + p0 + o_px;
    uint4 xres[2];
    if (o_valid) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + o_row * 32) + 2 * o_half;
      xres[0] = __ldg(xr);
      xres[1] = __ldg(xr + 1);
    }
    mbar_wait(&mma_bar, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    if (tr) tr[5] = clock64();
    // ---- 6: + b' + x (residual, re-read from L2) -> bf16 -> global ----
    {
      uint32_t raw[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(16 * o_half), raw);
      tmem_ld_wait();
      if (o_valid) {
        uint4* op = reinterpret_cast<uint4*>(out + o_row * 32) + 2 * o_half;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 xv = xres[c];
          const float* bb = &s_fb[16 * o_half + c * 8];
          float2 a;
          uint4 q;
          a = unpack_bf16x2(xv.x); q.x = pack_bf16x2(__uint_as_float(raw[8 * c + 0]) + bb[0] + a.x, __uint_as_float(raw[8 * c + 1]) + bb[1] + a.y);
          a = unpack_bf16x2(xv.y); q.y = pack_bf16x2(__uint_as_float(raw[8 * c + 2]) + bb[2] + a.x, __uint_as_float(raw[8 * c + 3]) + bb[3] + a.y);
          a = unpack_bf16x2(xv.z); q.z = pack_bf16x2(__uint_as_float(raw[8 * c + 4]) + bb[4] + a.x, __uint_as_float(raw[8 * c + 5]) + bb[5] + a.y);
          a = unpack_bf16x2(xv.w); q.w = pack_bf16x2(__uint_as_float(raw[8 * c + 6]) + bb[6] + a.x, __uint_as_float(raw[8 * c + 7]) + bb[7] + a.y);
          op[c] = q;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM columns and the y / Z buffer are free for the next tile
    tc_fence_after();
    if (tr) tr[6] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
  }
}

}  // namespace

bool mha_folded_tc_applicable(int F, int HW) { return F >= 1 && F <= 16 && HW >= 8; }

int mha_folded_tc_launch(const void* x, const void* fa, const float* fu, const void* fm, const float* fb, void* out, int B,
                         int F, int H, int W, cudaStream_t st) {
  const int HW = H * W;
  const int PXT = 8;  // pixels per tile, 16 token rows each
  FtMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  {
    // x viewed as (C, F, HW, B): a box of (32, F, 1, 1) = the F frames of one pixel lands as F rows of 64 bytes
    const uint64_t dims[4] = {32u, (uint64_t)F, (uint64_t)HW, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)HW * 64u, 64u, (uint64_t)F * HW * 64u};
    const uint32_t box[4] = {32u, (uint32_t)F, 1u, 1u};
    if ((rc = encode_tmap_bf16(&maps.x, x, 4, dims, str, box, 64))) return rc;
  }
  {
    const uint64_t dims[2] = {32u, 256u};
    const uint64_t str[1] = {64u};
    const uint32_t box[2] = {32u, 256u};
    if ((rc = encode_tmap_bf16(&maps.a, fa, 2, dims, str, box, 64))) return rc;
    if ((rc = encode_tmap_bf16(&maps.m, fm, 2, dims, str, box, 64))) return rc;
  }
  const int tiles_per_img = (HW + PXT - 1) / PXT;
  const int n_tiles = B * tiles_per_img;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(mha_folded_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmem);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "mha_folded_tc cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const int grid = std::min(n_tiles, 2 * num_sms());
  cudaError_t le = launch_pdl(mha_folded_tc_kernel, dim3(grid), dim3(kFtThreads), (size_t)kFtSmem, st, 1, maps,
                              reinterpret_cast<const bf16*>(x), fu, fb, reinterpret_cast<bf16*>(out), B, F, HW, PXT,
                              tiles_per_img, n_tiles, debug_trace_ptr());
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_folded_tc launch: %s", cudaGetErrorString(le));
  return check_launch("mha_folded_tc");
}

}  // namespace vdn
