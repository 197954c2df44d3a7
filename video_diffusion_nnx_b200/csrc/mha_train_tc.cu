// Temporal attention forward of the TRAINING engines at C = 32 with its q|k|v projection on tcgen05
// (modules.py:261-323 in the 'b f h w c -> b (h w) f c' arrangement of unet3d.py:86-96; same contract as
// vdn_mha_temporal_fused_fwd: x -> q|k|v [P][768] (kept for the backward), o [P][256], lse [P][8]).
//
// The register-resident warp-MMA kernel (mha_temporal_mma_fwd_kernel<true>, mha_mma.cu) spends 24 of its 32 mma.sync per
// (pixel, head) on the projection and runs at the legacy-MMA issue rate (122 us at 163840 token rows). Here the
// projection is a tcgen05 GEMM and only the F x F core stays on the warp-level path:
//   tile = 8 pixels x 16 token rows (row = px * 16 + f, rows f >= F stay zero) = one UMMA M tile; per tile four passes
//   of two heads: [128 x 192] = X W_pass^T (q | k | v of both heads) into TMEM -> registers -> + bias -> bf16 -> (a) the
//   q|k|v tensor in global memory (64-byte pieces per row), (b) six [128 x 32] blocks in shared memory -> per (pixel,
//   head) S = q k^T, softmax, O = P v as 8 mma.sync with ldmatrix operands (each warp runs the two heads of one pixel
//   interleaved) -> o, lse to global memory.
// Two CTAs per SM; bounds: the q|k|v write (252 MB at 163840 rows) and the TMEM read (384 KB per tile at 64 B/clk).
#include <algorithm>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {
namespace {

constexpr int kTtThreads = 256;
constexpr int kTtXBytes = 128 * 64;          // x tile: 128 rows x 32 ch bf16, SW64
constexpr int kTtWBytes = 768 * 64;          // head-major q|k|v weights [768][32], SW64
constexpr int kTtBlk = 128 * 64;             // one staged [128 rows][32 features] block (64-byte rows, SW64 pattern)
constexpr int kTtSBytes = 6 * kTtBlk;        // q, k, v of the two heads of a pass
constexpr int kTtSmem = 1024 + kTtXBytes + kTtWBytes + kTtSBytes;
constexpr int kTtPassCols = 192;

struct TtMaps {
  CUtensorMap x, w;
};

__device__ __forceinline__ void tt_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tt_ldsm(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void tt_stsm(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}
__device__ __forceinline__ void tt_ldsm_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__global__ void __launch_bounds__(kTtThreads, 2) mha_train_tc_kernel(const __grid_constant__ TtMaps maps,
                                                                    const float* __restrict__ bias_hm,
                                                                    bf16* __restrict__ o, bf16* __restrict__ qkv,
                                                                    float* __restrict__ lse, int B, int F, int HW,
                                                                    int tiles_per_img, int n_tiles, long long* trace) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_bar, w_bar, mma_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[768];

  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* sX = smem;
  uint8_t* sW = sX + kTtXBytes;
  uint8_t* sS = sW + kTtWBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int PXT = 8;

  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w);
    mbar_init(&x_bar, 1);
    mbar_init(&w_bar, 1);
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 256u);
    tmem_relinquish();
  }
  // rows f >= F of the x tile are never written by TMA: zero once (they only feed discarded / masked entries)
  for (int i = threadIdx.x; i < kTtXBytes / 16; i += kTtThreads) reinterpret_cast<uint4*>(sX)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  for (int i = threadIdx.x; i < 768; i += kTtThreads) s_bias[i] = bias_hm ? bias_hm[i] : 0.f;

  int tile = blockIdx.x;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&w_bar, (uint32_t)kTtWBytes);
    for (int i = 0; i < 3; ++i) tma_load_2d(sW + i * 256 * 64, &maps.w, &w_bar, 0, i * 256);
    if (tile < n_tiles) {
      const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * PXT;
      mbar_expect_tx(&x_bar, (uint32_t)(PXT * F * 64));
      for (int px = 0; px < PXT; ++px) tma_load_4d(sX + px * 1024, &maps.x, &x_bar, 0, 0, p0 + px, b);
    }
  }
  mbar_wait(&w_bar, 0);
  __syncthreads();  // s_bias visible

  const uint32_t desc_hi64 = ((8u * 64u) >> 4) | (1u << 14) | (umma_layout_type(64) << 29);
  const uint32_t sx16 = smem_u32(sX) >> 4, sw16 = smem_u32(sW) >> 4;
  const uint32_t sS_u = smem_u32(sS);
  uint32_t x_ph = 0u, mma_ph = 0u;
  int tile_i = 0;
  const int g = lane >> 2, t = lane & 3;
  const float scale = rsqrtf(32.f);
  // per-lane ldmatrix offsets inside a staged [128][32] block (a pixel = 16 rows)
  uint32_t a_off0, a_off1, b_off0, b_off1, t_off0, t_off1;
  bool cv[2][2];
  {
    const uint32_t row16 = (uint32_t)((lane & 7) + 8 * ((lane >> 3) & 1));
    const uint32_t sw16r = (row16 >> 1) & 3u;
    a_off0 = row16 * 64u + ((((uint32_t)(lane >> 4)) ^ sw16r) << 4);        // A operand / transposed B: k step 0 | dims 0-15
    a_off1 = row16 * 64u + (((2u + (uint32_t)(lane >> 4)) ^ sw16r) << 4);   //                            k step 1 | dims 16-31
    t_off0 = a_off0;
    t_off1 = a_off1;
    const uint32_t row8 = (uint32_t)(lane & 7), c4 = (uint32_t)(lane >> 3);
    b_off0 = row8 * 64u + ((c4 ^ ((row8 >> 1) & 3u)) << 4);                 // plain B operand: keys 0-7
    b_off1 = (row8 + 8u) * 64u + ((c4 ^ (((row8 + 8u) >> 1) & 3u)) << 4);   //                  keys 8-15
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i) cv[nt][i] = 8 * nt + 2 * t + i < F;
  }

  for (; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, p0 = (tile - b * tiles_per_img) * PXT;
    const int n_px = min(PXT, HW - p0);
    long long* tr = (trace && blockIdx.x == 0 && threadIdx.x == 0 && tile_i < 60) ? trace + 16 * tile_i : nullptr;
    ++tile_i;
    if (tr) tr[0] = clock64();
    mbar_wait(&x_bar, x_ph);
    x_ph ^= 1u;
    if (tr) tr[1] = clock64();
    // this thread's token row in the TMEM phases
    const int quarter = warp & 3, hl_c = warp >> 2;
    const int r = quarter * 32 + lane;
    auto issue_gemm = [&](int pass) {
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, kTtPassCols, 0, 0);
        const uint32_t b16 = sw16 + (uint32_t)((pass * kTtPassCols * 64) >> 4);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi64) << 32) | ((sx16 + 2u * k) | (1u << 16)),
                    (static_cast<uint64_t>(desc_hi64) << 32) | ((b16 + 2u * k) | (1u << 16)), idesc, k);
        tc_commit(&mma_bar);
      }
    };
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
      // ---- q | k | v of heads 2*pass, 2*pass + 1: [128 x 192] = X W_pass^T (issued one pass ahead, see below) ----
      if (pass == 0) issue_gemm(0);
      mbar_wait(&mma_bar, mma_ph);
      mma_ph ^= 1u;
      tc_fence_after();
      if (tr) tr[2 + 3 * pass] = clock64();
      // ---- TMEM -> + bias -> bf16 -> global q|k|v and the staged blocks (warps 0-3: first head, 4-7: second) ----
      {
        const int h = 2 * pass + hl_c;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(hl_c * 96);
        auto stage_part = [&](const uint32_t (&raw)[32], int part) {
          uint8_t* rowp = sS + (hl_c * 3 + part) * kTtBlk + r * 64;
          const float* bb = &s_bias[h * 96 + part * 32];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 u0 = *reinterpret_cast<const float4*>(bb + c * 8);
            const float4 u1 = *reinterpret_cast<const float4*>(bb + c * 8 + 4);
            uint4 q;
            q.x = pack_bf16x2(__uint_as_float(raw[8 * c + 0]) + u0.x, __uint_as_float(raw[8 * c + 1]) + u0.y);
            q.y = pack_bf16x2(__uint_as_float(raw[8 * c + 2]) + u0.z, __uint_as_float(raw[8 * c + 3]) + u0.w);
            q.z = pack_bf16x2(__uint_as_float(raw[8 * c + 4]) + u1.x, __uint_as_float(raw[8 * c + 5]) + u1.y);
            q.w = pack_bf16x2(__uint_as_float(raw[8 * c + 6]) + u1.z, __uint_as_float(raw[8 * c + 7]) + u1.w);
            *reinterpret_cast<uint4*>(rowp + (((uint32_t)c ^ (uint32_t)((r >> 1) & 3)) << 4)) = q;
          }
        };
        uint32_t raw_a[32], raw_b[32];
        tmem_ld_32x32(taddr, raw_a);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 32u, raw_b);
        stage_part(raw_a, 0);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + 64u, raw_a);
        stage_part(raw_b, 1);
        tmem_ld_wait();
        stage_part(raw_a, 2);
      }
      tc_fence_before();
      __syncthreads();
      // the accumulator has been drained: the next pass's GEMM runs behind this pass's stores and core; after the last
      // pass the x tile has fed its last GEMM and the next tile's can be fetched
      if (pass < 3) {
        issue_gemm(pass + 1);
      } else if (threadIdx.x == 0) {
        const int nt_ = tile + gridDim.x;
        if (nt_ < n_tiles) {
          const int nb = nt_ / tiles_per_img, np0 = (nt_ - nb * tiles_per_img) * PXT;
          mbar_expect_tx(&x_bar, (uint32_t)(PXT * F * 64));
          for (int px = 0; px < PXT; ++px) tma_load_4d(sX + px * 1024, &maps.x, &x_bar, 0, 0, np0 + px, nb);
        }
      }
      // ---- q | k | v of this pass to global memory, coalesced: four lanes per 64-byte (row, head, part) piece. (A thread
      //      storing its own row straight from registers touches 32 cache lines per warp instruction: 2500 cycles.) ----
      {
        const int sub = threadIdx.x & 3, row0 = threadIdx.x >> 2;  // 16-byte chunk, row within a group of 64 rows
#pragma unroll
        for (int blk = 0; blk < 6; ++blk) {
          const int hl = blk / 3, part = blk - 3 * hl;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int rr = row0 + 64 * half, ppx = rr >> 4, ff = rr & 15;
            if (ff < F && ppx < n_px) {
              const uint4 q = *reinterpret_cast<const uint4*>(sS + blk * kTtBlk + rr * 64 + (((uint32_t)sub ^ (uint32_t)((rr >> 1) & 3)) << 4));
              const long grow = ((long)b * F + ff) * HW + p0 + ppx;
              reinterpret_cast<uint4*>(qkv + grow * 768 + part * 256 + (2 * pass + hl) * 32)[sub] = q;
            }
          }
        }
      }
      __syncthreads();  // the q rows of a pixel are about to be replaced by its o rows
      if (tr) tr[3 + 3 * pass] = clock64();
      // ---- core: warp w = pixel w, its two heads interleaved ----
      if (warp < n_px) {
        struct Unit {
          uint32_t qa[2][4], pa[4];
          float S[2][4], O[4][4], m_lo, m_hi, l_lo, l_hi;
        };
        const uint32_t po = (uint32_t)warp * (16u * 64u);
        auto scores = [&](Unit& u, int hl) {
          const uint32_t qb = sS_u + (uint32_t)((hl * 3 + 0) * kTtBlk) + po, kb = sS_u + (uint32_t)((hl * 3 + 1) * kTtBlk) + po;
          tt_ldsm(u.qa[0], qb + a_off0);
          tt_ldsm(u.qa[1], qb + a_off1);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            uint32_t xb[4];
            tt_ldsm(xb, kb + (nt ? b_off1 : b_off0));
            u.S[nt][0] = u.S[nt][1] = u.S[nt][2] = u.S[nt][3] = 0.f;
            tt_mma(u.S[nt], u.qa[0], xb[0], xb[1]);
            tt_mma(u.S[nt], u.qa[1], xb[2], xb[3]);
          }
        };
        auto softmax = [&](Unit& u) {
          float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              u.S[nt][i] = cv[nt][i] ? u.S[nt][i] * scale : -INFINITY;
              u.S[nt][2 + i] = cv[nt][i] ? u.S[nt][2 + i] * scale : -INFINITY;
              m_lo = fmaxf(m_lo, u.S[nt][i]);
              m_hi = fmaxf(m_hi, u.S[nt][2 + i]);
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
              u.S[nt][i] = __expf(u.S[nt][i] - m_lo);
              u.S[nt][2 + i] = __expf(u.S[nt][2 + i] - m_hi);
              l_lo += u.S[nt][i];
              l_hi += u.S[nt][2 + i];
            }
          l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
          l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
          l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
          l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
          u.m_lo = m_lo; u.m_hi = m_hi; u.l_lo = l_lo; u.l_hi = l_hi;
          u.pa[0] = pack_bf16x2(u.S[0][0], u.S[0][1]);
          u.pa[1] = pack_bf16x2(u.S[0][2], u.S[0][3]);
          u.pa[2] = pack_bf16x2(u.S[1][0], u.S[1][1]);
          u.pa[3] = pack_bf16x2(u.S[1][2], u.S[1][3]);
        };
        auto mix = [&](Unit& u, int hl) {  // O = P v
          const uint32_t vb = sS_u + (uint32_t)((hl * 3 + 2) * kTtBlk) + po;
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t xb[4];
            tt_ldsm_t(xb, vb + (np ? t_off1 : t_off0));
            u.O[2 * np][0] = u.O[2 * np][1] = u.O[2 * np][2] = u.O[2 * np][3] = 0.f;
            u.O[2 * np + 1][0] = u.O[2 * np + 1][1] = u.O[2 * np + 1][2] = u.O[2 * np + 1][3] = 0.f;
            tt_mma(u.O[2 * np], u.pa, xb[0], xb[1]);
            tt_mma(u.O[2 * np + 1], u.pa, xb[2], xb[3]);
          }
        };
        auto put = [&](const Unit& u, int hl) {  // o (bf16) over the pixel's q rows; lse straight to global memory
          const int h = 2 * pass + hl;
          const float inv_lo = __fdividef(1.f, u.l_lo), inv_hi = __fdividef(1.f, u.l_hi);
          const uint32_t qb = sS_u + (uint32_t)((hl * 3 + 0) * kTtBlk) + po;
#pragma unroll
          for (int np = 0; np < 2; ++np)
            tt_stsm(qb + (np ? a_off1 : a_off0), pack_bf16x2(u.O[2 * np][0] * inv_lo, u.O[2 * np][1] * inv_lo),
                    pack_bf16x2(u.O[2 * np][2] * inv_hi, u.O[2 * np][3] * inv_hi),
                    pack_bf16x2(u.O[2 * np + 1][0] * inv_lo, u.O[2 * np + 1][1] * inv_lo),
                    pack_bf16x2(u.O[2 * np + 1][2] * inv_hi, u.O[2 * np + 1][3] * inv_hi));
          if (t == 0) {
            if (g < F) lse[(((long)b * F + g) * HW + p0 + warp) * 8 + h] = u.m_lo + __logf(u.l_lo);
            if (g + 8 < F) lse[(((long)b * F + g + 8) * HW + p0 + warp) * 8 + h] = u.m_hi + __logf(u.l_hi);
          }
        };
        Unit ua, ub;
        scores(ua, 0);
        scores(ub, 1);
        softmax(ua);
        softmax(ub);
        mix(ua, 0);
        mix(ub, 1);
        put(ua, 0);
        put(ub, 1);
      }
      __syncthreads();
      {  // o of the two heads to global memory, coalesced like q | k | v
        const int sub = threadIdx.x & 3, row0 = threadIdx.x >> 2;
#pragma unroll
        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int rr = row0 + 64 * half, ppx = rr >> 4, ff = rr & 15;
            if (ff < F && ppx < n_px) {
              const uint4 q = *reinterpret_cast<const uint4*>(sS + (hl * 3) * kTtBlk + rr * 64 + (((uint32_t)sub ^ (uint32_t)((rr >> 1) & 3)) << 4));
              const long grow = ((long)b * F + ff) * HW + p0 + ppx;
              reinterpret_cast<uint4*>(o + grow * 256 + (2 * pass + hl) * 32)[sub] = q;
            }
          }
      }
      __syncthreads();  // staged blocks and the TMEM columns are free for the next pass
      tc_fence_after();
      if (tr) tr[4 + 3 * pass] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
  }
}

}  // namespace

bool mha_train_tc_applicable(int F, int HW) { return F >= 1 && F <= 16 && HW >= 8; }

int mha_train_tc_launch(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv, float* lse, int B, int F,
                        int H, int W, cudaStream_t st) {
  const int HW = H * W;
  TtMaps maps;
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
    const uint64_t dims[2] = {32u, 768u};
    const uint64_t str[1] = {64u};
    const uint32_t box[2] = {32u, 256u};
    if ((rc = encode_tmap_bf16(&maps.w, w_hm, 2, dims, str, box, 64))) return rc;
  }
  const int tiles_per_img = (HW + 7) / 8;
  const int n_tiles = B * tiles_per_img;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(mha_train_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTtSmem);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "mha_train_tc cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const int grid = std::min(n_tiles, 2 * num_sms());
  cudaError_t le = launch_pdl(mha_train_tc_kernel, dim3(grid), dim3(kTtThreads), (size_t)kTtSmem, st, 1, maps, bias_hm,
                              reinterpret_cast<bf16*>(o), reinterpret_cast<bf16*>(qkv), lse, B, F, HW, tiles_per_img, n_tiles,
                              debug_trace_ptr());
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_train_tc launch: %s", cudaGetErrorString(le));
  return check_launch("mha_train_tc");
}

}  // namespace vdn
