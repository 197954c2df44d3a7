// Temporal attention forward entirely on tensor cores (tcgen05): QKV projection, S = Q K^T, O = P V.
// (reference: unet3d.py:86-96,118-120 + modules.py:285-323; out projection + residual stay in tapgemm)
//
// One CTA owns PX adjacent pixels x all F frames (PX*F <= 128 tokens). Per head:
//   1. projection  [128 x C] x [C x 96] -> TMEM (q|k|v of the head), operands by TMA (5-D box = the
//      'b f h w c -> b (h w) f c' rearrangement)
//   2. the worker threads move q,k,v (+bias, bf16) from TMEM into three K-major smem tiles, PERMUTING
//      rows to pixel-major order r' = px*F + f, so that the scores a token needs are F contiguous columns
//   3. S = Q K^T as ONE 128x128x32 MMA over the whole tile: only the F x F diagonal blocks (same pixel)
//      are used - 8x more MACs than needed, still ~30x cheaper than doing them on CUDA cores
//   4. each worker thread pulls the <= 48-column window of its S row out of TMEM, selects its F scores,
//      softmax in fp32, writes F bf16 probabilities into the (pre-zeroed) block-diagonal P tile
//   5. O = P V as a 128x32x128 MMA (V read MN-major from the same smem tile), 1/l applied on the way out
// TMEM: [0,96) qkv, [96,224) S, [224,256) O -> 256 columns, two CTAs per SM.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kTcThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 workers
constexpr int kTcStages = 2;

struct TcMaps {
  CUtensorMap x;  // (C, W, H, F, B) bf16
  CUtensorMap w;  // (C, 768) bf16, head-major rows (h*96 + part*32 + d)
};
struct TcArgs {
  int B, H, W, C, PX, chunks;  // a tile = PX consecutive pixels of the flattened H*W index (all F frames)
  const float* bias;  // [768] head-major
  bf16* o;            // [P][256]
  bf16* qkv;          // [P][768] or null
  float* lse;         // [P][8] or null
};

// byte offset of 16-byte chunk `c` of row `r` in a K-major tile with 64-byte rows (SWIZZLE_64B)
__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
// byte offset of element column `col` (bf16) of row `r` in a K-major [128 x 128] tile made of two
// [128 x 64] SWIZZLE_128B atoms-columns (16 KB each)
__device__ __forceinline__ uint32_t sw128_off(int r, int col) {
  const int atom = col >> 6, cc = col & 63;
  return (uint32_t)(atom * 16384 + r * 128 + ((((cc >> 3) ^ (r & 7)) << 4)) + (cc & 7) * 2);
}

template <int BK, int F>
__global__ void __launch_bounds__(kTcThreads) mha_tc_fwd_kernel(const __grid_constant__ TcMaps maps, const TcArgs a) {
  constexpr int kSw = BK * 2;
  constexpr int kXBytes = 128 * BK * 2;
  constexpr int kWBytes = 96 * BK * 2;
  constexpr int kWBytesPad = (kWBytes + 1023) & ~1023;
  constexpr int kStageBytes = kXBytes + kWBytesPad;
  constexpr uint32_t kLayout = umma_layout_type(kSw);
  constexpr uint32_t kSBO = 8 * kSw;
  constexpr int kTile = 128 * 64;  // bytes of a [128 x 32] bf16 tile

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTcStages], empty_bar[kTcStages];
  __shared__ __align__(8) uint64_t qkv_full, qkv_empty, qk_ready, s_full, p_ready, o_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[768];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* sQ = smem + kTcStages * kStageBytes;
  uint8_t* sK = sQ + kTile;
  uint8_t* sV = sK + kTile;
  uint8_t* sP = sV + kTile;  // 32 KB

  const int PX = a.PX;
  const int rows = PX * F;
  const int HW = a.H * a.W;
  const int tiles_img = (HW + PX - 1) / PX;  // the last tile of a batch element may be partial (TMA zero fill)
  const int tile = blockIdx.x;
  const int b = tile / tiles_img;
  const int p0 = (tile - b * tiles_img) * PX;

  pdl_trigger();
  // zero V (rows >= PX*F must be finite zeros: they meet P's zero columns) and the block-diagonal P tile
  for (int i = threadIdx.x; i < (kTile + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sV)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&qkv_full, 1);
    mbar_init(&qkv_empty, 4);
    mbar_init(&qk_ready, 4);
    mbar_init(&s_full, 1);
    mbar_init(&p_ready, 4);
    mbar_init(&o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 256);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // zero fill visible to the tensor core proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t t_qkv = tmem_base, t_s = tmem_base + 96, t_o = tmem_base + 224;
  pdl_wait();  // smem zero fill / barrier init / TMEM alloc above overlapped the previous kernel
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_bias[i] = a.bias ? a.bias[i] : 0.f;
  __syncthreads();

  if (warp == 0) {
    if (elect_one()) {  // one elected lane (not `lane == 0`): keeps the issue loop on the uniform datapath
      const int n_steps = 8 * a.chunks;
      int h = 0, c = 0;
      for (int it = 0; it < n_steps; ++it) {
        const int st = it % kTcStages;
        const uint32_t ph = (uint32_t)(it / kTcStages) & 1u;
        mbar_wait(&empty_bar[st], ph ^ 1u);
        uint8_t* sx = smem + st * kStageBytes;
        mbar_expect_tx(&full_bar[st], (uint32_t)(rows * BK * 2 + kWBytes));
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(sx)),
            "l"(reinterpret_cast<uint64_t>(&maps.x)), "r"(smem_u32(&full_bar[st])), "r"(c * BK), "r"(p0), "r"(0),
            "r"(0), "r"(b)
            : "memory");
        tma_load_2d(sx + kXBytes, &maps.w, &full_bar[st], c * BK, h * 96);
        if (++c == a.chunks) {
          c = 0;
          ++h;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t id_proj = umma_idesc_bf16(128, 96, 0, 0);
      const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t id_o = umma_idesc_bf16(128, 32, 0, 1);  // B = V read MN-major
      int it = 0;
      auto project = [&](int h) {
        mbar_wait(&qkv_empty, (uint32_t)(h & 1) ^ 1u);
        tc_fence_after();
        for (int c = 0; c < a.chunks; ++c, ++it) {
          const int st = it % kTcStages;
          const uint32_t ph = (uint32_t)(it / kTcStages) & 1u;
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint32_t sx = smem_u32(smem + st * kStageBytes);
          const uint64_t da = umma_smem_desc(sx, 16, kSBO, kLayout);
          const uint64_t db = umma_smem_desc(sx + kXBytes, 16, kSBO, kLayout);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(t_qkv, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), id_proj, (c | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[st]);
        }
        tc_commit(&qkv_full);
      };
      project(0);
      for (int h = 0; h < 8; ++h) {
        const uint32_t par = (uint32_t)(h & 1);
        // S = Q K^T : [128 x 32] x [128 x 32]^T, 64-byte rows (SWIZZLE_64B), two K=16 slices
        mbar_wait(&qk_ready, par);
        tc_fence_after();
        {
          const uint64_t dq = umma_smem_desc(smem_u32(sQ), 16, 512, umma_layout_type(64));
          const uint64_t dk = umma_smem_desc(smem_u32(sK), 16, 512, umma_layout_type(64));
          umma_bf16(t_s, dq, dk, id_s, 0u);
          umma_bf16(t_s, dq + 2, dk + 2, id_s, 1u);
          tc_commit(&s_full);
        }
        if (h + 1 < 8) project(h + 1);
        // O = P V : A = P [128 x 128] K-major (two SWIZZLE_128B atoms), B = V [128 tokens x 32] MN-major
        mbar_wait(&p_ready, par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t pa = smem_u32(sP) + (uint32_t)((k >> 2) * 16384 + (k & 3) * 32);
          const uint32_t vb = smem_u32(sV) + (uint32_t)(k * 16 * 64);
          const uint64_t dp = umma_smem_desc(pa, 16, 1024, umma_layout_type(128));
          const uint64_t dv = umma_smem_desc(vb, (uint32_t)kTile, 512, umma_layout_type(64));
          umma_bf16(t_o, dp, dv, id_o, k != 0 ? 1u : 0u);
        }
        tc_commit(&o_full);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;       // TMEM lane of this thread
    // role A: projection row r is token (f, px) in frame-major order (the TMA box order)
    const bool inA = r < rows;
    const int fA = inA ? r / PX : 0, pxA = inA ? r % PX : 0;
    const bool validA = inA && p0 + pxA < HW;
    const int rpA = pxA * F + fA;            // its row in the pixel-major attention tiles
    const long growA = ((long)b * F + fA) * HW + p0 + pxA;
    // role B: attention row r is token (px, f) in pixel-major order
    const bool inB = r < rows;
    const int pxB = inB ? r / F : 0, fB = inB ? r % F : 0;
    const bool validB = inB && p0 + pxB < HW;
    const long growB = ((long)b * F + fB) * HW + p0 + pxB;
    const int px_lo = (quarter * 32) / F;    // first pixel of this warp's rows
    constexpr int kSlots = (32 % F == 0) ? (32 / F) : ((31 / F) + 2);  // pixels intersecting a warp's 32 rows
    constexpr int kWin = kSlots * F > 32 ? 48 : 32;
    // first S column this warp needs, clamped so the window stays inside the 128-column tile; the clamp
    // shifts the window by a whole number of F-wide slots (F = 10: 90 -> 80), which `sel` absorbs
    const int c0 = min(px_lo * F, 128 - kWin);
    const int sel = pxB - px_lo + (px_lo * F - c0) / F;  // which F-wide slot of the window is mine
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const float scale = rsqrtf(32.f);

    for (int h = 0; h < 8; ++h) {
      const uint32_t par = (uint32_t)(h & 1);
      // ---- 2. q, k, v: TMEM -> (+bias, bf16) -> smem tiles at the permuted row ----
      mbar_wait(&qkv_full, par);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        uint32_t raw[32];
        tmem_ld_32x32(t_qkv + lane_sel + (uint32_t)(part * 32), raw);
        tmem_ld_wait();
        const float* bh = s_bias + h * 96 + part * 32;
        uint32_t packed[16];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bh + e);
          packed[e >> 1] = pack_bf16x2(__uint_as_float(raw[e]) + b4.x, __uint_as_float(raw[e + 1]) + b4.y);
          packed[(e >> 1) + 1] = pack_bf16x2(__uint_as_float(raw[e + 2]) + b4.z, __uint_as_float(raw[e + 3]) + b4.w);
        }
        if (inA) {
          uint8_t* dst = part == 0 ? sQ : (part == 1 ? sK : sV);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(dst + sw64_off(rpA, c)) =
                make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
          if (a.qkv && validA) {
            uint4* gp = reinterpret_cast<uint4*>(a.qkv + growA * 768 + part * 256 + h * 32);
#pragma unroll
            for (int c = 0; c < 4; ++c) gp[c] = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
          }
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&qkv_empty);  // TMEM qkv columns may be overwritten by the next head's projection
        mbar_arrive(&qk_ready);   // q, k (and v) tiles are in smem
      }
      // ---- 4. my S row: window of <= 48 columns, select my F scores, softmax ----
      mbar_wait(&s_full, par);
      tc_fence_after();
      float sc[F];
      {
        uint32_t win[kWin];
        tmem_ld_32x32(t_s + lane_sel + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&win[0]));
        if (kWin == 48) tmem_ld_32x16(t_s + lane_sel + (uint32_t)(c0 + 32), *reinterpret_cast<uint32_t(*)[16]>(&win[32]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < F; ++j) {
          float v = __uint_as_float(win[j]);
#pragma unroll
          for (int s = 1; s < kSlots; ++s)
            if (s * F + j < kWin) v = (sel == s) ? __uint_as_float(win[s * F + j]) : v;
          sc[j] = v * scale;
        }
      }
      float mx = sc[0];
#pragma unroll
      for (int j = 1; j < F; ++j) mx = fmaxf(mx, sc[j]);
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) {
        sc[j] = __expf(sc[j] - mx);
        l += sc[j];
      }
      if (inB) {
#pragma unroll
        for (int j = 0; j < F; j += 2) {
          // columns pxB*F + j, +1 of my P row (F is even for the instantiated kernels -> aligned bf16x2 stores)
          const uint32_t v2 = pack_bf16x2(sc[j], (j + 1 < F) ? sc[j + 1] : 0.f);
          *reinterpret_cast<uint32_t*>(sP + sw128_off(r, pxB * F + j)) = v2;
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready);
      // ---- 5. O row ----
      mbar_wait(&o_full, par);
      tc_fence_after();
      {
        uint32_t raw[32];
        tmem_ld_32x32(t_o + lane_sel, raw);
        tmem_ld_wait();
        if (validB) {
          const float inv = 1.f / l;
          uint4* op = reinterpret_cast<uint4*>(a.o + growB * 256 + h * 32);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(raw[8 * c + 0]) * inv, __uint_as_float(raw[8 * c + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(raw[8 * c + 2]) * inv, __uint_as_float(raw[8 * c + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(raw[8 * c + 4]) * inv, __uint_as_float(raw[8 * c + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(raw[8 * c + 6]) * inv, __uint_as_float(raw[8 * c + 7]) * inv);
            op[c] = u;
          }
          if (a.lse) a.lse[growB * 8 + h] = mx + __logf(l);
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int BK, int F>
static int launch_tc(const TcMaps& maps, const TcArgs& a, int n_tiles, cudaStream_t st) {
  constexpr int stage = 128 * BK * 2 + ((96 * BK * 2 + 1023) & ~1023);
  const int smem = kTcStages * stage + 3 * 8192 + 32768 + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(mha_tc_fwd_kernel<BK, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "mha_tc cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  cudaError_t le = launch_pdl(mha_tc_fwd_kernel<BK, F>, dim3(n_tiles), dim3(kTcThreads), (size_t)smem, st, 1, maps, a);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_tc_fwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_tc_fwd_kernel");
}

}  // namespace vdn

using namespace vdn;

// Returns 1 if (F, C) is served by the tensor-core kernel (otherwise use vdn_mha_temporal_fused_fwd).
extern "C" int vdn_mha_temporal_tc_supported(int F, int C) {
  if (C == 32 && F >= 1 && F <= 16) return 1;  // register-resident warp-MMA kernel (mha_mma.cu)
  return (F == 10 || F == 16) && (C % 32 == 0) ? 1 : 0;
}

// Same contract as vdn_mha_temporal_fused_fwd.
extern "C" int vdn_mha_temporal_tc_fwd(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv,
                                       float* lse, int B, int F, int H, int W, int C, void* stream) {
  VDN_REQUIRE(x && w_hm && o && B > 0 && H > 0 && W > 0, VDN_E_SHAPE, "mha_tc: bad args");
  VDN_REQUIRE(vdn_mha_temporal_tc_supported(F, C), VDN_E_SHAPE, "mha_tc: F=%d C=%d not instantiated", F, C);
  const bool force_tcgen05 = tune_on("VDN_MHA_TC_FWD");  // A/B comparison only
  // training forward (q|k|v and lse kept for the backward): projection on tcgen05, F x F core on warp-level MMAs
  if (C == 32 && qkv && lse && vdn::mha_train_tc_applicable(F, H * W) && !tune_on("VDN_MHA_TRAIN_MMA") && !force_tcgen05)
    return vdn::mha_train_tc_launch(x, w_hm, bias_hm, o, qkv, lse, B, F, H, W, reinterpret_cast<cudaStream_t>(stream));
  if (C == 32 && !(force_tcgen05 && (F == 10 || F == 16)))
    return vdn::mha_temporal_mma_fwd_launch(x, w_hm, bias_hm, o, qkv, lse, B, F, H, W, reinterpret_cast<cudaStream_t>(stream));
  const int PX = std::min(128 / F, H * W);  // pixels per 128-row tile (12 for F = 10, 8 for F = 16)
  const int BK = (C % 64 == 0) ? 64 : 32;
  TcArgs a;
  a.B = B; a.H = H; a.W = W; a.C = C; a.PX = PX; a.chunks = C / BK;
  a.bias = bias_hm;
  a.o = reinterpret_cast<bf16*>(o);
  a.qkv = reinterpret_cast<bf16*>(qkv);
  a.lse = lse;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    // pixels flattened: (C, H*W, 1, F, B); a box = BK channels x PX pixels x F frames
    const uint64_t HWl = (uint64_t)H * W;
    const uint64_t dims[5] = {(uint64_t)C, HWl, 1u, (uint64_t)F, (uint64_t)B};
    const uint64_t str[4] = {(uint64_t)C * 2, HWl * C * 2, HWl * C * 2, (uint64_t)F * HWl * C * 2};
    const uint32_t box[5] = {(uint32_t)BK, (uint32_t)PX, 1u, (uint32_t)F, 1u};
    int rc = encode_tmap_bf16(&maps.x, x, 5, dims, str, box, BK * 2);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, 768};
    const uint64_t str[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {(uint32_t)BK, 96u};
    int rc = encode_tmap_bf16(&maps.w, w_hm, 2, dims, str, box, BK * 2);
    if (rc) return rc;
  }
  const int n_tiles = B * ((H * W + PX - 1) / PX);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (F == 10) return BK == 64 ? launch_tc<64, 10>(maps, a, n_tiles, st) : launch_tc<32, 10>(maps, a, n_tiles, st);
  return BK == 64 ? launch_tc<64, 16>(maps, a, n_tiles, st) : launch_tc<32, 16>(maps, a, n_tiles, st);
}

// ---------------------------------------------------------------------------------------
// Temporal attention core backward on tensor cores. Per (pixel tile, head), rows pixel-major:
//   S = Q K^T, dP = dO V^T                      (two 128x128x32 MMAs, block-diagonal use)
//   P = exp(S/sqrt(d) - lse), D = sum_j P dP, dS = P (dP - D)/sqrt(d)      (worker threads, fp32)
//   dQ = dS K, dK = dS^T Q, dV = P^T dO         (three 128x32x128 MMAs; the transposes are MN-major
//                                                reads of the same shared-memory tiles)
// TMEM: S [0,128) and dP [128,256); dQ|dK|dV reuse [0,96) once the workers hold their S/dP windows.
// ---------------------------------------------------------------------------------------
namespace vdn {

template <int F>
__global__ void __launch_bounds__(160) mha_tc_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                         const float* __restrict__ lse, bf16* __restrict__ dqkv, int B,
                                                         int H, int W, int PX) {
  constexpr int kTile = 128 * 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t in_ready, sdp_full, pds_ready, grad_full;
  __shared__ uint32_t tmem_base_smem;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTile;
  uint8_t* sV = sK + kTile;
  uint8_t* sG = sV + kTile;
  uint8_t* sP = sG + kTile;      // 32 KB
  uint8_t* sS = sP + 32768;      // 32 KB
  const int rows = PX * F;
  const int HW = H * W;
  const int tiles_img = (HW + PX - 1) / PX;
  const int tile = blockIdx.x;
  const int b = tile / tiles_img;
  const int p0 = (tile - b * tiles_img) * PX;

  pdl_trigger();
  for (int i = threadIdx.x; i < (4 * kTile + 65536) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&in_ready, 4);
    mbar_init(&sdp_full, 1);
    mbar_init(&pds_ready, 4);
    mbar_init(&grad_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 256);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t id_ss = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t id_dq = umma_idesc_bf16(128, 32, 0, 1);   // A K-major (dS), B MN-major (K)
      const uint32_t id_tt = umma_idesc_bf16(128, 32, 1, 1);   // A MN-major (dS^T / P^T), B MN-major
      const uint32_t l64 = umma_layout_type(64), l128 = umma_layout_type(128);
      for (int h = 0; h < 8; ++h) {
        const uint32_t par = (uint32_t)(h & 1);
        mbar_wait(&in_ready, par);
        tc_fence_after();
        {
          const uint64_t dq = umma_smem_desc(smem_u32(sQ), 16, 512, l64), dk = umma_smem_desc(smem_u32(sK), 16, 512, l64);
          const uint64_t dg = umma_smem_desc(smem_u32(sG), 16, 512, l64), dv = umma_smem_desc(smem_u32(sV), 16, 512, l64);
          umma_bf16(tmem_base, dq, dk, id_ss, 0u);
          umma_bf16(tmem_base, dq + 2, dk + 2, id_ss, 1u);
          umma_bf16(tmem_base + 128, dg, dv, id_ss, 0u);
          umma_bf16(tmem_base + 128, dg + 2, dv + 2, id_ss, 1u);
          tc_commit(&sdp_full);
        }
        mbar_wait(&pds_ready, par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t a_km = (uint32_t)((k >> 2) * 16384 + (k & 3) * 32);  // K-major slice of a [128x128] tile
          const uint32_t a_mn = (uint32_t)(k * 2048);                          // MN-major slice (16 rows)
          const uint32_t b_mn = (uint32_t)(k * 1024);                          // 16 token rows of a [128x32] tile
          const uint32_t acc = k != 0 ? 1u : 0u;
          umma_bf16(tmem_base, umma_smem_desc(smem_u32(sS) + a_km, 16, 1024, l128),
                    umma_smem_desc(smem_u32(sK) + b_mn, (uint32_t)kTile, 512, l64), id_dq, acc);
          umma_bf16(tmem_base + 32, umma_smem_desc(smem_u32(sS) + a_mn, 16384, 1024, l128),
                    umma_smem_desc(smem_u32(sQ) + b_mn, (uint32_t)kTile, 512, l64), id_tt, acc);
          umma_bf16(tmem_base + 64, umma_smem_desc(smem_u32(sP) + a_mn, 16384, 1024, l128),
                    umma_smem_desc(smem_u32(sG) + b_mn, (uint32_t)kTile, 512, l64), id_tt, acc);
        }
        tc_commit(&grad_full);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool in_tile = r < rows;
    const int px = in_tile ? r / F : 0, f = in_tile ? r % F : 0;
    const bool valid = in_tile && p0 + px < HW;  // rows of a partial last tile stay zero in every smem tile
    const long grow = ((long)b * F + f) * HW + p0 + px;
    const int px_lo = (quarter * 32) / F;
    constexpr int kSlots = (32 % F == 0) ? (32 / F) : ((31 / F) + 2);
    constexpr int kWin = kSlots * F > 32 ? 48 : 32;
    const int c0 = min(px_lo * F, 128 - kWin);   // clamped to the tile; the shift is a whole number of slots
    const int sel = px - px_lo + (px_lo * F - c0) / F;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const float scale = rsqrtf(32.f);
    for (int h = 0; h < 8; ++h) {
      const uint32_t par = (uint32_t)(h & 1);
      float L = 0.f;
      if (valid) {
        const uint4* qp = reinterpret_cast<const uint4*>(qkv + grow * 768 + h * 32);
        const uint4* gp = reinterpret_cast<const uint4*>(d_o + grow * 256 + h * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(sQ + sw64_off(r, c)) = qp[c];
          *reinterpret_cast<uint4*>(sK + sw64_off(r, c)) = qp[32 + c];   // +256 bf16 = 32 uint4
          *reinterpret_cast<uint4*>(sV + sw64_off(r, c)) = qp[64 + c];
          *reinterpret_cast<uint4*>(sG + sw64_off(r, c)) = gp[c];
        }
        L = lse[grow * 8 + h];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_ready);
      mbar_wait(&sdp_full, par);
      tc_fence_after();
      float sc[F], dp[F];
      {
        uint32_t win[kWin];
        tmem_ld_32x32(tmem_base + lane_sel + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&win[0]));
        if (kWin == 48) tmem_ld_32x16(tmem_base + lane_sel + (uint32_t)(c0 + 32), *reinterpret_cast<uint32_t(*)[16]>(&win[32]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < F; ++j) {
          float v = __uint_as_float(win[j]);
#pragma unroll
          for (int s = 1; s < kSlots; ++s)
            if (s * F + j < kWin) v = (sel == s) ? __uint_as_float(win[s * F + j]) : v;
          sc[j] = v;
        }
        tmem_ld_32x32(tmem_base + 128 + lane_sel + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&win[0]));
        if (kWin == 48) tmem_ld_32x16(tmem_base + 128 + lane_sel + (uint32_t)(c0 + 32), *reinterpret_cast<uint32_t(*)[16]>(&win[32]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < F; ++j) {
          float v = __uint_as_float(win[j]);
#pragma unroll
          for (int s = 1; s < kSlots; ++s)
            if (s * F + j < kWin) v = (sel == s) ? __uint_as_float(win[s * F + j]) : v;
          dp[j] = v;
        }
      }
      float D = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) {
        sc[j] = __expf(sc[j] * scale - L);   // P_ij
        D = fmaf(sc[j], dp[j], D);
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < F; j += 2) {
          const float ds0 = sc[j] * (dp[j] - D) * scale;
          const float ds1 = (j + 1 < F) ? sc[j + 1] * (dp[j + 1] - D) * scale : 0.f;
          const uint32_t off = sw128_off(r, px * F + j);
          *reinterpret_cast<uint32_t*>(sP + off) = pack_bf16x2(sc[j], (j + 1 < F) ? sc[j + 1] : 0.f);
          *reinterpret_cast<uint32_t*>(sS + off) = pack_bf16x2(ds0, ds1);
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_ready);
      mbar_wait(&grad_full, par);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + lane_sel + (uint32_t)(part * 32), raw);
        tmem_ld_wait();
        if (valid) {
          uint4* gp = reinterpret_cast<uint4*>(dqkv + grow * 768 + part * 256 + h * 32);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(raw[8 * c + 0]), __uint_as_float(raw[8 * c + 1]));
            u.y = pack_bf16x2(__uint_as_float(raw[8 * c + 2]), __uint_as_float(raw[8 * c + 3]));
            u.z = pack_bf16x2(__uint_as_float(raw[8 * c + 4]), __uint_as_float(raw[8 * c + 5]));
            u.w = pack_bf16x2(__uint_as_float(raw[8 * c + 6]), __uint_as_float(raw[8 * c + 7]));
            gp[c] = u;
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int F>
static int launch_tc_bwd(const bf16* qkv, const bf16* d_o, const float* lse, bf16* dqkv, int B, int H, int W, int PX,
                         cudaStream_t st) {
  const int smem = 4 * 8192 + 65536 + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(mha_tc_bwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "mha_tc_bwd cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  cudaError_t le = launch_pdl(mha_tc_bwd_kernel<F>, dim3(B * ((H * W + PX - 1) / PX)), dim3(160), (size_t)smem, st, 1, qkv, d_o,
                              lse, dqkv, B, H, W, PX);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_tc_bwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_tc_bwd_kernel");
}

}  // namespace vdn

// Same contract as vdn_mha_temporal_bwd (o is not needed: D_i = sum_j P_ij dP_ij). F in {10, 16}.
extern "C" int vdn_colsum(const void* dy, float* db, long P, int C, void* stream);

extern "C" int vdn_mha_temporal_tc_bwd(const void* qkv, const void* d_o, const float* lse, void* dqkv, float* dbias,
                                       int B, int F, int H, int W, void* stream) {
  VDN_REQUIRE(qkv && d_o && lse && dqkv && F >= 1 && F <= 16, VDN_E_SHAPE, "mha_tc_bwd: bad args (F <= 16)");
  const bool use_tcgen05 = tune_on("VDN_MHA_TC_BWD");  // block-diagonal tcgen05 kernel (A/B only)
  if (!use_tcgen05)
  {
    // (accumulating the bias column sums inside the MMA kernel costs more than the separate pass: measured
    //  +90 us on the 64x64 level from the extra live registers, vs 40 us for vdn_colsum)
    int rc = vdn::mha_temporal_mma_bwd_launch(qkv, d_o, lse, dqkv, B, F, H, W, reinterpret_cast<cudaStream_t>(stream));
    if (rc == 0 && dbias) rc = vdn_colsum(dqkv, dbias, (long)B * F * H * W, 768, stream);
    return rc;
  }
  VDN_REQUIRE(F == 10 || F == 16, VDN_E_SHAPE, "mha_tc_bwd: the tcgen05 kernel is instantiated for F in {10,16}");
  // The backward is bound by its per-row global loads / stores, not by the MMA chain: 120-row tiles (PX = 12)
  // measured slower than 80-row tiles, so it keeps power-of-two pixel counts (8 for F = 10 and F = 16).
  int PX = 1;
  while (PX * 2 * F <= 128 && PX * 2 <= H * W) PX *= 2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const vdn::bf16* q = reinterpret_cast<const vdn::bf16*>(qkv);
  const vdn::bf16* g = reinterpret_cast<const vdn::bf16*>(d_o);
  vdn::bf16* dq = reinterpret_cast<vdn::bf16*>(dqkv);
  int rc = F == 10 ? vdn::launch_tc_bwd<10>(q, g, lse, dq, B, H, W, PX, st) : vdn::launch_tc_bwd<16>(q, g, lse, dq, B, H, W, PX, st);
  if (rc == 0 && dbias) rc = vdn_colsum(dqkv, dbias, (long)B * F * H * W, 768, stream);
  return rc;
}
