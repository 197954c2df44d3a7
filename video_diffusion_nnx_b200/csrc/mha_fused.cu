// Fused temporal attention forward: QKV projection on tcgen05 + the F x F attention core, without
// materialising qkv (reference: Residual(PreNorm(EinopsToAndFrom('b f h w c','b (h w) f c', MHA))),
// unet3d.py:86-96,118-120, modules.py:285-323; the out projection + residual stay in tapgemm).
//
// One CTA owns PX adjacent pixels (along w) x all F frames = PX*F <= 128 token rows, r = f*PX + px,
// loaded by ONE 5-D TMA box (channels, PX, 1, F, 1) per K chunk - the 'b f h w c -> b (h w) f c'
// rearrangement is just this box. The projection runs in 4 passes of 2 heads (N = 192 = 2 x (q|k|v) x 32,
// head-major repacked weights) into TMEM; two groups of 4 epilogue warps (one head each) pull their
// head's q,k,v out of TMEM (+bias), exchange k,v through shared memory and run the softmax(q k^T) v
// over the F frames of their own pixel on CUDA cores (fp32), then store o (and, for training, qkv and
// the log-sum-exp needed by the backward).
#include <algorithm>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kFThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2-5 head A, warps 6-9 head B
constexpr int kFStages = 2;
constexpr int kPassCols = 192;      // 2 heads x 96
constexpr int kKvPitch = 68;        // floats per smem row (64 + 4 pad: conflict-free LDS.128 across rows)

struct FusedMaps {
  CUtensorMap x;   // (C, W, H, F, B) bf16
  CUtensorMap w;   // (C, 768) bf16 head-major rows
};

struct FusedArgs {
  int B, F, H, W, C;
  int PX;            // pixels per tile
  int chunks;        // C / BK
  const float* bias; // [768] head-major, or null
  bf16* o;           // [P][256]
  bf16* qkv;         // [P][768] (q|k|v layout of the unfused path) or null
  float* lse;        // [P][8] or null
};

template <int BK>
__global__ void __launch_bounds__(kFThreads, 2) mha_fused_fwd_kernel(const __grid_constant__ FusedMaps maps,
                                                                  const FusedArgs a) {
  constexpr int kSw = BK * 2;
  constexpr int kXBytes = 128 * BK * 2;
  constexpr int kWBytes = kPassCols * BK * 2;  // 192 rows
  constexpr int kWBytesPad = (kWBytes + 1023) & ~1023;
  constexpr int kStageBytes = kXBytes + kWBytesPad;
  constexpr uint32_t kLayout = umma_layout_type(kSw);
  constexpr uint32_t kSBO = 8 * kSw;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kFStages];
  __shared__ __align__(8) uint64_t empty_bar[kFStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ __align__(8) uint64_t tmem_empty_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[768];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;
  // 1024-byte aligned view of the dynamic smem; offset arithmetic keeps the pointer in the shared space
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_bias[i] = a.bias ? a.bias[i] : 0.f;
  float* skv = reinterpret_cast<float*>(smem + kFStages * kStageBytes);  // [2 groups][128][kKvPitch]

  const int tiles_x = a.W / a.PX;
  const int tile = blockIdx.x;
  const int x0 = (tile % tiles_x) * a.PX;
  const int y = (tile / tiles_x) % a.H;
  const int b = tile / (tiles_x * a.H);
  const int n_steps = 4 * a.chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w);
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, 8);  // one arrival per epilogue warp
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < n_steps; ++it) {
        const int pass = it / a.chunks, c = it - pass * a.chunks;
        const int st = it % kFStages;
        const uint32_t ph = (uint32_t)(it / kFStages) & 1u;
        mbar_wait(&empty_bar[st], ph ^ 1u);
        uint8_t* sx = smem + st * kStageBytes;
        uint8_t* sw = sx + kXBytes;
        mbar_expect_tx(&full_bar[st], (uint32_t)(a.PX * a.F * BK * 2 + kWBytes));
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(sx)),
            "l"(reinterpret_cast<uint64_t>(&maps.x)), "r"(smem_u32(&full_bar[st])), "r"(c * BK), "r"(x0), "r"(y),
            "r"(0), "r"(b)
            : "memory");
        tma_load_2d(sw, &maps.w, &full_bar[st], c * BK, pass * kPassCols);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, kPassCols, 0, 0);
      int it = 0;
      for (int pass = 0; pass < 4; ++pass) {
        mbar_wait(&tmem_empty_bar, (uint32_t)(pass & 1) ^ 1u);  // epilogue drained the previous pass
        tc_fence_after();
        for (int c = 0; c < a.chunks; ++c, ++it) {
          const int st = it % kFStages;
          const uint32_t ph = (uint32_t)(it / kFStages) & 1u;
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint32_t sx = smem_u32(smem + st * kStageBytes);
          const uint32_t sw = sx + kXBytes;
          const uint64_t da = umma_smem_desc(sx, 16, kSBO, kLayout);
          const uint64_t db = umma_smem_desc(sw, 16, kSBO, kLayout);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (c | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[st]);
        }
        tc_commit(&tmem_full_bar);
      }
    }
    __syncwarp();
  } else {
    // ---------------- attention: group g = head (2*pass + g), thread = token row r = f*PX + px ----------------
    const int g = (warp - 2) >> 2;          // 0 / 1
    const int quarter = warp & 3;           // TMEM lane quarter of this warp
    const int r = quarter * 32 + lane;
    const int rows = a.PX * a.F;
    const bool valid = r < rows;
    const int f = valid ? r / a.PX : 0;
    const int px = valid ? r - f * a.PX : 0;
    const long grow = (((long)b * a.F + f) * a.H + y) * a.W + x0 + px;  // global token row
    float* kv = skv + (size_t)g * rows * kKvPitch;   // only the PX*F real rows are exchanged
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 96);
    const float scale = rsqrtf(32.f);
    const int bar_id = 1 + g;  // named barrier per head group
    const int nw_active = (rows + 31) >> 5;        // warps of the group that own real rows
    const bool warp_active = quarter < nw_active;  // rows >= PX*F are padding: those warps only hand TMEM back
    for (int pass = 0; pass < 4; ++pass) {
      const int h = pass * 2 + g;
      mbar_wait(&tmem_full_bar, (uint32_t)(pass & 1));
      tc_fence_after();
      if (!warp_active) {
        if (lane == 0) mbar_arrive(&tmem_empty_bar);
        continue;
      }
      const float* bh = s_bias + h * 96;
      float* myrow = kv + (size_t)(valid ? r : 0) * kKvPitch;
      float q[32];
      // v, k, then q: one 32-column TMEM load each (small live register set). In training mode the
      // values are rounded to bf16 exactly as stored for the backward; in inference they stay fp32.
#pragma unroll
      for (int part = 2; part >= 0; --part) {
        uint32_t raw[32];
        tmem_ld_32x32(taddr + (uint32_t)(part * 32), raw);
        tmem_ld_wait();
        if (part == 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar);  // this warp's TMEM reads of the pass are done
        }
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bh + part * 32 + e);
          v[e] = __uint_as_float(raw[e]) + b4.x;
          v[e + 1] = __uint_as_float(raw[e + 1]) + b4.y;
          v[e + 2] = __uint_as_float(raw[e + 2]) + b4.z;
          v[e + 3] = __uint_as_float(raw[e + 3]) + b4.w;
        }
        if (a.qkv) {
          uint32_t packed[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            packed[e >> 1] = pack_bf16x2(v[e], v[e + 1]);
            const float2 f2 = unpack_bf16x2(packed[e >> 1]);
            v[e] = f2.x;
            v[e + 1] = f2.y;
          }
          if (valid) {
            // the unfused layout [P][q(256) | k(256) | v(256)] consumed by the backward kernels
            uint4* dstp = reinterpret_cast<uint4*>(a.qkv + grow * 768 + part * 256 + h * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dstp[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
        }
        if (part == 0) {
#pragma unroll
          for (int e = 0; e < 32; ++e) q[e] = v[e] * scale;
        } else if (valid) {
          float* dstrow = myrow + (part == 1 ? 0 : 32);
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dstrow + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nw_active * 32) : "memory");
      // scores for all F <= 16 keys first, then one softmax, then the weighted sum of v (no rescaling)
      float sc[16];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        sc[j] = -INFINITY;
        if (j < a.F) {
          const float* rowj = kv + (size_t)(j * a.PX + px) * kKvPitch;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(rowj + e);
            s0 = fmaf(q[e], k4.x, s0); s1 = fmaf(q[e + 1], k4.y, s1);
            s2 = fmaf(q[e + 2], k4.z, s2); s3 = fmaf(q[e + 3], k4.w, s3);
          }
          sc[j] = (s0 + s1) + (s2 + s3);
          mx = fmaxf(mx, sc[j]);
        }
      }
      float acc[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = 0.f;
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < a.F) {
          const float p = __expf(sc[j] - mx);
          l += p;
          const float* rowj = kv + (size_t)(j * a.PX + px) * kKvPitch + 32;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(rowj + e);
            acc[e] = fmaf(p, v4.x, acc[e]); acc[e + 1] = fmaf(p, v4.y, acc[e + 1]);
            acc[e + 2] = fmaf(p, v4.z, acc[e + 2]); acc[e + 3] = fmaf(p, v4.w, acc[e + 3]);
          }
        }
      }
      const float inv = 1.f / l;
      if (valid) {
        bf16* op = a.o + grow * 256 + h * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_bf16x2(acc[8 * j + 0] * inv, acc[8 * j + 1] * inv);
          u.y = pack_bf16x2(acc[8 * j + 2] * inv, acc[8 * j + 3] * inv);
          u.z = pack_bf16x2(acc[8 * j + 4] * inv, acc[8 * j + 5] * inv);
          u.w = pack_bf16x2(acc[8 * j + 6] * inv, acc[8 * j + 7] * inv);
          reinterpret_cast<uint4*>(op)[j] = u;
        }
        if (a.lse) a.lse[grow * 8 + h] = mx + __logf(l);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nw_active * 32) : "memory");  // kv buffer free for the next head
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// head-major repack of the fused qkv projection: dst[(h*96 + part*32 + d)][c] = w[c][part*256 + h*32 + d]
__global__ void qkv_headmajor_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                          bf16* __restrict__ dst, float* __restrict__ bias_dst, int C) {
  const int total = 768 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, row = i / C;
    const int h = row / 96, part = (row % 96) / 32, d = row % 32;
    dst[i] = __float2bfloat16(w[(long)c * 768 + part * 256 + h * 32 + d]);
  }
  if (bias && blockIdx.x == 0)
    for (int row = threadIdx.x; row < 768; row += blockDim.x) {
      const int h = row / 96, part = (row % 96) / 32, d = row % 32;
      bias_dst[row] = bias[part * 256 + h * 32 + d];
    }
}

template <int BK>
static int launch_fused(const FusedMaps& maps, const FusedArgs& a, int n_tiles, cudaStream_t st) {
  constexpr int stage = 128 * BK * 2 + ((kPassCols * BK * 2 + 1023) & ~1023);
  const int smem = kFStages * stage + 2 * (a.PX * a.F) * kKvPitch * 4 + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(mha_fused_fwd_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "mha_fused cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  mha_fused_fwd_kernel<BK><<<n_tiles, kFThreads, smem, st>>>(maps, a);
  return check_launch("mha_fused_fwd_kernel");
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_qkv_headmajor_pack(const float* w, const float* bias, void* dst, float* bias_dst, int C,
                                      void* stream) {
  VDN_REQUIRE(w && dst && C > 0, VDN_E_SHAPE, "qkv_headmajor_pack: bad args");
  qkv_headmajor_pack_kernel<<<std::min(148 * 4, (768 * C + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, bias, reinterpret_cast<bf16*>(dst), bias_dst, C);
  return check_launch("qkv_headmajor_pack");
}

// x bf16 (B,F,H,W,C); w_hm bf16 [768][C] and bias_hm fp32 [768] from vdn_qkv_headmajor_pack;
// o bf16 [P][256]; qkv bf16 [P][768] and lse fp32 [P][8] optional (training).
extern "C" int vdn_mha_temporal_fused_fwd(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv,
                                          float* lse, int B, int F, int H, int W, int C, void* stream) {
  VDN_REQUIRE(x && w_hm && o && B > 0 && F > 0 && F <= 16 && H > 0 && W > 0, VDN_E_SHAPE, "mha_fused: bad args (F <= 16)");
  VDN_REQUIRE(C % 16 == 0, VDN_E_SHAPE, "mha_fused: C=%d must be a multiple of 16", C);
  int PX = 1;
  while (PX * 2 * F <= 128 && PX * 2 <= W && W % (PX * 2) == 0) PX *= 2;
  const int BK = (C % 64 == 0) ? 64 : (C % 32 == 0) ? 32 : 16;
  FusedArgs a;
  a.B = B; a.F = F; a.H = H; a.W = W; a.C = C; a.PX = PX; a.chunks = C / BK;
  a.bias = bias_hm;
  a.o = reinterpret_cast<bf16*>(o);
  a.qkv = reinterpret_cast<bf16*>(qkv);
  a.lse = lse;
  FusedMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)F, (uint64_t)B};
    const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)F * H * W * C * 2};
    const uint32_t box[5] = {(uint32_t)BK, (uint32_t)PX, 1u, (uint32_t)F, 1u};
    int rc = encode_tmap_bf16(&maps.x, x, 5, dims, str, box, BK * 2);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, 768};
    const uint64_t str[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)kPassCols};
    int rc = encode_tmap_bf16(&maps.w, w_hm, 2, dims, str, box, BK * 2);
    if (rc) return rc;
  }
  const int n_tiles = B * H * (W / PX);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (BK == 64) return launch_fused<64>(maps, a, n_tiles, st);
  if (BK == 32) return launch_fused<32>(maps, a, n_tiles, st);
  return launch_fused<16>(maps, a, n_tiles, st);
}

// ---------------------------------------------------------------------------------------
// Temporal attention core backward, one kernel: grid (pixel tiles, heads), thread = token (f, px).
// k, v, q, dO of the tile's head are exchanged through shared memory; P and dS (F x F per pixel) are
// computed once by the query threads and re-used by the key threads:
//   D_i = dO_i.O_i ; P_ij = exp(q_i.k_j/sqrt(d) - lse_i) ; dS_ij = P_ij (dO_i.v_j - D_i)/sqrt(d)
//   dq_i = sum_j dS_ij k_j ; dk_j = sum_i dS_ij q_i ; dv_j = sum_i P_ij dO_i
// ---------------------------------------------------------------------------------------
namespace vdn {

constexpr int kBwPitch = 36;  // floats per smem row (32 + 4): conflict-free LDS.128 across the 8 rows of a warp

__global__ void __launch_bounds__(128, 4) mha_temporal_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                               const bf16* __restrict__ d_o,
                                                               const float* __restrict__ lse, bf16* __restrict__ dqkv,
                                                               int B, int F, int H, int W, int PX) {
  extern __shared__ __align__(16) float sm[];
  const int rows = PX * F;
  float* sK = sm;
  float* sV = sK + rows * kBwPitch;
  float* sQ = sV + rows * kBwPitch;
  float* sG = sQ + rows * kBwPitch;          // dO
  float* sP = sG + rows * kBwPitch;          // [px][i][j]
  float* sS = sP + PX * F * F;               // dS [px][i][j]
  const int tiles_x = W / PX;
  const int tile = blockIdx.x, h = blockIdx.y;
  const int x0 = (tile % tiles_x) * PX;
  const int y = (tile / tiles_x) % H;
  const int b = tile / (tiles_x * H);
  const int r = threadIdx.x;
  const bool valid = r < rows;
  const int f = valid ? r / PX : 0;
  const int px = valid ? r - f * PX : 0;
  const long grow = (((long)b * F + f) * H + y) * W + x0 + px;
  const float scale = rsqrtf(32.f);
  float q[32], g[32];
  float D = 0.f, L = 0.f;
  if (valid) {
    float t[32];
    const bf16* qp = qkv + grow * 768 + h * 32;
    // own q, dO (kept in registers and published to smem), k, v (published)
    auto ld = [&](const bf16* p, float (&v)[32]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 u = reinterpret_cast<const uint4*>(p)[c];
        float2 x2;
        x2 = unpack_bf16x2(u.x); v[8 * c + 0] = x2.x; v[8 * c + 1] = x2.y;
        x2 = unpack_bf16x2(u.y); v[8 * c + 2] = x2.x; v[8 * c + 3] = x2.y;
        x2 = unpack_bf16x2(u.z); v[8 * c + 4] = x2.x; v[8 * c + 5] = x2.y;
        x2 = unpack_bf16x2(u.w); v[8 * c + 6] = x2.x; v[8 * c + 7] = x2.y;
      }
    };
    auto st = [&](float* dst, const float (&v)[32]) {
#pragma unroll
      for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
    };
    ld(qp, q);
    st(sQ + r * kBwPitch, q);
    ld(qp + 256, t);
    st(sK + r * kBwPitch, t);
    ld(qp + 512, t);
    st(sV + r * kBwPitch, t);
    ld(d_o + grow * 256 + h * 32, g);
    st(sG + r * kBwPitch, g);
    ld(o + grow * 256 + h * 32, t);
#pragma unroll
    for (int e = 0; e < 32; ++e) D = fmaf(g[e], t[e], D);
    L = lse[grow * 8 + h];
  }
  __syncthreads();
  float acc[32];
  if (valid) {
    // ---- as query i = f: P_ij, dS_ij, dq_i ----
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = 0.f;
    for (int j = 0; j < F; ++j) {
      const float* kj = sK + (j * PX + px) * kBwPitch;
      const float* vj = sV + (j * PX + px) * kBwPitch;
      float s0 = 0.f, s1 = 0.f, d0 = 0.f, d1 = 0.f;
      float kreg[32];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(kj + e);
        const float4 v4 = *reinterpret_cast<const float4*>(vj + e);
        kreg[e] = k4.x; kreg[e + 1] = k4.y; kreg[e + 2] = k4.z; kreg[e + 3] = k4.w;
        s0 = fmaf(q[e], k4.x, s0); s1 = fmaf(q[e + 1], k4.y, s1);
        s0 = fmaf(q[e + 2], k4.z, s0); s1 = fmaf(q[e + 3], k4.w, s1);
        d0 = fmaf(g[e], v4.x, d0); d1 = fmaf(g[e + 1], v4.y, d1);
        d0 = fmaf(g[e + 2], v4.z, d0); d1 = fmaf(g[e + 3], v4.w, d1);
      }
      const float p = __expf((s0 + s1) * scale - L);
      const float ds = p * ((d0 + d1) - D) * scale;
      sP[(px * F + f) * F + j] = p;
      sS[(px * F + f) * F + j] = ds;
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = fmaf(ds, kreg[e], acc[e]);
    }
    bf16* dq = dqkv + grow * 768 + h * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 u;
      u.x = pack_bf16x2(acc[8 * c + 0], acc[8 * c + 1]); u.y = pack_bf16x2(acc[8 * c + 2], acc[8 * c + 3]);
      u.z = pack_bf16x2(acc[8 * c + 4], acc[8 * c + 5]); u.w = pack_bf16x2(acc[8 * c + 6], acc[8 * c + 7]);
      reinterpret_cast<uint4*>(dq)[c] = u;
    }
  }
  __syncthreads();
  if (valid) {
    // ---- as key j = f: dk_j, dv_j ----
    float dv[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = dv[e] = 0.f;
    for (int i = 0; i < F; ++i) {
      const float p = sP[(px * F + i) * F + f];
      const float ds = sS[(px * F + i) * F + f];
      const float* qi = sQ + (i * PX + px) * kBwPitch;
      const float* gi = sG + (i * PX + px) * kBwPitch;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(qi + e);
        const float4 g4 = *reinterpret_cast<const float4*>(gi + e);
        acc[e] = fmaf(ds, q4.x, acc[e]); acc[e + 1] = fmaf(ds, q4.y, acc[e + 1]);
        acc[e + 2] = fmaf(ds, q4.z, acc[e + 2]); acc[e + 3] = fmaf(ds, q4.w, acc[e + 3]);
        dv[e] = fmaf(p, g4.x, dv[e]); dv[e + 1] = fmaf(p, g4.y, dv[e + 1]);
        dv[e + 2] = fmaf(p, g4.z, dv[e + 2]); dv[e + 3] = fmaf(p, g4.w, dv[e + 3]);
      }
    }
    bf16* dk = dqkv + grow * 768 + 256 + h * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 u;
      u.x = pack_bf16x2(acc[8 * c + 0], acc[8 * c + 1]); u.y = pack_bf16x2(acc[8 * c + 2], acc[8 * c + 3]);
      u.z = pack_bf16x2(acc[8 * c + 4], acc[8 * c + 5]); u.w = pack_bf16x2(acc[8 * c + 6], acc[8 * c + 7]);
      reinterpret_cast<uint4*>(dk)[c] = u;
      u.x = pack_bf16x2(dv[8 * c + 0], dv[8 * c + 1]); u.y = pack_bf16x2(dv[8 * c + 2], dv[8 * c + 3]);
      u.z = pack_bf16x2(dv[8 * c + 4], dv[8 * c + 5]); u.w = pack_bf16x2(dv[8 * c + 6], dv[8 * c + 7]);
      reinterpret_cast<uint4*>(dk + 256)[c] = u;
    }
  }
}

}  // namespace vdn

// qkv bf16 [P][768], o / d_o bf16 [P][256], lse fp32 [P][8] (from the forward) -> dqkv bf16 [P][768]
extern "C" int vdn_mha_temporal_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                                    int B, int F, int H, int W, void* stream) {
  VDN_REQUIRE(qkv && o && d_o && lse && dqkv && F >= 1 && F <= 16, VDN_E_SHAPE, "mha_temporal_bwd: bad args (F <= 16)");
  int PX = 1;
  while (PX * 2 * F <= 128 && PX * 2 <= W && W % (PX * 2) == 0) PX *= 2;
  const int rows = PX * F;
  const size_t smem = (size_t)(4 * rows * kBwPitch + 2 * PX * F * F) * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(mha_temporal_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cfg = true;
  }
  mha_temporal_bwd_kernel<<<dim3(B * H * (W / PX), 8), 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o), lse,
      reinterpret_cast<bf16*>(dqkv), B, F, H, W, PX);
  return check_launch("mha_temporal_bwd_kernel");
}
