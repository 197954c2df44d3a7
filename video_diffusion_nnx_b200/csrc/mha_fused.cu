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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
    if (lane == 0) {
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
    if (lane == 0) {
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
    float* kv = skv + (size_t)g * 128 * kKvPitch;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 96);
    const float scale = rsqrtf(32.f);
    const int bar_id = 1 + g;  // named barrier per head group (128 threads)
    for (int pass = 0; pass < 4; ++pass) {
      const int h = pass * 2 + g;
      mbar_wait(&tmem_full_bar, (uint32_t)(pass & 1));
      tc_fence_after();
      const float* bh = a.bias ? a.bias + h * 96 : nullptr;
      float* myrow = kv + (size_t)r * kKvPitch;
      float q[32];
      // k, v, q one after the other (keeps the live register set small). Values are rounded to bf16
      // exactly like the unfused path stores them, so forward and backward see the same q, k, v.
#pragma unroll
      for (int part = 2; part >= 0; --part) {   // v, k, then q
        uint32_t raw[32];
        tmem_ld_32x32(taddr + (uint32_t)(part * 32), raw);
        tmem_ld_wait();
        if (part == 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar);  // this warp's TMEM reads of the pass are done
        }
        uint32_t packed[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float v0 = __uint_as_float(raw[e]), v1 = __uint_as_float(raw[e + 1]);
          if (bh) {
            v0 += __ldg(bh + part * 32 + e);
            v1 += __ldg(bh + part * 32 + e + 1);
          }
          packed[e >> 1] = pack_bf16x2(v0, v1);
        }
        if (a.qkv && valid) {
          // the unfused layout [P][q(256) | k(256) | v(256)] consumed by the backward kernels
          uint4* dstp = reinterpret_cast<uint4*>(a.qkv + grow * 768 + part * 256 + h * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dstp[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        if (part == 0) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 f2 = unpack_bf16x2(packed[e >> 1]);
            q[e] = f2.x * scale;
            q[e + 1] = f2.y * scale;
          }
        } else {
          float* dstrow = myrow + (part == 1 ? 0 : 32);
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float2 f0 = unpack_bf16x2(packed[e >> 1]), f1 = unpack_bf16x2(packed[(e >> 1) + 1]);
            *reinterpret_cast<float4*>(dstrow + e) = make_float4(f0.x, f0.y, f1.x, f1.y);
          }
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      float acc[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) acc[e] = 0.f;
      float mx = -INFINITY, l = 0.f;
      for (int j = 0; j < a.F; ++j) {
        const float* rowj = kv + (size_t)(j * a.PX + px) * kKvPitch;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 k4 = *reinterpret_cast<const float4*>(rowj + e);
          s0 = fmaf(q[e], k4.x, s0); s1 = fmaf(q[e + 1], k4.y, s1);
          s2 = fmaf(q[e + 2], k4.z, s2); s3 = fmaf(q[e + 3], k4.w, s3);
        }
        const float sc = (s0 + s1) + (s2 + s3);
        const float mn = fmaxf(mx, sc);
        const float corr = __expf(mx - mn);
        const float p = __expf(sc - mn);
        l = l * corr + p;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(rowj + 32 + e);
          acc[e] = fmaf(acc[e], corr, p * v4.x); acc[e + 1] = fmaf(acc[e + 1], corr, p * v4.y);
          acc[e + 2] = fmaf(acc[e + 2], corr, p * v4.z); acc[e + 3] = fmaf(acc[e + 3], corr, p * v4.w);
        }
        mx = mn;
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
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // kv buffer free for the next head
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
  const int smem = kFStages * stage + 2 * 128 * kKvPitch * 4 + 1024;
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
  VDN_REQUIRE(x && w_hm && o && B > 0 && F > 0 && F <= 128 && H > 0 && W > 0, VDN_E_SHAPE, "mha_fused: bad args");
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
