// SpatialLinearAttention apply pass of the inference engines (C = 32) as two tcgen05 GEMMs:
//   out = x + to_out(softmax_D(x W_q) ctx)                       (modules.py:105-123, q NOT scaled; unet3d.py:170-178)
// With ctx_h [32 x 32] known per (image, head) - pass 1 of vdn_sla_fused_fwd - everything after the feature softmax is
// linear, so ctx and to_out fold into one [256 x 32] matrix per image:
//   out[n, c] = x[n, c] + sum_h sum_d q~_h[n, d] G_h[d][c],   G_h = ctx_h W_out,h                 (sla_fold_g_kernel)
// and a tile of 128 tokens is
//   1. TMA: x tile [128 x 32] -> smem (K-major, SW64)                 2. tcgen05: Q[128 x 256] = X W_q -> TMEM
//   3. TMEM -> registers: softmax over the 32 features of each head, entirely inside the thread that owns the row
//      -> bf16 -> smem in the K-major SW128 chunk layout              4. tcgen05: O[128 x 32] = Q~ G_b (K = 256) -> TMEM
//   5. TMEM -> + x -> bf16 -> global.
// No q / k / v / tok tensors, no mma.sync: the warp-MMA version (sla_apply_fused_kernel, sla_mma.cu) ran at the
// legacy-MMA issue rate (241 us at 16 samples). Two CTAs per SM; the bound is the TMEM read of Q (128 KB per tile at
// 64 B/clk per SM).
#include <algorithm>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {
namespace {

constexpr int kSaThreads = 256;
constexpr int kSaXBytes = 128 * 64;         // x tile, SW64
constexpr int kSaWBytes = 256 * 64;         // W_q [256][32], SW64
constexpr int kSaGBytes = 4 * 32 * 128;     // G_b^T [32][256] as 4 K-chunks of [32 rows][128 B], SW128
constexpr int kSaQBytes = 4 * 128 * 128;    // q~: 4 K-chunks of [128 rows][128 B], SW128
constexpr int kSaSmem = 1024 + kSaXBytes + kSaWBytes + kSaGBytes + kSaQBytes;

struct SaMaps {
  CUtensorMap x, w, g;
};

// Gt[img][c][h*32 + d] = sum_e ctx[img][h][d][e] * w_out[c][h*32 + e]   (bf16, the K-major B operand of GEMM 2)
__global__ void __launch_bounds__(256) sla_fold_g_kernel(const float* __restrict__ ctx, const bf16* __restrict__ w_out,
                                                         bf16* __restrict__ gt) {
  __shared__ float s_ctx[32][33];
  __shared__ float s_w[32][33];  // [c][e]
  const int h = blockIdx.x & 7, img = blockIdx.x >> 3;
  const float* cp = ctx + ((size_t)img * 8 + h) * 1024;
  for (int i = threadIdx.x; i < 1024; i += 256) {
    s_ctx[i >> 5][i & 31] = cp[i];
    s_w[i >> 5][i & 31] = __bfloat162float(w_out[(i >> 5) * 256 + h * 32 + (i & 31)]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += 256) {
    const int c = i >> 5, d = i & 31;
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) acc = fmaf(s_ctx[d][e], s_w[c][e], acc);
    gt[((size_t)img * 32 + c) * 256 + h * 32 + d] = __float2bfloat16(acc);
  }
}

__global__ void __launch_bounds__(kSaThreads, 2) sla_apply_tc_kernel(const __grid_constant__ SaMaps maps,
                                                                    const bf16* __restrict__ x, bf16* __restrict__ out,
                                                                    int tiles_per_img, int n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_bar, g_bar, w_bar, mma_bar;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* sX = smem;
  uint8_t* sW = sX + kSaXBytes;
  uint8_t* sG = sW + kSaWBytes;
  uint8_t* sQ = sG + kSaGBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.g);
    mbar_init(&x_bar, 1);
    mbar_init(&g_bar, 1);
    mbar_init(&w_bar, 1);
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 256u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  int tile = blockIdx.x;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&w_bar, (uint32_t)kSaWBytes);
    tma_load_2d(sW, &maps.w, &w_bar, 0, 0);
    if (tile < n_tiles) {
      mbar_expect_tx(&x_bar, (uint32_t)kSaXBytes);
      tma_load_2d(sX, &maps.x, &x_bar, 0, tile * 128);
      mbar_expect_tx(&g_bar, (uint32_t)kSaGBytes);
      for (int kc = 0; kc < 4; ++kc) tma_load_2d(sG + kc * 4096, &maps.g, &g_bar, kc * 64, (tile / tiles_per_img) * 32);
    }
  }
  mbar_wait(&w_bar, 0);

  const uint32_t desc_hi64 = ((8u * 64u) >> 4) | (1u << 14) | (umma_layout_type(64) << 29);
  const uint32_t desc_hi128 = ((8u * 128u) >> 4) | (1u << 14) | (umma_layout_type(128) << 29);
  const uint32_t sx16 = smem_u32(sX) >> 4, sw16 = smem_u32(sW) >> 4, sg16 = smem_u32(sG) >> 4, sq16 = smem_u32(sQ) >> 4;
  uint32_t x_ph = 0u, g_ph = 0u, mma_ph = 0u;
  const int quarter = warp & 3, half = warp >> 2;
  const int r = quarter * 32 + lane;  // tile row of this thread in the TMEM phases
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);

  for (; tile < n_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    // ---- 1 + 2: Q = X W_q ----
    mbar_wait(&x_bar, x_ph);
    x_ph ^= 1u;
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi64) << 32) | ((sx16 + 2u * k) | (1u << 16)),
                  (static_cast<uint64_t>(desc_hi64) << 32) | ((sw16 + 2u * k) | (1u << 16)), idesc, k);
      tc_commit(&mma_bar);
    }
    // residual rows of x for the epilogue: requested now, used after GEMM 2 (the x tile in smem is recycled before)
    const long o_row = (long)tile * 128 + r;
    uint4 xres[2];
    {
      const uint4* xr = reinterpret_cast<const uint4*>(x + o_row * 32) + 2 * half;
      xres[0] = __ldg(xr);
      xres[1] = __ldg(xr + 1);
    }
    mbar_wait(&mma_bar, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    if (threadIdx.x == 0 && next < n_tiles) {  // the x tile has been consumed: prefetch the next one
      mbar_expect_tx(&x_bar, (uint32_t)kSaXBytes);
      tma_load_2d(sX, &maps.x, &x_bar, 0, next * 128);
    }
    // ---- 3: softmax over the 32 features of each head (thread-local), -> bf16 -> smem ----
    {
      auto stage_head = [&](const uint32_t (&raw)[32], int hh) {
        float v[32];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = __uint_as_float(raw[i]);
          m = fmaxf(m, v[i]);
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          v[i] = __expf(v[i] - m);
          v[i + 1] = __expf(v[i + 1] - m);
          s0 += v[i];
          s1 += v[i + 1];
        }
        const float inv = __fdividef(1.f, s0 + s1);
        uint8_t* rowp = sQ + (hh >> 1) * (128 * 128) + r * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 q;
          q.x = pack_bf16x2(v[8 * c + 0] * inv, v[8 * c + 1] * inv);
          q.y = pack_bf16x2(v[8 * c + 2] * inv, v[8 * c + 3] * inv);
          q.z = pack_bf16x2(v[8 * c + 4] * inv, v[8 * c + 5] * inv);
          q.w = pack_bf16x2(v[8 * c + 6] * inv, v[8 * c + 7] * inv);
          *reinterpret_cast<uint4*>(rowp + (((uint32_t)((hh & 1) * 4 + c) ^ (uint32_t)(r & 7)) << 4)) = q;
        }
      };
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
    fence_proxy_async_smem();
    __syncthreads();
    // ---- 4: O = Q~ G_b ----
    if (threadIdx.x == 0) {
      mbar_wait(&g_bar, g_ph);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        const uint32_t a0 = sq16 + (uint32_t)((kc * 128 * 128) >> 4);
        const uint32_t b0 = sg16 + (uint32_t)((kc * 32 * 128) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi128) << 32) | ((a0 + 2u * k) | (1u << 16)),
                    (static_cast<uint64_t>(desc_hi128) << 32) | ((b0 + 2u * k) | (1u << 16)), idesc, (kc | k) != 0 ? 1u : 0u);
      }
      tc_commit(&mma_bar);
    }
    g_ph ^= 1u;
    mbar_wait(&mma_bar, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    if (threadIdx.x == 0 && next < n_tiles) {  // G_b has been consumed: fetch the next tile's image matrix
      mbar_expect_tx(&g_bar, (uint32_t)kSaGBytes);
      for (int kc = 0; kc < 4; ++kc) tma_load_2d(sG + kc * 4096, &maps.g, &g_bar, kc * 64, (next / tiles_per_img) * 32);
    }
    // ---- 5: + x -> bf16 -> global (4 lane quarters x 2 column halves) ----
    {
      uint32_t raw[16];
      tmem_ld_32x16(taddr + (uint32_t)(16 * half), raw);
      tmem_ld_wait();
      uint4* op = reinterpret_cast<uint4*>(out + o_row * 32) + 2 * half;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint4 xv = xres[c];
        float2 a;
        uint4 q;
        a = unpack_bf16x2(xv.x); q.x = pack_bf16x2(__uint_as_float(raw[8 * c + 0]) + a.x, __uint_as_float(raw[8 * c + 1]) + a.y);
        a = unpack_bf16x2(xv.y); q.y = pack_bf16x2(__uint_as_float(raw[8 * c + 2]) + a.x, __uint_as_float(raw[8 * c + 3]) + a.y);
        a = unpack_bf16x2(xv.z); q.z = pack_bf16x2(__uint_as_float(raw[8 * c + 4]) + a.x, __uint_as_float(raw[8 * c + 5]) + a.y);
        a = unpack_bf16x2(xv.w); q.w = pack_bf16x2(__uint_as_float(raw[8 * c + 6]) + a.x, __uint_as_float(raw[8 * c + 7]) + a.y);
        op[c] = q;
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM columns and the q~ buffer are free for the next tile
    tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
  }
}

}  // namespace

bool sla_apply_tc_applicable(int N) { return N >= 128 && N % 128 == 0; }
size_t sla_apply_tc_scratch_bytes(int n_img) { return (size_t)n_img * 32 * 256 * sizeof(bf16); }

// gt_ws: scratch of sla_apply_tc_scratch_bytes(n_img), 16-byte aligned (the folded per-image matrices)
int sla_apply_tc_launch(const void* x, const void* w_qkv, const void* w_out, const float* ctx, void* gt_ws, void* out,
                        int n_img, int N, cudaStream_t st) {
  sla_fold_g_kernel<<<n_img * 8, 256, 0, st>>>(ctx, reinterpret_cast<const bf16*>(w_out), reinterpret_cast<bf16*>(gt_ws));
  int rc = check_launch("sla_fold_g");
  if (rc) return rc;
  SaMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    const uint64_t dims[2] = {32u, (uint64_t)n_img * N};
    const uint64_t str[1] = {64u};
    const uint32_t box[2] = {32u, 128u};
    if ((rc = encode_tmap_bf16(&maps.x, x, 2, dims, str, box, 64))) return rc;
  }
  {
    const uint64_t dims[2] = {32u, 256u};  // the q rows of the packed [768][32] projection
    const uint64_t str[1] = {64u};
    const uint32_t box[2] = {32u, 256u};
    if ((rc = encode_tmap_bf16(&maps.w, w_qkv, 2, dims, str, box, 64))) return rc;
  }
  {
    const uint64_t dims[2] = {256u, (uint64_t)n_img * 32};
    const uint64_t str[1] = {512u};
    const uint32_t box[2] = {64u, 32u};
    if ((rc = encode_tmap_bf16(&maps.g, gt_ws, 2, dims, str, box, 128))) return rc;
  }
  const int tiles_per_img = N / 128;
  const int n_tiles = n_img * tiles_per_img;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(sla_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSaSmem);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "sla_apply_tc cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const int grid = std::min(n_tiles, 2 * num_sms());
  // plain stream order (no programmatic dependent launch): the kernel's TMA reads gt_ws, written by the launch above
  sla_apply_tc_kernel<<<grid, kSaThreads, kSaSmem, st>>>(maps, reinterpret_cast<const bf16*>(x),
                                                        reinterpret_cast<bf16*>(out), tiles_per_img, n_tiles);
  return check_launch("sla_apply_tc");
}

}  // namespace vdn
