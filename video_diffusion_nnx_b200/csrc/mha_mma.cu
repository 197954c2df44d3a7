// Temporal attention core backward on warp-level tensor-core MMAs (modules.py:285-324 differentiated by
// jax.value_and_grad, trainer.py:361; sequences are the F <= 16 frames of one pixel, unet3d.py:86-96).
//
// One warp owns one (pixel, head): its whole problem is Q, K, V, dO of shape [F x 32] and fits the
// m16n8k16 fragments of a single warp, so everything between the global loads and the global stores lives
// in registers - no shared memory, no TMEM, no barriers:
//   S   = Q K^T          dP   = dO V^T         (rows = query tokens)
//   P^T, dS^T            8x8-tile transposes of the fragments (movmatrix), rows = key tokens
//   P = exp(S/sqrt(d) - lse),  D = rowsum(P dP),  dS = P (dP - D)/sqrt(d)
//   dQ = dS K            dK = dS^T Q           dV = P^T dO
// Operand access uses the same permuted-fragment trick as sla_mma.cu: with the contraction index over the 32
// head features permuted, lane (g, j) feeds the MMAs straight from the 16-byte chunk j of token rows g / g+8;
// operands that contract over TOKENS are gathered as 32-bit words of the token rows 2j, 2j+1, 2j+8, 2j+9 and
// transposed in registers with prmt. The tcgen05 version of this kernel (mha_tc.cu) pads the F x F problem
// to a 128 x 128 block-diagonal tile and is bound by its per-row global accesses; this one moves each byte
// once and is bound by HBM.
#include <algorithm>
#include <cstdlib>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// C[16 x 16] = A[16 x 32] B[16 x 32]^T with both operands given as the lane's 16-byte chunks of rows g / g+8
// (feature contraction, permuted k). c[t] = n-tile t (columns 8t .. 8t+7 = rows of B).
__device__ __forceinline__ void chunk_abt(const uint4& a_lo, const uint4& a_hi, const uint4& b_lo, const uint4& b_hi,
                                          float (&c)[2][4]) {
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  mma16816(c[0], a_lo.x, a_hi.x, a_lo.y, a_hi.y, b_lo.x, b_lo.y);
  mma16816(c[0], a_lo.z, a_hi.z, a_lo.w, a_hi.w, b_lo.z, b_lo.w);
  mma16816(c[1], a_lo.x, a_hi.x, a_lo.y, a_hi.y, b_hi.x, b_hi.y);
  mma16816(c[1], a_lo.z, a_hi.z, a_lo.w, a_hi.w, b_hi.z, b_hi.w);
}

// Words g and g+8 (features 2g,2g+1 | 2g+16,2g+17) of the head slice of token rows 2j, 2j+1, 2j+8, 2j+9.
struct TokWords {
  uint32_t w0[4], w1[4];
};
__device__ __forceinline__ void load_tokwords(TokWords& w, const bf16* base, long tok_stride, int F, int g, int j) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int tok = 2 * j + (c & 1) + 8 * (c >> 1);
    if (tok < F) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(base + tok * tok_stride);
      w.w0[c] = __ldg(row + g);
      w.w1[c] = __ldg(row + 8 + g);
    } else {
      w.w0[c] = 0u;
      w.w1[c] = 0u;
    }
  }
}
// transpose of an 8x8 bf16 tile held one 32-bit register per lane (row = lane/4, columns 2*(lane%4), +1)
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
// out[16 x 32] = A[16 x 16 tokens] * X[16 tokens x 32]: A as packed A-fragment registers {a0,a1,a2,a3}, X as token
// words. o[t][i]: n-tile t (features 2n | 2n+1 | 2n+16 | 2n+17), fragment element i.
__device__ __forceinline__ void afrag_times_tokens(const uint32_t (&af)[4], const TokWords& x, float (&o)[4][4]) {
  const uint32_t a0 = af[0], a1 = af[1], a2 = af[2], a3 = af[3];
  uint32_t b[4][2];
  b[0][0] = __byte_perm(x.w0[0], x.w0[1], 0x5410); b[0][1] = __byte_perm(x.w0[2], x.w0[3], 0x5410);
  b[1][0] = __byte_perm(x.w0[0], x.w0[1], 0x7632); b[1][1] = __byte_perm(x.w0[2], x.w0[3], 0x7632);
  b[2][0] = __byte_perm(x.w1[0], x.w1[1], 0x5410); b[2][1] = __byte_perm(x.w1[2], x.w1[3], 0x5410);
  b[3][0] = __byte_perm(x.w1[0], x.w1[1], 0x7632); b[3][1] = __byte_perm(x.w1[2], x.w1[3], 0x7632);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[t][i] = 0.f;
    mma16816(o[t], a0, a1, a2, a3, b[t][0], b[t][1]);
  }
}
// same with A given as C-fragments of two 8-column tiles (rows g / g+8)
__device__ __forceinline__ void frag_times_tokens(const float (&a)[2][4], const TokWords& x, float (&o)[4][4]) {
  const uint32_t af[4] = {pack_bf16x2(a[0][0], a[0][1]), pack_bf16x2(a[0][2], a[0][3]), pack_bf16x2(a[1][0], a[1][1]),
                          pack_bf16x2(a[1][2], a[1][3])};
  afrag_times_tokens(af, x, o);
}
// Stores the [16 x 32] result fragments of frag_times_tokens: per row two 8-byte pieces (features 4j..4j+3 and
// 4j+16..4j+19).
__device__ __forceinline__ void store_rows(const float (&o)[4][4], bf16* row_lo, bf16* row_hi, bool v_lo, bool v_hi,
                                           int j) {
  if (v_lo) {
    *reinterpret_cast<uint2*>(row_lo + 4 * j) = make_uint2(pack_bf16x2(o[0][0], o[1][0]), pack_bf16x2(o[0][1], o[1][1]));
    *reinterpret_cast<uint2*>(row_lo + 16 + 4 * j) = make_uint2(pack_bf16x2(o[2][0], o[3][0]), pack_bf16x2(o[2][1], o[3][1]));
  }
  if (v_hi) {
    *reinterpret_cast<uint2*>(row_hi + 4 * j) = make_uint2(pack_bf16x2(o[0][2], o[1][2]), pack_bf16x2(o[0][3], o[1][3]));
    *reinterpret_cast<uint2*>(row_hi + 16 + 4 * j) = make_uint2(pack_bf16x2(o[2][2], o[3][2]), pack_bf16x2(o[2][3], o[3][3]));
  }
}

// grid: blocks of 8 warps = the 8 heads of one pixel; a block walks pixels blockIdx.x, +gridDim.x, ...
__global__ void __launch_bounds__(256, 4) mha_temporal_mma_bwd_kernel(const bf16* __restrict__ qkv,
                                                                   const bf16* __restrict__ d_o,
                                                                   const float* __restrict__ lse,
                                                                   bf16* __restrict__ dqkv, int B, int F, long HW) {
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  const float scale = rsqrtf(32.f);
  const long n_pix = (long)B * HW;
  const bool v_lo = g < F, v_hi = g + 8 < F;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (long pix = blockIdx.x; pix < n_pix; pix += gridDim.x) {
    const long b = pix / HW, p = pix - b * HW;
    const long row0 = b * F * HW + p;          // token f lives at row0 + f*HW
    const long r_lo = row0 + (long)g * HW, r_hi = row0 + (long)(g + 8) * HW;
    // ---- 16-byte chunk operands (rows g / g+8) ----
    uint4 q_lo = zero4, q_hi = zero4, k_lo = zero4, k_hi = zero4, vv_lo = zero4, vv_hi = zero4, g_lo = zero4, g_hi = zero4;
    float L_lo = 0.f, L_hi = 0.f;
    if (v_lo) {
      const uint4* pr = reinterpret_cast<const uint4*>(qkv + r_lo * 768 + h * 32) + j;
      q_lo = __ldg(pr);
      k_lo = __ldg(pr + 32);
      vv_lo = __ldg(pr + 64);
      g_lo = __ldg(reinterpret_cast<const uint4*>(d_o + r_lo * 256 + h * 32) + j);
      L_lo = __ldg(lse + r_lo * 8 + h);
    }
    if (v_hi) {
      const uint4* pr = reinterpret_cast<const uint4*>(qkv + r_hi * 768 + h * 32) + j;
      q_hi = __ldg(pr);
      k_hi = __ldg(pr + 32);
      vv_hi = __ldg(pr + 64);
      g_hi = __ldg(reinterpret_cast<const uint4*>(d_o + r_hi * 256 + h * 32) + j);
      L_hi = __ldg(lse + r_hi * 8 + h);
    }
    // ---- operands that contract over TOKENS (K for dQ, Q for dK, dO for dV) are the 8x8-tile transposes of the chunk
    //      registers already loaded (movmatrix), as in the forward kernel's P V: 24 register shuffles instead of 24
    //      four-byte global loads, their address arithmetic and the byte permutes; n-tile t then holds the features of
    //      chunk register t, so the products land in the chunk layout and leave with 16-byte stores ----
    auto times_tokens = [&](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, const uint4& lo, const uint4& hi,
                            bf16* dst_lo, bf16* dst_hi) {
      const uint32_t bt[4][2] = {{movmatrix_trans(lo.x), movmatrix_trans(hi.x)}, {movmatrix_trans(lo.y), movmatrix_trans(hi.y)},
                                 {movmatrix_trans(lo.z), movmatrix_trans(hi.z)}, {movmatrix_trans(lo.w), movmatrix_trans(hi.w)}};
      uint32_t u_lo[4], u_hi[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(c, a0, a1, a2, a3, bt[t][0], bt[t][1]);
        u_lo[t] = pack_bf16x2(c[0], c[1]);
        u_hi[t] = pack_bf16x2(c[2], c[3]);
      }
      if (v_lo) reinterpret_cast<uint4*>(dst_lo)[j] = make_uint4(u_lo[0], u_lo[1], u_lo[2], u_lo[3]);
      if (v_hi) reinterpret_cast<uint4*>(dst_hi)[j] = make_uint4(u_hi[0], u_hi[1], u_hi[2], u_hi[3]);
    };

    // ---- query-major pass: P, dS (rows = query tokens g / g+8, columns = key tokens) ----
    float S[2][4], dP[2][4];
    chunk_abt(q_lo, q_hi, k_lo, k_hi, S);
    chunk_abt(g_lo, g_hi, vv_lo, vv_hi, dP);
    float D_lo = 0.f, D_hi = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool cv = 8 * t + 2 * j + i < F;  // key column valid
        S[t][i] = (cv && v_lo) ? __expf(S[t][i] * scale - L_lo) : 0.f;
        S[t][2 + i] = (cv && v_hi) ? __expf(S[t][2 + i] * scale - L_hi) : 0.f;
        D_lo = fmaf(S[t][i], dP[t][i], D_lo);
        D_hi = fmaf(S[t][2 + i], dP[t][2 + i], D_hi);
      }
    D_lo += __shfl_xor_sync(0xffffffffu, D_lo, 1);
    D_lo += __shfl_xor_sync(0xffffffffu, D_lo, 2);
    D_hi += __shfl_xor_sync(0xffffffffu, D_hi, 1);
    D_hi += __shfl_xor_sync(0xffffffffu, D_hi, 2);
    float dS[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        dS[t][i] = S[t][i] * (dP[t][i] - D_lo) * scale;
        dS[t][2 + i] = S[t][2 + i] * (dP[t][2 + i] - D_hi) * scale;
      }
    // dQ = dS K
    times_tokens(pack_bf16x2(dS[0][0], dS[0][1]), pack_bf16x2(dS[0][2], dS[0][3]), pack_bf16x2(dS[1][0], dS[1][1]),
                 pack_bf16x2(dS[1][2], dS[1][3]), k_lo, k_hi, dqkv + r_lo * 768 + h * 32, dqkv + r_hi * 768 + h * 32);

    // ---- key-major operands: P^T and dS^T are the 8x8-tile transposes of the query-major fragments
    //      (movmatrix), A fragment = {T(q0-7 x k0-7), T(q0-7 x k8-15), T(q8-15 x k0-7), T(q8-15 x k8-15)} ----
    uint32_t pT[4], dsT[4];
    pT[0] = movmatrix_trans(pack_bf16x2(S[0][0], S[0][1]));
    pT[1] = movmatrix_trans(pack_bf16x2(S[1][0], S[1][1]));
    pT[2] = movmatrix_trans(pack_bf16x2(S[0][2], S[0][3]));
    pT[3] = movmatrix_trans(pack_bf16x2(S[1][2], S[1][3]));
    dsT[0] = movmatrix_trans(pack_bf16x2(dS[0][0], dS[0][1]));
    dsT[1] = movmatrix_trans(pack_bf16x2(dS[1][0], dS[1][1]));
    dsT[2] = movmatrix_trans(pack_bf16x2(dS[0][2], dS[0][3]));
    dsT[3] = movmatrix_trans(pack_bf16x2(dS[1][2], dS[1][3]));
    // dK = dS^T Q, dV = P^T dO
    times_tokens(dsT[0], dsT[1], dsT[2], dsT[3], q_lo, q_hi, dqkv + r_lo * 768 + 256 + h * 32, dqkv + r_hi * 768 + 256 + h * 32);
    times_tokens(pT[0], pT[1], pT[2], pT[3], g_lo, g_hi, dqkv + r_lo * 768 + 512 + h * 32, dqkv + r_hi * 768 + 512 + h * 32);
  }
}

// ---------------------------------------------------------------------------------------
// Fused forward for C = 32: QKV projection + attention core of one (pixel, head) per warp, all in registers
// (reference: modules.py:285-323; same contract as vdn_mha_temporal_fused_fwd).
//   q|k|v = x W_h + b_h     [F x 32] each  (x rows as 16-byte chunk operands, weights as B fragments)
//   v^T                     8x8-tile transposes of the v registers (movmatrix): the B operand of P V
//   S = q k^T / sqrt(32), P = softmax(S), O = P V, lse = max + log(sum)
// The output-feature order of each projection is permuted (by choosing which weight row feeds which
// fragment column) so that a lane's accumulator registers ARE the 16-byte chunk j of token rows g / g+8:
// q/k/v feed the next MMA without any data movement, and qkv (training) / o are written with 16-byte stores.
// Weight fragments stay in registers across the pixel loop (48 registers for C = 32).
// ---------------------------------------------------------------------------------------
struct ProjW {
  uint32_t b[4][2][2];  // [n-tile][k-step][2]  B fragments of x W (output feature psi(t, g))
  float bias[4][2];     // bias of output features psi(t, 2j), psi(t, 2j+1)
};
// psi(t, c): feature held by fragment column c of n-tile t  ->  lane j ends up with features 8j .. 8j+7
__device__ __forceinline__ int psi(int t, int c) { return 8 * (c >> 1) + 4 * (t >> 1) + 2 * (t & 1) + (c & 1); }
__device__ __forceinline__ void load_projw(ProjW& w, const bf16* w_rows /*[32 feat][32 c]*/, const float* bias, int g,
                                           int j) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(w_rows + psi(t, g) * 32);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      w.b[t][s][0] = __ldg(row + 4 * j + 2 * s);      // channels 8j+4s, 8j+4s+1
      w.b[t][s][1] = __ldg(row + 4 * j + 2 * s + 1);  // channels 8j+4s+2, 8j+4s+3
    }
    w.bias[t][0] = __ldg(bias + psi(t, 2 * j));
    w.bias[t][1] = __ldg(bias + psi(t, 2 * j + 1));
  }
}
// Fragment sets live in shared memory as [24 words][32 lanes] (lane innermost: conflict-free LDS), written once
// per warp: keeping them in registers (68+ per thread) caps the kernel at 8 warps per SM, and its chains of
// dependent MMAs need more resident warps than that to hide their latency.
constexpr int kProjWords = 24;  // 16 fragment words + 8 bias floats
__device__ __forceinline__ void store_projw(uint32_t* sm, const ProjW& w, int lane) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      sm[((t * 2 + s) * 2 + 0) * 32 + lane] = w.b[t][s][0];
      sm[((t * 2 + s) * 2 + 1) * 32 + lane] = w.b[t][s][1];
    }
    sm[(16 + 2 * t) * 32 + lane] = __float_as_uint(w.bias[t][0]);
    sm[(16 + 2 * t + 1) * 32 + lane] = __float_as_uint(w.bias[t][1]);
  }
}
// lo/hi = bf16 chunk j of the projected rows g / g+8
__device__ __forceinline__ void project_chunks(const uint4& x_lo, const uint4& x_hi, const uint32_t* sm, int lane,
                                               uint4& lo, uint4& hi) {
  uint32_t l[4], h[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float b0 = __uint_as_float(sm[(16 + 2 * t) * 32 + lane]), b1 = __uint_as_float(sm[(16 + 2 * t + 1) * 32 + lane]);
    float c[4] = {b0, b1, b0, b1};
    mma16816(c, x_lo.x, x_hi.x, x_lo.y, x_hi.y, sm[((t * 2 + 0) * 2 + 0) * 32 + lane], sm[((t * 2 + 0) * 2 + 1) * 32 + lane]);
    mma16816(c, x_lo.z, x_hi.z, x_lo.w, x_hi.w, sm[((t * 2 + 1) * 2 + 0) * 32 + lane], sm[((t * 2 + 1) * 2 + 1) * 32 + lane]);
    l[t] = pack_bf16x2(c[0], c[1]);
    h[t] = pack_bf16x2(c[2], c[3]);
  }
  lo = make_uint4(l[0], l[1], l[2], l[3]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
}

template <bool kWriteQkv>
__global__ void __launch_bounds__(256, 2) mha_temporal_mma_fwd_kernel(const bf16* __restrict__ x,
                                                                   const bf16* __restrict__ w_hm,
                                                                   const float* __restrict__ bias_hm,
                                                                   bf16* __restrict__ o, bf16* __restrict__ qkv,
                                                                   float* __restrict__ lse, int B, int F, long HW) {
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  const float scale = rsqrtf(32.f);
  const long n_pix = (long)B * HW;
  const bool v_lo = g < F, v_hi = g + 8 < F;
  extern __shared__ uint32_t sm_w[];  // per head: q | k | v^T (| v) fragment sets
  constexpr int kHeadWords = 3 * kProjWords * 32;
  uint32_t* s_q = sm_w + h * kHeadWords;
  uint32_t* s_k = s_q + kProjWords * 32;
  uint32_t* s_v = s_k + kProjWords * 32;
  {
    ProjW w;
    load_projw(w, w_hm + (h * 96) * 32, bias_hm + h * 96, g, j);
    store_projw(s_q, w, lane);
    load_projw(w, w_hm + (h * 96 + 32) * 32, bias_hm + h * 96 + 32, g, j);
    store_projw(s_k, w, lane);
    load_projw(w, w_hm + (h * 96 + 64) * 32, bias_hm + h * 96 + 64, g, j);
    store_projw(s_v, w, lane);
  }
  __syncwarp();
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  // the x chunks of the NEXT pixel are fetched while the current one is processed (register double buffer)
  uint4 nx_lo = zero4, nx_hi = zero4;
  if (blockIdx.x < n_pix) {
    const long b = blockIdx.x / HW, p = blockIdx.x - b * HW;
    const long row0 = b * F * HW + p;
    if (v_lo) nx_lo = __ldg(reinterpret_cast<const uint4*>(x + (row0 + (long)g * HW) * 32) + j);
    if (v_hi) nx_hi = __ldg(reinterpret_cast<const uint4*>(x + (row0 + (long)(g + 8) * HW) * 32) + j);
  }
  for (long pix = blockIdx.x; pix < n_pix; pix += gridDim.x) {
    const long b = pix / HW, p = pix - b * HW;
    const long row0 = b * F * HW + p;
    const long r_lo = row0 + (long)g * HW, r_hi = row0 + (long)(g + 8) * HW;
    const uint4 x_lo = nx_lo, x_hi = nx_hi;
    {
      const long np = pix + gridDim.x;
      if (np < n_pix) {
        const long nb = np / HW, pp = np - nb * HW;
        const long nrow0 = nb * F * HW + pp;
        if (v_lo) nx_lo = __ldg(reinterpret_cast<const uint4*>(x + (nrow0 + (long)g * HW) * 32) + j);
        if (v_hi) nx_hi = __ldg(reinterpret_cast<const uint4*>(x + (nrow0 + (long)(g + 8) * HW) * 32) + j);
      }
    }
    uint4 q_lo, q_hi, k_lo, k_hi;
    project_chunks(x_lo, x_hi, s_q, lane, q_lo, q_hi);
    project_chunks(x_lo, x_hi, s_k, lane, k_lo, k_hi);
    uint4 vv_lo, vv_hi;
    project_chunks(x_lo, x_hi, s_v, lane, vv_lo, vv_hi);
    if (kWriteQkv) {  // training: the backward reads q | k | v
      if (v_lo) {
        uint4* pr = reinterpret_cast<uint4*>(qkv + r_lo * 768 + h * 32) + j;
        pr[0] = q_lo;
        pr[32] = k_lo;
        pr[64] = vv_lo;
      }
      if (v_hi) {
        uint4* pr = reinterpret_cast<uint4*>(qkv + r_hi * 768 + h * 32) + j;
        pr[0] = q_hi;
        pr[32] = k_hi;
        pr[64] = vv_hi;
      }
    }
    // B fragments of P V (contraction over tokens) = 8x8-tile transposes of the v chunk registers (movmatrix):
    // n-tile t holds the features psi(t, .) of the chunk layout, so O lands in the chunk layout as well
    uint32_t vb[4][2];
    vb[0][0] = movmatrix_trans(vv_lo.x); vb[0][1] = movmatrix_trans(vv_hi.x);
    vb[1][0] = movmatrix_trans(vv_lo.y); vb[1][1] = movmatrix_trans(vv_hi.y);
    vb[2][0] = movmatrix_trans(vv_lo.z); vb[2][1] = movmatrix_trans(vv_hi.z);
    vb[3][0] = movmatrix_trans(vv_lo.w); vb[3][1] = movmatrix_trans(vv_hi.w);
    // S = q k^T, softmax over the key tokens (columns 8t + 2j + i)
    float S[2][4];
    chunk_abt(q_lo, q_hi, k_lo, k_hi, S);
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool cv = 8 * t + 2 * j + i < F;
        S[t][i] = cv ? S[t][i] * scale : -INFINITY;
        S[t][2 + i] = cv ? S[t][2 + i] * scale : -INFINITY;
        m_lo = fmaxf(m_lo, S[t][i]);
        m_hi = fmaxf(m_hi, S[t][2 + i]);
      }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        S[t][i] = __expf(S[t][i] - m_lo);
        S[t][2 + i] = __expf(S[t][2 + i] - m_hi);
        l_lo += S[t][i];
        l_hi += S[t][2 + i];
      }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    // O = P V: A = P fragments (keys 2j,2j+1 | 8+2j,9+2j), B = transposed v tiles
    const uint32_t a0 = pack_bf16x2(S[0][0], S[0][1]), a1 = pack_bf16x2(S[0][2], S[0][3]);
    const uint32_t a2 = pack_bf16x2(S[1][0], S[1][1]), a3 = pack_bf16x2(S[1][2], S[1][3]);
    float o_lo[8], o_hi[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(c, a0, a1, a2, a3, vb[t][0], vb[t][1]);
      o_lo[2 * t] = c[0];
      o_lo[2 * t + 1] = c[1];
      o_hi[2 * t] = c[2];
      o_hi[2 * t + 1] = c[3];
    }
    const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
    if (v_lo) {
      uint4 u4;
      u4.x = pack_bf16x2(o_lo[0] * inv_lo, o_lo[1] * inv_lo);
      u4.y = pack_bf16x2(o_lo[2] * inv_lo, o_lo[3] * inv_lo);
      u4.z = pack_bf16x2(o_lo[4] * inv_lo, o_lo[5] * inv_lo);
      u4.w = pack_bf16x2(o_lo[6] * inv_lo, o_lo[7] * inv_lo);
      reinterpret_cast<uint4*>(o + r_lo * 256 + h * 32)[j] = u4;
      if (lse && j == 0) lse[r_lo * 8 + h] = m_lo + __logf(l_lo);
    }
    if (v_hi) {
      uint4 u4;
      u4.x = pack_bf16x2(o_hi[0] * inv_hi, o_hi[1] * inv_hi);
      u4.y = pack_bf16x2(o_hi[2] * inv_hi, o_hi[3] * inv_hi);
      u4.z = pack_bf16x2(o_hi[4] * inv_hi, o_hi[5] * inv_hi);
      u4.w = pack_bf16x2(o_hi[6] * inv_hi, o_hi[7] * inv_hi);
      reinterpret_cast<uint4*>(o + r_hi * 256 + h * 32)[j] = u4;
      if (lse && j == 0) lse[r_hi * 8 + h] = m_hi + __logf(l_hi);
    }
  }
}

int mha_temporal_mma_fwd_launch(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv, float* lse,
                                int B, int F, int H, int W, cudaStream_t st) {
  const long HW = (long)H * W;
  const long n_pix = (long)B * HW;
  const int grid = (int)std::min<long>(n_pix, 148L * tune_int("VDN_MHA_FWD_CTAS", 16));
  const size_t smem_t = (size_t)8 * 3 * kProjWords * 32 * 4, smem_i = smem_t;
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(mha_temporal_mma_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
    cudaFuncSetAttribute(mha_temporal_mma_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i);
    cfg = true;
  }
  cudaError_t le =
      qkv ? launch_pdl(mha_temporal_mma_fwd_kernel<true>, dim3(grid), dim3(256), smem_t, st, 1,
                       reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(w_hm), bias_hm,
                       reinterpret_cast<bf16*>(o), reinterpret_cast<bf16*>(qkv), lse, B, F, HW)
          : launch_pdl(mha_temporal_mma_fwd_kernel<false>, dim3(grid), dim3(256), smem_i, st, 1,
                       reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(w_hm), bias_hm,
                       reinterpret_cast<bf16*>(o), reinterpret_cast<bf16*>(qkv), lse, B, F, HW);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_temporal_mma_fwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_temporal_mma_fwd");
}

// ---------------------------------------------------------------------------------------
// Folded temporal attention block for inference engines, C = 32: out = x + out_proj(MHA(x)) in one kernel.
// Without a backward pass nothing forces q, k, v, o into memory, and the algebra folds:
//   S_ij = q_i k_j / sqrt(d) = (x_i A_h + u_h) . x_j + (terms constant in j: softmax drops them)
//          A_h = W_q,h W_k,h^T / sqrt(d) [C x C],  u_h = b_q,h W_k,h^T / sqrt(d)
//   out  = sum_h softmax(S_h) v_h W_o,h + b_o = sum_h (P_h x) M_h + b',  M_h = W_v,h W_o,h,  b' = sum_h b_v,h W_o,h + b_o
// so a (pixel, head) costs 24 MMAs (y = xA+u: 8, S: 4, Z = P x: 4, Z M_h: 8) instead of 32 + a separate out-projection
// GEMM, and neither o nor the out-projection input ever exist. The 8 heads of a pixel are the 8 warps of a block and
// are summed through shared memory. The folded matrices are rebuilt by mha_fold_pack_kernel after every weight
// update (a few hundred kFLOP).
// ---------------------------------------------------------------------------------------
__global__ void mha_fold_pack_kernel(const float* __restrict__ w_qkv /*[32][768]*/, const float* __restrict__ b_qkv,
                                     const float* __restrict__ w_out /*[256][32]*/, const float* __restrict__ b_out,
                                     bf16* __restrict__ fa /*[8][32 n][32 c]*/, float* __restrict__ fu /*[8][32]*/,
                                     bf16* __restrict__ fm /*[8][32 c][32 c']*/, float* __restrict__ fb /*[32]*/) {
  const int h = blockIdx.x;
  const float scale = rsqrtf(32.f);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int n = i >> 5, c = i & 31;
    // fa[h][n][c] = scale * A[c][n] = scale * sum_d W_q[c][h,d] W_k[n][h,d]
    float a = 0.f, m = 0.f;
    for (int d = 0; d < 32; ++d) {
      a = fmaf(w_qkv[c * 768 + h * 32 + d], w_qkv[n * 768 + 256 + h * 32 + d], a);
      // fm[h][n][c] = M[c][n] = sum_d W_v[c][h,d] W_o[h*32+d][n]
      m = fmaf(w_qkv[c * 768 + 512 + h * 32 + d], w_out[(h * 32 + d) * 32 + n], m);
    }
    fa[(h * 32 + n) * 32 + c] = __float2bfloat16(a * scale);
    fm[(h * 32 + n) * 32 + c] = __float2bfloat16(m);
  }
  for (int n = threadIdx.x; n < 32; n += blockDim.x) {
    float u = 0.f;
    for (int d = 0; d < 32; ++d) u = fmaf(b_qkv ? b_qkv[h * 32 + d] : 0.f, w_qkv[n * 768 + 256 + h * 32 + d], u);
    fu[h * 32 + n] = u * scale;
    if (h == 0) {  // b'[n] = b_o[n] + sum_{h,d} b_v[h,d] W_o[h*32+d][n]
      float bsum = b_out ? b_out[n] : 0.f;
      for (int r = 0; r < 256; ++r) bsum = fmaf(b_qkv ? b_qkv[512 + r] : 0.f, w_out[r * 32 + n], bsum);
      fb[n] = bsum;
    }
  }
}

__global__ void __launch_bounds__(256, 2) mha_temporal_folded_fwd_kernel(const bf16* __restrict__ x,
                                                                       const bf16* __restrict__ fa,
                                                                       const float* __restrict__ fu,
                                                                       const bf16* __restrict__ fm,
                                                                       const float* __restrict__ fb,
                                                                       bf16* __restrict__ out, int B, int F, long HW) {
  __shared__ float s_part[8][16][33];
  extern __shared__ uint32_t sm_w[];  // per head: y-projection fragment set (24 words) | M_h fragment set (16 words)
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  const long n_pix = (long)B * HW;
  const bool v_lo = g < F, v_hi = g + 8 < F;
  uint32_t* s_y = sm_w + h * (kProjWords + 16) * 32;
  uint32_t* s_m = s_y + kProjWords * 32;
  {
    ProjW w;
    load_projw(w, fa + (h * 32) * 32, fu + h * 32, g, j);  // psi-permuted outputs: y lands in chunk layout
    store_projw(s_y, w, lane);
    // B[k = c'][n = c] = M_h[c'][c] = fm[h][c][c'], sigma columns (result in chunk layout), chunk-order k
#pragma unroll
    for (int sx = 0; sx < 2; ++sx)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int n = 8 * (g >> 1) + 2 * t + (g & 1);
        const uint32_t* row = reinterpret_cast<const uint32_t*>(fm + (h * 32 + n) * 32);
        s_m[((sx * 4 + t) * 2 + 0) * 32 + lane] = __ldg(row + 4 * j + 2 * sx);
        s_m[((sx * 4 + t) * 2 + 1) * 32 + lane] = __ldg(row + 4 * j + 2 * sx + 1);
      }
  }
  __syncwarp();
  const float bias0 = fb[2 * (threadIdx.x & 15)], bias1 = fb[2 * (threadIdx.x & 15) + 1];
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (long pix = blockIdx.x; pix < n_pix; pix += gridDim.x) {  // uniform trip count for the whole block
    const long b = pix / HW, p = pix - b * HW;
    const long row0 = b * F * HW + p;
    const long r_lo = row0 + (long)g * HW, r_hi = row0 + (long)(g + 8) * HW;
    const uint4 x_lo = v_lo ? __ldg(reinterpret_cast<const uint4*>(x + r_lo * 32) + j) : zero4;
    const uint4 x_hi = v_hi ? __ldg(reinterpret_cast<const uint4*>(x + r_hi * 32) + j) : zero4;
    uint4 y_lo, y_hi;
    project_chunks(x_lo, x_hi, s_y, lane, y_lo, y_hi);  // y = x A_h + u_h (already scaled by 1/sqrt(d))
    float S[2][4];
    chunk_abt(y_lo, y_hi, x_lo, x_hi, S);               // S'_ij = y_i . x_j
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool cv = 8 * t + 2 * j + i < F;
        S[t][i] = cv ? S[t][i] : -INFINITY;
        S[t][2 + i] = cv ? S[t][2 + i] : -INFINITY;
        m_lo = fmaxf(m_lo, S[t][i]);
        m_hi = fmaxf(m_hi, S[t][2 + i]);
      }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        S[t][i] = __expf(S[t][i] - m_lo);
        S[t][2 + i] = __expf(S[t][2 + i] - m_hi);
        l_lo += S[t][i];
        l_hi += S[t][2 + i];
      }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    // Z = P x: B fragments = transposed x tiles; Z lands in chunk layout (channels 8j..8j+7)
    const uint32_t a0 = pack_bf16x2(S[0][0], S[0][1]), a1 = pack_bf16x2(S[0][2], S[0][3]);
    const uint32_t a2 = pack_bf16x2(S[1][0], S[1][1]), a3 = pack_bf16x2(S[1][2], S[1][3]);
    const uint32_t xb[4][2] = {{movmatrix_trans(x_lo.x), movmatrix_trans(x_hi.x)},
                               {movmatrix_trans(x_lo.y), movmatrix_trans(x_hi.y)},
                               {movmatrix_trans(x_lo.z), movmatrix_trans(x_hi.z)},
                               {movmatrix_trans(x_lo.w), movmatrix_trans(x_hi.w)}};
    const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
    uint32_t z_lo[4], z_hi[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(c, a0, a1, a2, a3, xb[t][0], xb[t][1]);
      z_lo[t] = pack_bf16x2(c[0] * inv_lo, c[1] * inv_lo);
      z_hi[t] = pack_bf16x2(c[2] * inv_hi, c[3] * inv_hi);
    }
    // out_h = Z M_h (chunk layout), summed over the block's 8 heads through shared memory
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(d, z_lo[0], z_hi[0], z_lo[1], z_hi[1], s_m[((0 * 4 + t) * 2 + 0) * 32 + lane], s_m[((0 * 4 + t) * 2 + 1) * 32 + lane]);
      mma16816(d, z_lo[2], z_hi[2], z_lo[3], z_hi[3], s_m[((1 * 4 + t) * 2 + 0) * 32 + lane], s_m[((1 * 4 + t) * 2 + 1) * 32 + lane]);
      s_part[h][g][8 * j + 2 * t] = d[0];
      s_part[h][g][8 * j + 2 * t + 1] = d[1];
      s_part[h][g + 8][8 * j + 2 * t] = d[2];
      s_part[h][g + 8][8 * j + 2 * t + 1] = d[3];
    }
    __syncthreads();
    {
      const int r = threadIdx.x >> 4, c = (threadIdx.x & 15) * 2;  // token r, channels c, c+1
      if (r < F) {
        float v0 = bias0, v1 = bias1;
#pragma unroll
        for (int hh = 0; hh < 8; ++hh) {
          v0 += s_part[hh][r][c];
          v1 += s_part[hh][r][c + 1];
        }
        const long row = row0 + (long)r * HW;
        const float2 xr = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + row * 32 + c));
        *reinterpret_cast<uint32_t*>(out + row * 32 + c) = pack_bf16x2(v0 + xr.x, v1 + xr.y);
      }
    }
    __syncthreads();
  }
}

int mha_fold_pack_launch(const float* w_qkv, const float* b_qkv, const float* w_out, const float* b_out, void* fa,
                         float* fu, void* fm, float* fb, cudaStream_t st) {
  mha_fold_pack_kernel<<<8, 256, 0, st>>>(w_qkv, b_qkv, w_out, b_out, reinterpret_cast<bf16*>(fa), fu,
                                          reinterpret_cast<bf16*>(fm), fb);
  return check_launch("mha_fold_pack");
}

int mha_temporal_folded_fwd_launch(const void* x, const void* fa, const float* fu, const void* fm, const float* fb,
                                   void* out, int B, int F, int H, int W, cudaStream_t st) {
  const long HW = (long)H * W;
  const long n_pix = (long)B * HW;
  const int grid = (int)std::min<long>(n_pix, 148L * 16);
  const size_t smem = (size_t)8 * (kProjWords + 16) * 32 * 4;
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(mha_temporal_folded_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  cudaError_t le = launch_pdl(mha_temporal_folded_fwd_kernel, dim3(grid), dim3(256), smem, st, 1,
                              reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(fa), fu,
                              reinterpret_cast<const bf16*>(fm), fb, reinterpret_cast<bf16*>(out), B, F, HW);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_temporal_folded_fwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_temporal_folded_fwd");
}

// ---------------------------------------------------------------------------------------
// Attention core forward on a materialised q|k|v tensor (any C: the projection ran as a tap-GEMM at full tensor-core
// rate): one warp per (pixel, head), everything in registers, every byte moved once with 16-byte accesses.
//   S = q k^T / sqrt(32), P = softmax(S), O = P V, lse = max + log(sum)     (modules.py:296-323)
// Same fragment scheme as the fused C = 32 kernel above: q, k, v arrive as the lane's 16-byte chunk j of token rows
// g / g+8; the B operand of P V is the movmatrix transpose of the v chunk registers.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) mha_temporal_mma_core_fwd_kernel(const bf16* __restrict__ qkv,
                                                                           bf16* __restrict__ o,
                                                                           float* __restrict__ lse, int B, int F,
                                                                           long HW) {
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  const float scale = rsqrtf(32.f);
  const long n_pix = (long)B * HW;
  const bool v_lo = g < F, v_hi = g + 8 < F;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (long pix = blockIdx.x; pix < n_pix; pix += gridDim.x) {
    const long b = pix / HW, p = pix - b * HW;
    const long row0 = b * F * HW + p;  // token f lives at row0 + f*HW
    const long r_lo = row0 + (long)g * HW, r_hi = row0 + (long)(g + 8) * HW;
    uint4 q_lo = zero4, q_hi = zero4, k_lo = zero4, k_hi = zero4, vv_lo = zero4, vv_hi = zero4;
    if (v_lo) {
      const uint4* pr = reinterpret_cast<const uint4*>(qkv + r_lo * 768 + h * 32) + j;
      q_lo = __ldg(pr);
      k_lo = __ldg(pr + 32);
      vv_lo = __ldg(pr + 64);
    }
    if (v_hi) {
      const uint4* pr = reinterpret_cast<const uint4*>(qkv + r_hi * 768 + h * 32) + j;
      q_hi = __ldg(pr);
      k_hi = __ldg(pr + 32);
      vv_hi = __ldg(pr + 64);
    }
    uint32_t vb[4][2];
    vb[0][0] = movmatrix_trans(vv_lo.x); vb[0][1] = movmatrix_trans(vv_hi.x);
    vb[1][0] = movmatrix_trans(vv_lo.y); vb[1][1] = movmatrix_trans(vv_hi.y);
    vb[2][0] = movmatrix_trans(vv_lo.z); vb[2][1] = movmatrix_trans(vv_hi.z);
    vb[3][0] = movmatrix_trans(vv_lo.w); vb[3][1] = movmatrix_trans(vv_hi.w);
    float S[2][4];
    chunk_abt(q_lo, q_hi, k_lo, k_hi, S);
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool cv = 8 * t + 2 * j + i < F;
        S[t][i] = cv ? S[t][i] * scale : -INFINITY;
        S[t][2 + i] = cv ? S[t][2 + i] * scale : -INFINITY;
        m_lo = fmaxf(m_lo, S[t][i]);
        m_hi = fmaxf(m_hi, S[t][2 + i]);
      }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    float l_lo = 0.f, l_hi = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        S[t][i] = __expf(S[t][i] - m_lo);
        S[t][2 + i] = __expf(S[t][2 + i] - m_hi);
        l_lo += S[t][i];
        l_hi += S[t][2 + i];
      }
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const uint32_t a0 = pack_bf16x2(S[0][0], S[0][1]), a1 = pack_bf16x2(S[0][2], S[0][3]);
    const uint32_t a2 = pack_bf16x2(S[1][0], S[1][1]), a3 = pack_bf16x2(S[1][2], S[1][3]);
    float o_lo[8], o_hi[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(c, a0, a1, a2, a3, vb[t][0], vb[t][1]);
      o_lo[2 * t] = c[0];
      o_lo[2 * t + 1] = c[1];
      o_hi[2 * t] = c[2];
      o_hi[2 * t + 1] = c[3];
    }
    const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
    if (v_lo) {
      uint4 u4;
      u4.x = pack_bf16x2(o_lo[0] * inv_lo, o_lo[1] * inv_lo);
      u4.y = pack_bf16x2(o_lo[2] * inv_lo, o_lo[3] * inv_lo);
      u4.z = pack_bf16x2(o_lo[4] * inv_lo, o_lo[5] * inv_lo);
      u4.w = pack_bf16x2(o_lo[6] * inv_lo, o_lo[7] * inv_lo);
      reinterpret_cast<uint4*>(o + r_lo * 256 + h * 32)[j] = u4;
      if (lse && j == 0) lse[r_lo * 8 + h] = m_lo + __logf(l_lo);
    }
    if (v_hi) {
      uint4 u4;
      u4.x = pack_bf16x2(o_hi[0] * inv_hi, o_hi[1] * inv_hi);
      u4.y = pack_bf16x2(o_hi[2] * inv_hi, o_hi[3] * inv_hi);
      u4.z = pack_bf16x2(o_hi[4] * inv_hi, o_hi[5] * inv_hi);
      u4.w = pack_bf16x2(o_hi[6] * inv_hi, o_hi[7] * inv_hi);
      reinterpret_cast<uint4*>(o + r_hi * 256 + h * 32)[j] = u4;
      if (lse && j == 0) lse[r_hi * 8 + h] = m_hi + __logf(l_hi);
    }
  }
}

int mha_temporal_mma_core_fwd_launch(const void* qkv, void* o, float* lse, int B, int F, int H, int W, cudaStream_t st) {
  const long HW = (long)H * W;
  const long n_pix = (long)B * HW;
  const int grid = (int)std::min<long>(n_pix, 148L * 24);
  cudaError_t le = launch_pdl(mha_temporal_mma_core_fwd_kernel, dim3(grid), dim3(256), (size_t)0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(o), lse, B, F, HW);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_temporal_mma_core_fwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_temporal_mma_core_fwd");
}

int mha_temporal_mma_bwd_launch(const void* qkv, const void* d_o, const float* lse, void* dqkv, int B, int F, int H,
                                int W, cudaStream_t st) {
  const long HW = (long)H * W;
  const long n_pix = (long)B * HW;
  const int grid = (int)std::min<long>(n_pix, 148L * tune_int("VDN_MHA_BWD_CTAS", 4));  // = the four resident CTAs per SM (64 registers)
  cudaError_t le = launch_pdl(mha_temporal_mma_bwd_kernel, dim3(grid), dim3(256), (size_t)0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(d_o), lse,
                              reinterpret_cast<bf16*>(dqkv), B, F, HW);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_temporal_mma_bwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_temporal_mma_bwd");
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_mha_temporal_core_fwd(const void* qkv, void* o, float* lse, int B, int F, int H, int W, void* stream) {
  VDN_REQUIRE(qkv && o && B > 0 && H > 0 && W > 0, VDN_E_SHAPE, "mha_temporal_core_fwd: bad args");
  VDN_REQUIRE(F >= 1 && F <= 16, VDN_E_SHAPE, "mha_temporal_core_fwd: F=%d unsupported (F <= 16)", F);
  return mha_temporal_mma_core_fwd_launch(qkv, o, lse, B, F, H, W, reinterpret_cast<cudaStream_t>(stream));
}

// Folded inference-only temporal attention block (C = 32): see mha_temporal_folded_fwd_kernel.
// fold_pack: w_qkv fp32 [32][768] (q|k|v), b_qkv fp32 [768] or NULL, w_out fp32 [256][32], b_out fp32 [32] or NULL ->
//            fa bf16 [8][32][32], fu fp32 [8][32], fm bf16 [8][32][32], fb fp32 [32].
extern "C" int vdn_mha_fold_pack(const float* w_qkv, const float* b_qkv, const float* w_out, const float* b_out,
                                 void* fa, float* fu, void* fm, float* fb, void* stream) {
  VDN_REQUIRE(w_qkv && w_out && fa && fu && fm && fb, VDN_E_SHAPE, "mha_fold_pack: bad args");
  return mha_fold_pack_launch(w_qkv, b_qkv, w_out, b_out, fa, fu, fm, fb, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int vdn_mha_temporal_folded_fwd(const void* x, const void* fa, const float* fu, const void* fm,
                                           const float* fb, void* out, int B, int F, int H, int W, int C,
                                           void* stream) {
  VDN_REQUIRE(x && fa && fu && fm && fb && out && B > 0 && H > 0 && W > 0, VDN_E_SHAPE, "mha_temporal_folded_fwd: bad args");
  VDN_REQUIRE(C == 32 && F >= 1 && F <= 16, VDN_E_SHAPE, "mha_temporal_folded_fwd: C=%d F=%d unsupported (C == 32, F <= 16)", C, F);
  // the shared-weight GEMMs on tcgen05 (mha_folded_tc.cu); VDN_MHA_FOLDED_MMA=1 keeps the all-mma.sync kernel
  if (mha_folded_tc_applicable(F, H * W) && !tune_on("VDN_MHA_FOLDED_MMA"))
    return mha_folded_tc_launch(x, fa, fu, fm, fb, out, B, F, H, W, reinterpret_cast<cudaStream_t>(stream));
  return mha_temporal_folded_fwd_launch(x, fa, fu, fm, fb, out, B, F, H, W, reinterpret_cast<cudaStream_t>(stream));
}
