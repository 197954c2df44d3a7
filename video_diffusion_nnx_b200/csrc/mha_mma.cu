// Temporal attention core backward on warp-level tensor-core MMAs (modules.py:285-324 differentiated by
// jax.value_and_grad, trainer.py:361; sequences are the F <= 16 frames of one pixel, unet3d.py:86-96).
//
// One warp owns one (pixel, head): its whole problem is Q, K, V, dO of shape [F x 32] and fits the
// m16n8k16 fragments of a single warp, so everything between the global loads and the global stores lives
// in registers - no shared memory, no TMEM, no barriers:
//   S   = Q K^T          dP   = dO V^T         (rows = query tokens)
//   S^T = K Q^T          dP^T = V dO^T         (rows = key tokens; recomputed instead of transposed)
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
// out[16 x 32] = A[16 x 16 tokens] * X[16 tokens x 32]: A as C-fragments of two 8-column tiles (rows g / g+8),
// X as token words. o[t][i]: n-tile t (features 2n | 2n+1 | 2n+16 | 2n+17), fragment element i.
__device__ __forceinline__ void frag_times_tokens(const float (&a)[2][4], const TokWords& x, float (&o)[4][4]) {
  const uint32_t a0 = pack_bf16x2(a[0][0], a[0][1]), a1 = pack_bf16x2(a[0][2], a[0][3]);
  const uint32_t a2 = pack_bf16x2(a[1][0], a[1][1]), a3 = pack_bf16x2(a[1][2], a[1][3]);
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
__global__ void __launch_bounds__(256) mha_temporal_mma_bwd_kernel(const bf16* __restrict__ qkv,
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
    const bf16* qb = qkv + row0 * 768 + h * 32;
    const bf16* gb = d_o + row0 * 256 + h * 32;
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
    // ---- token-word operands (contraction over tokens) ----
    TokWords wq, wk, wg;
    load_tokwords(wq, qb, HW * 768, F, g, j);
    load_tokwords(wk, qb + 256, HW * 768, F, g, j);
    load_tokwords(wg, gb, HW * 256, F, g, j);

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
    float o[4][4];
    frag_times_tokens(dS, wk, o);  // dQ = dS K
    store_rows(o, dqkv + r_lo * 768 + h * 32, dqkv + r_hi * 768 + h * 32, v_lo, v_hi, j);

    // ---- key-major pass: P^T, dS^T (rows = key tokens g / g+8, columns = query tokens) ----
    float ST[2][4], dPT[2][4];
    chunk_abt(k_lo, k_hi, q_lo, q_hi, ST);
    chunk_abt(vv_lo, vv_hi, g_lo, g_hi, dPT);
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int qc = 2 * j + i;  // query column within the tile; its L and D sit in the lanes of row qc
        const float Lc = __shfl_sync(0xffffffffu, t == 0 ? L_lo : L_hi, 4 * qc);
        const float Dc = __shfl_sync(0xffffffffu, t == 0 ? D_lo : D_hi, 4 * qc);
        const bool cv = 8 * t + qc < F;
        const float p_lo = (cv && v_lo) ? __expf(ST[t][i] * scale - Lc) : 0.f;
        const float p_hi = (cv && v_hi) ? __expf(ST[t][2 + i] * scale - Lc) : 0.f;
        ST[t][i] = p_lo;
        ST[t][2 + i] = p_hi;
        dPT[t][i] = p_lo * (dPT[t][i] - Dc) * scale;        // dS^T
        dPT[t][2 + i] = p_hi * (dPT[t][2 + i] - Dc) * scale;
      }
    frag_times_tokens(dPT, wq, o);  // dK = dS^T Q
    store_rows(o, dqkv + r_lo * 768 + 256 + h * 32, dqkv + r_hi * 768 + 256 + h * 32, v_lo, v_hi, j);
    frag_times_tokens(ST, wg, o);   // dV = P^T dO
    store_rows(o, dqkv + r_lo * 768 + 512 + h * 32, dqkv + r_hi * 768 + 512 + h * 32, v_lo, v_hi, j);
  }
}

int mha_temporal_mma_bwd_launch(const void* qkv, const void* d_o, const float* lse, void* dqkv, int B, int F, int H,
                                int W, cudaStream_t st) {
  const long HW = (long)H * W;
  const long n_pix = (long)B * HW;
  const int grid = (int)std::min<long>(n_pix, 148L * 16);
  cudaError_t le = launch_pdl(mha_temporal_mma_bwd_kernel, dim3(grid), dim3(256), (size_t)0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(d_o), lse,
                              reinterpret_cast<bf16*>(dqkv), B, F, HW);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_temporal_mma_bwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_temporal_mma_bwd");
}

}  // namespace vdn
