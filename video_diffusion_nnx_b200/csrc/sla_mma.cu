// SpatialLinearAttention per-token products on warp-level tensor-core MMAs (modules.py:105-123).
//
// The token-side work of linear attention is, per (token, head), a handful of 32-vector x (32x32) products
// against the per-frame context matrices:
//   forward   tok  = q~ ctx                      q~ = softmax_d(q)
//   backward  dq~  = dtok ctx^T ,  dq = q~ (dq~ - <q~, dq~>)
//             dk~  = v dctx^T   ,  dk = k~ (dk~ - r),   r[d] = sum_e ctx[d][e] dctx[d][e],  k~ = exp(k - m) / S
//             dv   = k~ dctx
// On CUDA cores these are 1024 FMAs per (token, head) and instruction bound (~35 % of the FMA peak); here a
// warp owns 16 tokens of one head and runs them as m16n8k16 bf16 MMAs with fp32 accumulation, which leaves
// the kernels bound by the qkv / dqkv HBM traffic. The matrices are tiny (32x32) and coupled to row-wise
// softmax algebra, so register fragments (mma.sync) are the right granularity: a tcgen05 formulation needs a
// TMEM round trip per product and block-diagonal padding to reach M = 128.
//
// Fragment trick: the contraction index of A and B (and the output column index) may be permuted freely as
// long as both operands agree. The permutations below make lane (g, j) of the warp own exactly the 16-byte
// chunk j (features 8j..8j+7) of token rows g and g+8 - for the A operand AND for the result - so operands go
// from global memory to MMA registers and results back to global memory with plain 16-byte accesses, no
// shared-memory staging and no shuffles:
//   k-step s, A regs (a0,a2) = features (8j+4s, 8j+4s+1), (8j+4s+2, 8j+4s+3) of the chunk
//   n-tile t, column c of the tile = output feature 8*(c/2) + 2t + (c%2)
#include <algorithm>
#include <cstdlib>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kSmHeads = 8;
constexpr int kSmDh = 32;
constexpr int kSmHD = 256;
constexpr int kSmQKV = 768;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// B fragments (2 k-steps x 4 n-tiles x 2 regs) of the 32x32 matrix Bm[k][n] = M[k*ks + n*ns] (fp32, shared memory)
// under the permuted index maps described above.
struct BFrag {
  uint32_t r[2][4][2];
};
__device__ __forceinline__ void load_bfrag(BFrag& f, const float* M, int ks, int ns, int g, int j) {
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int n = 8 * (g >> 1) + 2 * t + (g & 1);
      const int k0 = 8 * j + 4 * s;
      f.r[s][t][0] = pack_bf16x2(M[k0 * ks + n * ns], M[(k0 + 1) * ks + n * ns]);
      f.r[s][t][1] = pack_bf16x2(M[(k0 + 2) * ks + n * ns], M[(k0 + 3) * ks + n * ns]);
    }
}

// out(row g | g+8)[8 features of chunk j] = A(16 tokens x 32) * B : lo/hi = the lane's chunk of rows g / g+8
__device__ __forceinline__ void chunk_matmul(const uint4& lo, const uint4& hi, const BFrag& b, float (&olo)[8],
                                             float (&ohi)[8]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16_16816(d, lo.x, hi.x, lo.y, hi.y, b.r[0][t][0], b.r[0][t][1]);
    mma_bf16_16816(d, lo.z, hi.z, lo.w, hi.w, b.r[1][t][0], b.r[1][t][1]);
    olo[2 * t] = d[0];
    olo[2 * t + 1] = d[1];
    ohi[2 * t] = d[2];
    ohi[2 * t + 1] = d[3];
  }
}

__device__ __forceinline__ void unpack_chunk(const uint4& q, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(q.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(q.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(q.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(q.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ uint4 pack_chunk(const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
// softmax over the 32 features of a token row held as 4 chunks by the 4 lanes of a quad (in place)
__device__ __forceinline__ void quad_softmax(float (&v)[8]) {
  float mx = v[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[i]);
  mx = quad_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = __expf(v[i] - mx);
    s += v[i];
  }
  const float inv = 1.f / quad_sum(s);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= inv;
}

// ---------------------------------------------------------------------------------------
// forward: tok[n, h*32 + e] = sum_d softmax_D(q[n,h,:])[d] * ctx[h][d][e]
// grid (chunks, n_img), 256 threads = 8 warps = 8 heads; a warp walks 16-token groups of its frame.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sla_apply_mma_kernel(const bf16* __restrict__ qkv, const float* __restrict__ ctx,
                                                            bf16* __restrict__ out, int N) {
  __shared__ float sctx[kSmHeads * 1024];  // [8][32][32]
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < kSmHeads * 1024; i += blockDim.x) sctx[i] = ctx[(long)img * kSmHeads * 1024 + i];
  __syncthreads();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  BFrag bc;
  load_bfrag(bc, sctx + h * 1024, 32, 1, g, j);  // B[k = d][n = e] = ctx[d][e]
  const int n_groups = (N + 15) / 16;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int n_lo = grp * 16 + g, n_hi = n_lo + 8;
    const bool v_lo = n_lo < N, v_hi = n_hi < N;
    const long r_lo = (long)img * N + (v_lo ? n_lo : N - 1), r_hi = (long)img * N + (v_hi ? n_hi : N - 1);
    const uint4 q_lo = __ldg(reinterpret_cast<const uint4*>(qkv + r_lo * kSmQKV + h * kSmDh) + j);
    const uint4 q_hi = __ldg(reinterpret_cast<const uint4*>(qkv + r_hi * kSmQKV + h * kSmDh) + j);
    float a_lo[8], a_hi[8], o_lo[8], o_hi[8];
    unpack_chunk(q_lo, a_lo);
    unpack_chunk(q_hi, a_hi);
    quad_softmax(a_lo);
    quad_softmax(a_hi);
    chunk_matmul(pack_chunk(a_lo), pack_chunk(a_hi), bc, o_lo, o_hi);
    if (v_lo) reinterpret_cast<uint4*>(out + r_lo * kSmHD + h * kSmDh)[j] = pack_chunk(o_lo);
    if (v_hi) reinterpret_cast<uint4*>(out + r_hi * kSmHD + h * kSmDh)[j] = pack_chunk(o_hi);
  }
}

// ---------------------------------------------------------------------------------------
// backward, per token: dq, dk, dv from ctx, dctx and the k statistics (m, S).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sla_bwd_tokens_mma_kernel(const bf16* __restrict__ qkv,
                                                                 const bf16* __restrict__ dtok,
                                                                 const float* __restrict__ ctx,
                                                                 const float* __restrict__ dctx,
                                                                 const float* __restrict__ kstat,
                                                                 bf16* __restrict__ dqkv, int N) {
  extern __shared__ float smem[];
  float* sctx = smem;                // [8][32][32]
  float* sdctx = smem + 8 * 1024;    // [8][32][32]
  float* sm_m = smem + 16 * 1024;    // [8][32]
  float* sm_is = sm_m + 256;         // [8][32]  1/S
  float* sm_r = sm_is + 256;         // [8][32]  r[d] = sum_e dctx[d][e]*ctx[d][e]
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) {
    sctx[i] = ctx[(long)img * 8 * 1024 + i];
    sdctx[i] = dctx[(long)img * 8 * 1024 + i];
  }
  __syncthreads();
  {
    const int hh = threadIdx.x >> 5, dd = threadIdx.x & 31;
    sm_m[threadIdx.x] = kstat[((long)img * kSmHeads + hh) * 64 + dd];
    sm_is[threadIdx.x] = 1.f / kstat[((long)img * kSmHeads + hh) * 64 + 32 + dd];
    float r = 0.f;
    for (int e = 0; e < 32; ++e) r += sdctx[hh * 1024 + dd * 32 + e] * sctx[hh * 1024 + dd * 32 + e];
    sm_r[threadIdx.x] = r;
  }
  __syncthreads();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  BFrag b_ctxT, b_dctxT, b_dctx;
  load_bfrag(b_ctxT, sctx + h * 1024, 1, 32, g, j);    // B[k = e][n = d] = ctx[d][e]    (dq~ = dtok ctx^T)
  load_bfrag(b_dctxT, sdctx + h * 1024, 1, 32, g, j);  // B[k = e][n = d] = dctx[d][e]   (dk~ = v dctx^T)
  load_bfrag(b_dctx, sdctx + h * 1024, 32, 1, g, j);   // B[k = d][n = e] = dctx[d][e]   (dv = k~ dctx)
  float km[8], kis[8], kr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    km[i] = sm_m[h * 32 + 8 * j + i];
    kis[i] = sm_is[h * 32 + 8 * j + i];
    kr[i] = sm_r[h * 32 + 8 * j + i];
  }
  const int n_groups = (N + 15) / 16;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int n_lo = grp * 16 + g, n_hi = n_lo + 8;
    const bool v_lo = n_lo < N, v_hi = n_hi < N;
    const long r_lo = (long)img * N + (v_lo ? n_lo : N - 1), r_hi = (long)img * N + (v_hi ? n_hi : N - 1);
    const uint4* p_lo = reinterpret_cast<const uint4*>(qkv + r_lo * kSmQKV + h * kSmDh) + j;
    const uint4* p_hi = reinterpret_cast<const uint4*>(qkv + r_hi * kSmQKV + h * kSmDh) + j;
    // all eight 16-byte loads of the group are issued before any use
    const uint4 q_lo = __ldg(p_lo), q_hi = __ldg(p_hi);
    const uint4 k_lo = __ldg(p_lo + 32), k_hi = __ldg(p_hi + 32);   // +256 bf16
    const uint4 v_lo4 = __ldg(p_lo + 64), v_hi4 = __ldg(p_hi + 64);
    const uint4 g_lo = __ldg(reinterpret_cast<const uint4*>(dtok + r_lo * kSmHD + h * kSmDh) + j);
    const uint4 g_hi = __ldg(reinterpret_cast<const uint4*>(dtok + r_hi * kSmHD + h * kSmDh) + j);
    uint4* o_lo = reinterpret_cast<uint4*>(dqkv + r_lo * kSmQKV + h * kSmDh) + j;
    uint4* o_hi = reinterpret_cast<uint4*>(dqkv + r_hi * kSmQKV + h * kSmDh) + j;
    float a_lo[8], a_hi[8], t_lo[8], t_hi[8];
    // ---- dq = q~ (dq~ - <q~, dq~>),  dq~ = dtok ctx^T ----
    unpack_chunk(q_lo, a_lo);
    unpack_chunk(q_hi, a_hi);
    quad_softmax(a_lo);
    quad_softmax(a_hi);
    chunk_matmul(g_lo, g_hi, b_ctxT, t_lo, t_hi);
    float dot_lo = 0.f, dot_hi = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dot_lo = fmaf(a_lo[i], t_lo[i], dot_lo);
      dot_hi = fmaf(a_hi[i], t_hi[i], dot_hi);
    }
    dot_lo = quad_sum(dot_lo);
    dot_hi = quad_sum(dot_hi);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t_lo[i] = a_lo[i] * (t_lo[i] - dot_lo);
      t_hi[i] = a_hi[i] * (t_hi[i] - dot_hi);
    }
    if (v_lo) *o_lo = pack_chunk(t_lo);
    if (v_hi) *o_hi = pack_chunk(t_hi);
    // ---- dk = k~ (v dctx^T - r),  k~ = exp(k - m) / S ----
    unpack_chunk(k_lo, a_lo);
    unpack_chunk(k_hi, a_hi);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a_lo[i] = __expf(a_lo[i] - km[i]) * kis[i];
      a_hi[i] = __expf(a_hi[i] - km[i]) * kis[i];
    }
    chunk_matmul(v_lo4, v_hi4, b_dctxT, t_lo, t_hi);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t_lo[i] = a_lo[i] * (t_lo[i] - kr[i]);
      t_hi[i] = a_hi[i] * (t_hi[i] - kr[i]);
    }
    if (v_lo) o_lo[32] = pack_chunk(t_lo);
    if (v_hi) o_hi[32] = pack_chunk(t_hi);
    // ---- dv = k~ dctx ----
    chunk_matmul(pack_chunk(a_lo), pack_chunk(a_hi), b_dctx, t_lo, t_hi);
    if (v_lo) o_lo[64] = pack_chunk(t_lo);
    if (v_hi) o_hi[64] = pack_chunk(t_hi);
  }
}

// Host launchers used by vdn_sla_core_fwd / vdn_sla_core_bwd (attn.cu).
int sla_apply_mma_launch(const void* qkv, const float* ctx, void* tok_out, int n_img, int N, cudaStream_t st) {
  const int n_groups = (N + 15) / 16;
  const int gx = std::max(1, std::min(n_groups, std::max(1, 148 * 8 / n_img)));
  cudaError_t le = launch_pdl(sla_apply_mma_kernel, dim3(gx, n_img), dim3(256), (size_t)0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), ctx, reinterpret_cast<bf16*>(tok_out), N);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "sla_apply launch: %s", cudaGetErrorString(le));
  return check_launch("sla_apply_mma");
}

int sla_bwd_tokens_mma_launch(const void* qkv, const void* d_tok, const float* ctx, const float* dctx,
                              const float* kstat, void* dqkv, int n_img, int N, cudaStream_t st) {
  const size_t smem = (16 * 1024 + 3 * 256) * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(sla_bwd_tokens_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int n_groups = (N + 15) / 16;
  const int gx = std::max(1, std::min(n_groups, std::max(1, 148 * 3 / n_img)));
  cudaError_t le = launch_pdl(sla_bwd_tokens_mma_kernel, dim3(gx, n_img), dim3(256), smem, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(d_tok), ctx, dctx, kstat,
                              reinterpret_cast<bf16*>(dqkv), N);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "sla_bwd_tokens launch: %s", cudaGetErrorString(le));
  return check_launch("sla_bwd_tokens_mma");
}

}  // namespace vdn
