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

// the same with the fragment set in shared memory, laid out [(s*4 + t)*2 + r][32 lanes]
__device__ __forceinline__ void store_bfrag(uint32_t* sb, const BFrag& f, int lane) {
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      sb[((s * 4 + t) * 2 + 0) * 32 + lane] = f.r[s][t][0];
      sb[((s * 4 + t) * 2 + 1) * 32 + lane] = f.r[s][t][1];
    }
}
__device__ __forceinline__ void chunk_matmul_s(const uint4& lo, const uint4& hi, const uint32_t* sb, int lane,
                                               float (&olo)[8], float (&ohi)[8]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16_16816(d, lo.x, hi.x, lo.y, hi.y, sb[((0 * 4 + t) * 2 + 0) * 32 + lane], sb[((0 * 4 + t) * 2 + 1) * 32 + lane]);
    mma_bf16_16816(d, lo.z, hi.z, lo.w, hi.w, sb[((1 * 4 + t) * 2 + 0) * 32 + lane], sb[((1 * 4 + t) * 2 + 1) * 32 + lane]);
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
__global__ void __launch_bounds__(256, 2) sla_bwd_tokens_mma_kernel(const bf16* __restrict__ qkv,
                                                                 const bf16* __restrict__ dtok,
                                                                 const float* __restrict__ ctx,
                                                                 const float* __restrict__ dctx,
                                                                 const float* __restrict__ kstat,
                                                                 bf16* __restrict__ dqkv, int N) {
  // The three B-fragment sets of the warp's head live in shared memory as [16 words][32 lanes] (lane innermost:
  // conflict-free) - in registers they cost 48 per thread and cap the kernel at 8 warps per SM, too few to cover
  // the global-load latency of a kernel that should run at HBM speed.
  extern __shared__ uint32_t smem_b[];
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  // stage ctx / dctx of the frame through the same shared memory (coalesced), build the fragments in registers,
  // then overwrite the staging area with the lane-major fragment sets
  float* stage = reinterpret_cast<float*>(smem_b);  // [2][8][32][32] fp32 = 64 KB
  {
    const float4* c4 = reinterpret_cast<const float4*>(ctx + (long)img * kSmHeads * 1024);
    const float4* d4 = reinterpret_cast<const float4*>(dctx + (long)img * kSmHeads * 1024);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
      reinterpret_cast<float4*>(stage)[i] = __ldg(c4 + i);
      reinterpret_cast<float4*>(stage + 8192)[i] = __ldg(d4 + i);
    }
  }
  __syncthreads();
  const float* ch = stage + h * 1024;
  const float* dch = stage + 8192 + h * 1024;
  BFrag f_ctxT, f_dctxT, f_dctx;
  load_bfrag(f_ctxT, ch, 1, 32, g, j);    // B[k = e][n = d] = ctx[d][e]    (dq~ = dtok ctx^T)
  load_bfrag(f_dctxT, dch, 1, 32, g, j);  // B[k = e][n = d] = dctx[d][e]   (dk~ = v dctx^T)
  load_bfrag(f_dctx, dch, 32, 1, g, j);   // B[k = d][n = e] = dctx[d][e]   (dv = k~ dctx)
  float km[8], kis[8], kr[8];
  {
    // r[d] = sum_e dctx[d][e] * ctx[d][e]: lane = d, then broadcast the lane's 8 features through shuffles
    float r = 0.f;
    for (int e = 0; e < 32; ++e) {
      const int ee = (e + lane) & 31;  // rotated: conflict-free shared-memory reads
      r = fmaf(dch[lane * 32 + ee], ch[lane * 32 + ee], r);
    }
    const float m = kstat[((long)img * kSmHeads + h) * 64 + lane];
    const float is = 1.f / kstat[((long)img * kSmHeads + h) * 64 + 32 + lane];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      km[i] = __shfl_sync(0xffffffffu, m, 8 * j + i);
      kis[i] = __shfl_sync(0xffffffffu, is, 8 * j + i);
      kr[i] = __shfl_sync(0xffffffffu, r, 8 * j + i);
    }
  }
  __syncthreads();  // every warp is done reading the staging area
  uint32_t* sb = smem_b + h * (3 * 16 * 32);
  store_bfrag(sb, f_ctxT, lane);
  store_bfrag(sb + 16 * 32, f_dctxT, lane);
  store_bfrag(sb + 32 * 32, f_dctx, lane);
  __syncwarp();
  const uint32_t* b_ctxT = sb;
  const uint32_t* b_dctxT = sb + 16 * 32;
  const uint32_t* b_dctx = sb + 32 * 32;
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
    chunk_matmul_s(g_lo, g_hi, b_ctxT, lane, t_lo, t_hi);
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
    chunk_matmul_s(v_lo4, v_hi4, b_dctxT, lane, t_lo, t_hi);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t_lo[i] = a_lo[i] * (t_lo[i] - kr[i]);
      t_hi[i] = a_hi[i] * (t_hi[i] - kr[i]);
    }
    if (v_lo) o_lo[32] = pack_chunk(t_lo);
    if (v_hi) o_hi[32] = pack_chunk(t_hi);
    // ---- dv = k~ dctx ----
    chunk_matmul_s(pack_chunk(a_lo), pack_chunk(a_hi), b_dctx, lane, t_lo, t_hi);
    if (v_lo) o_lo[64] = pack_chunk(t_lo);
    if (v_hi) o_hi[64] = pack_chunk(t_hi);
  }
}

// ---------------------------------------------------------------------------------------
// Token-contraction products:  C[d][e] = sum_n A[n][d] * B[n][e]   (ctx = p^T v, dctx = q~^T dtok)
// MMA roles: M = d, N = e, K = tokens. Both operands need two consecutive TOKENS of one feature packed in a
// register while memory packs two consecutive FEATURES of one token; the 16-bit transposition happens in
// registers (cvt.bf16x2 for values that pass through fp32, prmt for raw bf16), again with permuted indices:
//   lane (g, j) owns tokens 4j..4j+3 of a 16-token group (logical k = 2j,2j+1 | 2j+8,2j+9)
//   m-tile u: row g <-> feature 16u+2g, row g+8 <-> feature 16u+2g+1   (one 32-bit word per token)
//   n-tile t: column n <-> feature 2n (t=0), 2n+1 (t=1), 2n+16 (t=2), 2n+17 (t=3)   (words g and g+8)
// The accumulators live in registers, so the online softmax over tokens (ctx) rescales them exactly.
// ---------------------------------------------------------------------------------------
struct TokAcc {
  float c[2][4][4];  // [m-tile][n-tile][d0..d3]
};
__device__ __forceinline__ void tokacc_zero(TokAcc& a) {
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) a.c[u][t][i] = 0.f;
}
// af[u][i][c]: A value of feature 16u+2g+i at the lane's token c (fp32); bw0/bw1[c]: raw B words g / g+8 of token c
__device__ __forceinline__ void tok_mma(TokAcc& acc, const float (&af)[2][2][4], const uint32_t (&bw0)[4],
                                        const uint32_t (&bw1)[4]) {
  uint32_t b[4][2];
  b[0][0] = __byte_perm(bw0[0], bw0[1], 0x5410); b[0][1] = __byte_perm(bw0[2], bw0[3], 0x5410);
  b[1][0] = __byte_perm(bw0[0], bw0[1], 0x7632); b[1][1] = __byte_perm(bw0[2], bw0[3], 0x7632);
  b[2][0] = __byte_perm(bw1[0], bw1[1], 0x5410); b[2][1] = __byte_perm(bw1[2], bw1[3], 0x5410);
  b[3][0] = __byte_perm(bw1[0], bw1[1], 0x7632); b[3][1] = __byte_perm(bw1[2], bw1[3], 0x7632);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t a0 = pack_bf16x2(af[u][0][0], af[u][0][1]), a1 = pack_bf16x2(af[u][1][0], af[u][1][1]);
    const uint32_t a2 = pack_bf16x2(af[u][0][2], af[u][0][3]), a3 = pack_bf16x2(af[u][1][2], af[u][1][3]);
#pragma unroll
    for (int t = 0; t < 4; ++t) mma_bf16_16816(acc.c[u][t], a0, a1, a2, a3, b[t][0], b[t][1]);
  }
}
// scatter the accumulators to out[d*32 + e] (plain store or atomicAdd)
template <bool kAtomic>
__device__ __forceinline__ void tokacc_store(const TokAcc& acc, float* out, int g, int j) {
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = 16 * u + 2 * g + (i >> 1);
        const int n = 2 * j + (i & 1);
        const int e = 2 * n + (t & 1) + 16 * (t >> 1);
        if (kAtomic) atomicAdd(out + d * 32 + e, acc.c[u][t][i]);
        else out[d * 32 + e] = acc.c[u][t][i];
      }
}

// Partial context over a token range with exact online softmax over tokens (per feature d).
// grid (n_split, n_img), 256 threads = 8 warps = 8 heads. Outputs the format sla_ctx_merge_kernel consumes.
__global__ void __launch_bounds__(256) sla_ctx_partial_mma_kernel(const bf16* __restrict__ qkv, int N,
                                                                  int tokens_per_split,
                                                                  float* __restrict__ ctx_part /*[img][h][split][32][32]*/,
                                                                  float* __restrict__ ms_part /*[img][h][split][2][32]*/) {
  pdl_trigger();
  pdl_wait();
  const int split = blockIdx.x, img = blockIdx.y, n_split = gridDim.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  const int n_begin = split * tokens_per_split;
  const int n_end = min(N, n_begin + tokens_per_split);
  TokAcc acc;
  tokacc_zero(acc);
  float mrow[2][2], srow[2][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mrow[u][i] = -INFINITY;
      srow[u][i] = 0.f;
    }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(qkv + (long)img * N * kSmQKV + h * kSmDh);
  // the words of the NEXT 16-token group are requested before the current group is processed (register double buffer):
  // one group per iteration left every warp a load round trip per 8 MMAs
  uint32_t nkw[2][4], nvw0[4], nvw1[4];
  auto fetch = [&](int n0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = min(n0 + 4 * j + c, n_end - 1);
      const uint32_t* row = base + (long)n * (kSmQKV / 2);
      nkw[0][c] = __ldg(row + kSmHD / 2 + g);
      nkw[1][c] = __ldg(row + kSmHD / 2 + 8 + g);
      nvw0[c] = __ldg(row + kSmHD + g);
      nvw1[c] = __ldg(row + kSmHD + 8 + g);
    }
  };
  if (n_begin < n_end) fetch(n_begin);
  for (int n0 = n_begin; n0 < n_end; n0 += 16) {
    uint32_t kw[2][4], vw0[4], vw1[4];
    bool valid[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      valid[c] = n0 + 4 * j + c < n_end;
      kw[0][c] = nkw[0][c];
      kw[1][c] = nkw[1][c];
      vw0[c] = nvw0[c];
      vw1[c] = nvw1[c];
    }
    if (n0 + 16 < n_end) fetch(n0 + 16);
    float kf[2][2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 f = unpack_bf16x2(kw[u][c]);
        kf[u][0][c] = valid[c] ? f.x : -INFINITY;  // padded tokens: exp() -> 0
        kf[u][1][c] = valid[c] ? f.y : -INFINITY;
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float gm = fmaxf(fmaxf(kf[u][i][0], kf[u][i][1]), fmaxf(kf[u][i][2], kf[u][i][3]));
        gm = quad_max(gm);  // over the 16 tokens of the group
        const float m_new = fmaxf(mrow[u][i], gm);
        const float corr = __expf(mrow[u][i] - m_new);  // first group: exp(-inf) = 0 on zero accumulators
        mrow[u][i] = m_new;
        srow[u][i] *= corr;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc.c[u][t][2 * i] *= corr;
          acc.c[u][t][2 * i + 1] *= corr;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          kf[u][i][c] = __expf(kf[u][i][c] - m_new);
          srow[u][i] += kf[u][i][c];
        }
      }
    tok_mma(acc, kf, vw0, vw1);
  }
  const long blk = ((long)img * kSmHeads + h) * n_split + split;
  tokacc_store<false>(acc, ctx_part + blk * 1024, g, j);
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float S = quad_sum(srow[u][i]);
      if (j == 0) {
        const int d = 16 * u + 2 * g + i;
        ms_part[blk * 64 + d] = mrow[u][i];
        ms_part[blk * 64 + 32 + d] = S;
      }
    }
}

// dctx[h][d][e] += sum_n q~[n,d] * dtok[n,e] over a token range (atomic across splits; dctx pre-zeroed).
__global__ void __launch_bounds__(256) sla_dctx_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dtok,
                                                           int N, int tokens_per_split, float* __restrict__ dctx) {
  pdl_trigger();
  pdl_wait();
  const int split = blockIdx.x, img = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  const int n_begin = split * tokens_per_split;
  const int n_end = min(N, n_begin + tokens_per_split);
  TokAcc acc;
  tokacc_zero(acc);
  const uint32_t* qbase = reinterpret_cast<const uint32_t*>(qkv + (long)img * N * kSmQKV + h * kSmDh);
  const uint32_t* gbase = reinterpret_cast<const uint32_t*>(dtok + (long)img * N * kSmHD + h * kSmDh);
  uint32_t nqw[2][4], ngw0[4], ngw1[4];  // next group's words (register double buffer, see sla_ctx_partial_mma_kernel)
  auto fetch = [&](int n0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long nn = min(n0 + 4 * j + c, n_end - 1);
      nqw[0][c] = __ldg(qbase + nn * (kSmQKV / 2) + g);
      nqw[1][c] = __ldg(qbase + nn * (kSmQKV / 2) + 8 + g);
      ngw0[c] = __ldg(gbase + nn * (kSmHD / 2) + g);
      ngw1[c] = __ldg(gbase + nn * (kSmHD / 2) + 8 + g);
    }
  };
  if (n_begin < n_end) fetch(n_begin);
  for (int n0 = n_begin; n0 < n_end; n0 += 16) {
    uint32_t qw[2][4], gw0[4], gw1[4];
    bool valid[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      valid[c] = n0 + 4 * j + c < n_end;
      qw[0][c] = nqw[0][c];
      qw[1][c] = nqw[1][c];
      gw0[c] = ngw0[c];
      gw1[c] = ngw1[c];
    }
    if (n0 + 16 < n_end) fetch(n0 + 16);
    float qf[2][2][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // softmax over the 32 features of token c: the 8 lanes with this j hold 4 features each
      const float2 f0 = unpack_bf16x2(qw[0][c]), f1 = unpack_bf16x2(qw[1][c]);
      float mx = fmaxf(fmaxf(f0.x, f0.y), fmaxf(f1.x, f1.y));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      const float e00 = __expf(f0.x - mx), e01 = __expf(f0.y - mx), e10 = __expf(f1.x - mx), e11 = __expf(f1.y - mx);
      float sm = (e00 + e01) + (e10 + e11);
      sm += __shfl_xor_sync(0xffffffffu, sm, 4);
      sm += __shfl_xor_sync(0xffffffffu, sm, 8);
      sm += __shfl_xor_sync(0xffffffffu, sm, 16);
      const float inv = valid[c] ? 1.f / sm : 0.f;  // padded tokens contribute nothing
      qf[0][0][c] = e00 * inv;
      qf[0][1][c] = e01 * inv;
      qf[1][0][c] = e10 * inv;
      qf[1][1][c] = e11 * inv;
    }
    tok_mma(acc, qf, gw0, gw1);
  }
  tokacc_store<true>(acc, dctx + ((long)img * kSmHeads + h) * 1024, g, j);
}

// ---------------------------------------------------------------------------------------
// Fused SpatialLinearAttention forward for C = 32 (inference engines): x -> out with NO q/k/v or tok tensors in
// HBM (modules.py:99-129: to_q/to_k/to_v 1x1 convs without bias, core, to_out 1x1 conv, + residual).
//   pass 1 (sla_ctx_fused_kernel):   k^T = W_k^T x^T, v^T = W_v^T x^T per (16 tokens, head) as MMAs whose
//       accumulator fragments ARE the A / B fragments of ctx += p v (features x tokens), online softmax as above
//   pass 2 (sla_apply_fused_kernel): q = x W_q (chunk layout), q~ = softmax_D, tok = q~ ctx, out_h = tok W_o,h;
//       the 8 heads (= 8 warps of the block) are summed through shared memory, + x, one bf16 store.
// At the 64x64 level this replaces 252 MB of qkv writes + reads and 84 MB of tok traffic per call by 2 reads of x.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) sla_ctx_fused_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w_qkv /*[768][32]*/,
                                                               int N, int tokens_per_split,
                                                               float* __restrict__ ctx_part, float* __restrict__ ms_part) {
  pdl_trigger();
  pdl_wait();
  const int split = blockIdx.x, img = blockIdx.y, n_split = gridDim.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  const int n_begin = split * tokens_per_split;
  const int n_end = min(N, n_begin + tokens_per_split);
  // A fragments of W^T (rows = features 16u+g | 16u+8+g, k = channels in chunk order 8j+4s+{0,1} | +2)
  uint32_t wk[2][2][4], wv[2][2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t* k0 = reinterpret_cast<const uint32_t*>(w_qkv + (256 + h * 32 + 16 * u + g) * 32);
    const uint32_t* k1 = reinterpret_cast<const uint32_t*>(w_qkv + (256 + h * 32 + 16 * u + 8 + g) * 32);
    const uint32_t* v0 = reinterpret_cast<const uint32_t*>(w_qkv + (512 + h * 32 + 16 * u + g) * 32);
    const uint32_t* v1 = reinterpret_cast<const uint32_t*>(w_qkv + (512 + h * 32 + 16 * u + 8 + g) * 32);
#pragma unroll
    for (int sx = 0; sx < 2; ++sx) {
      wk[u][sx][0] = __ldg(k0 + 4 * j + 2 * sx); wk[u][sx][1] = __ldg(k1 + 4 * j + 2 * sx);
      wk[u][sx][2] = __ldg(k0 + 4 * j + 2 * sx + 1); wk[u][sx][3] = __ldg(k1 + 4 * j + 2 * sx + 1);
      wv[u][sx][0] = __ldg(v0 + 4 * j + 2 * sx); wv[u][sx][1] = __ldg(v1 + 4 * j + 2 * sx);
      wv[u][sx][2] = __ldg(v0 + 4 * j + 2 * sx + 1); wv[u][sx][3] = __ldg(v1 + 4 * j + 2 * sx + 1);
    }
  }
  float acc[2][4][4];  // [m-tile u: d = 16u+g | 16u+8+g][n-tile (u', r): e = 16u'+8r + 2j+{0,1}][frag]
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[u][t][i] = 0.f;
  float mrow[2][2], srow[2][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mrow[u][r] = -INFINITY;
      srow[u][r] = 0.f;
    }
  const bf16* xb = x + (long)img * N * 32;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int n0 = n_begin; n0 < n_end; n0 += 16) {
    // x^T operand: chunk j of token rows n0+g (n-tile 0) and n0+8+g (n-tile 1)
    const int t_lo = n0 + g, t_hi = n0 + 8 + g;
    const uint4 x_lo = t_lo < n_end ? __ldg(reinterpret_cast<const uint4*>(xb + (long)t_lo * 32) + j) : zero4;
    const uint4 x_hi = t_hi < n_end ? __ldg(reinterpret_cast<const uint4*>(xb + (long)t_hi * 32) + j) : zero4;
    float kt[2][2][4], vt[2][2][4];  // [u][token tile][frag]: rows = features, cols = tokens 8t+2j+{0,1}
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint4& xc = t == 0 ? x_lo : x_hi;
#pragma unroll
        for (int i = 0; i < 4; ++i) kt[u][t][i] = vt[u][t][i] = 0.f;
        mma_bf16_16816(kt[u][t], wk[u][0][0], wk[u][0][1], wk[u][0][2], wk[u][0][3], xc.x, xc.y);
        mma_bf16_16816(kt[u][t], wk[u][1][0], wk[u][1][1], wk[u][1][2], wk[u][1][3], xc.z, xc.w);
        mma_bf16_16816(vt[u][t], wv[u][0][0], wv[u][0][1], wv[u][0][2], wv[u][0][3], xc.x, xc.y);
        mma_bf16_16816(vt[u][t], wv[u][1][0], wv[u][1][1], wv[u][1][2], wv[u][1][3], xc.z, xc.w);
      }
    // k, v are rounded to bf16 exactly like the materialised path (projection output dtype)
    // online softmax over tokens, per feature row (u, r): the lane holds 4 of the group's 16 tokens
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float kv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int tok = n0 + 8 * (c >> 1) + 2 * j + (c & 1);
          const float kq = __bfloat162float(__float2bfloat16(kt[u][c >> 1][2 * r + (c & 1)]));
          kv[c] = tok < n_end ? kq : -INFINITY;  // padded tokens: exp() -> 0
        }
        float gm = fmaxf(fmaxf(kv[0], kv[1]), fmaxf(kv[2], kv[3]));
        gm = quad_max(gm);
        const float m_new = fmaxf(mrow[u][r], gm);
        const float corr = __expf(mrow[u][r] - m_new);
        mrow[u][r] = m_new;
        srow[u][r] *= corr;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          acc[u][t][2 * r] *= corr;
          acc[u][t][2 * r + 1] *= corr;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float pv = __expf(kv[c] - m_new);
          srow[u][r] += pv;
          kt[u][c >> 1][2 * r + (c & 1)] = pv;
        }
      }
    // ctx += p^T v: A = p fragments (rows = d), B = v^T fragments (k = tokens, n = e)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t a0 = pack_bf16x2(kt[u][0][0], kt[u][0][1]), a1 = pack_bf16x2(kt[u][0][2], kt[u][0][3]);
      const uint32_t a2 = pack_bf16x2(kt[u][1][0], kt[u][1][1]), a3 = pack_bf16x2(kt[u][1][2], kt[u][1][3]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {  // n-tile t = (u', r) = (t / 2, t % 2): e = 16u' + 8r + n
        const int up = t >> 1, r = t & 1;
        mma_bf16_16816(acc[u][t], a0, a1, a2, a3, pack_bf16x2(vt[up][0][2 * r], vt[up][0][2 * r + 1]),
                       pack_bf16x2(vt[up][1][2 * r], vt[up][1][2 * r + 1]));
      }
    }
  }
  const long blk = ((long)img * kSmHeads + h) * n_split + split;
  float* cp = ctx_part + blk * 1024;
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = 16 * u + 8 * (i >> 1) + g;
        const int e = 16 * (t >> 1) + 8 * (t & 1) + 2 * j + (i & 1);
        cp[d * 32 + e] = acc[u][t][i];
      }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float S = quad_sum(srow[u][r]);
      if (j == 0) {
        const int d = 16 * u + 8 * r + g;
        ms_part[blk * 64 + d] = mrow[u][r];
        ms_part[blk * 64 + 32 + d] = S;
      }
    }
}

// B fragments of a bf16 K-major matrix Wt[n][k] (row stride ld elements): B[k][n] = Wt[n_of(col)][k0 + k], with the
// chunk-order contraction index (8j+4s+{0..3}) and an arbitrary column -> row map given by the caller.
__device__ __forceinline__ void load_bfrag_kmajor(BFrag& f, const bf16* Wt, int ld, int k0, int g, int j, bool psi_cols) {
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      // psi_cols: projection whose OUTPUT should land in chunk layout (column c of tile t -> 8(c/2)+4(t/2)+2(t%2)+c%2);
      // otherwise the sigma map of chunk_matmul (8(c/2) + 2t + c%2) - both give lane j the features 8j..8j+7
      const int n = psi_cols ? 8 * (g >> 1) + 4 * (t >> 1) + 2 * (t & 1) + (g & 1) : 8 * (g >> 1) + 2 * t + (g & 1);
      const uint32_t* row = reinterpret_cast<const uint32_t*>(Wt + (long)n * ld + k0);
      f.r[s][t][0] = __ldg(row + 4 * j + 2 * s);
      f.r[s][t][1] = __ldg(row + 4 * j + 2 * s + 1);
    }
}

__global__ void __launch_bounds__(256, 2) sla_apply_fused_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w_qkv /*[768][32]*/,
                                                                 const bf16* __restrict__ w_out /*[32][256]*/,
                                                                 const float* __restrict__ ctx, bf16* __restrict__ out,
                                                                 int N) {
  __shared__ float s_part[8][16][33];  // per head: [16 tokens][32 channels] partial of the to_out projection
  extern __shared__ uint32_t s_frag[];  // per head: q-projection | ctx | to_out fragment sets, [16 words][32 lanes] each
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j = lane & 3;
  uint32_t* sb = s_frag + h * (3 * 16 * 32);
  {
    BFrag f;
    load_bfrag_kmajor(f, w_qkv + (long)(h * 32) * 32, 32, 0, g, j, false);  // q = x W_q,h (sigma columns: chunk layout)
    store_bfrag(sb, f, lane);
    load_bfrag(f, ctx + ((long)img * kSmHeads + h) * 1024, 32, 1, g, j);     // B[k = d][n = e] = ctx[d][e]
    store_bfrag(sb + 16 * 32, f, lane);
    load_bfrag_kmajor(f, w_out, 256, h * 32, g, j, false);                   // B[k = e][n = c] = W_o[h*32+e][c]
    store_bfrag(sb + 32 * 32, f, lane);
  }
  __syncwarp();
  const bf16* xb = x + (long)img * N * 32;
  bf16* ob = out + (long)img * N * 32;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const int n_groups = (N + 15) / 16;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {  // same trip count for every warp of the block
    const int n_lo = grp * 16 + g, n_hi = n_lo + 8;
    const uint4 x_lo = n_lo < N ? __ldg(reinterpret_cast<const uint4*>(xb + (long)n_lo * 32) + j) : zero4;
    const uint4 x_hi = n_hi < N ? __ldg(reinterpret_cast<const uint4*>(xb + (long)n_hi * 32) + j) : zero4;
    float a_lo[8], a_hi[8], o_lo[8], o_hi[8];
    chunk_matmul_s(x_lo, x_hi, sb, lane, a_lo, a_hi);                          // q (bf16-rounded like the GEMM output)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a_lo[i] = __bfloat162float(__float2bfloat16(a_lo[i]));
      a_hi[i] = __bfloat162float(__float2bfloat16(a_hi[i]));
    }
    quad_softmax(a_lo);
    quad_softmax(a_hi);
    chunk_matmul_s(pack_chunk(a_lo), pack_chunk(a_hi), sb + 16 * 32, lane, o_lo, o_hi);   // tok = q~ ctx
    chunk_matmul_s(pack_chunk(o_lo), pack_chunk(o_hi), sb + 32 * 32, lane, a_lo, a_hi);   // out_h = tok W_o,h
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s_part[h][g][8 * j + i] = a_lo[i];
      s_part[h][g + 8][8 * j + i] = a_hi[i];
    }
    __syncthreads();
    // sum the 8 heads, add the residual, store: thread -> (token r = tid / 16, channels 2*(tid % 16), +1)
    {
      const int r = threadIdx.x >> 4, c = (threadIdx.x & 15) * 2;
      const int n = grp * 16 + r;
      if (n < N) {
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int hh = 0; hh < 8; ++hh) {
          v0 += s_part[hh][r][c];
          v1 += s_part[hh][r][c + 1];
        }
        const float2 xr = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xb + (long)n * 32 + c));
        *reinterpret_cast<uint32_t*>(ob + (long)n * 32 + c) = pack_bf16x2(v0 + xr.x, v1 + xr.y);
      }
    }
    __syncthreads();
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
  const size_t smem = (size_t)2 * 8 * 1024 * sizeof(float);  // 64 KB staging, reused for the 48 KB of B fragments
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(sla_bwd_tokens_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int n_groups = (N + 15) / 16;
  const int gx = std::max(1, std::min(n_groups, std::max(1, 148 * 4 / n_img)));
  cudaError_t le = launch_pdl(sla_bwd_tokens_mma_kernel, dim3(gx, n_img), dim3(256), smem, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(d_tok), ctx, dctx, kstat,
                              reinterpret_cast<bf16*>(dqkv), N);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "sla_bwd_tokens launch: %s", cudaGetErrorString(le));
  return check_launch("sla_bwd_tokens_mma");
}

int sla_ctx_partial_mma_launch(const void* qkv, int N, int tokens_per_split, int n_split, float* ctx_part,
                               float* ms_part, int n_img, cudaStream_t st) {
  cudaError_t le = launch_pdl(sla_ctx_partial_mma_kernel, dim3(n_split, n_img), dim3(256), (size_t)0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), N, tokens_per_split, ctx_part, ms_part);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "sla_ctx_partial launch: %s", cudaGetErrorString(le));
  return check_launch("sla_ctx_partial_mma");
}

int sla_dctx_mma_launch(const void* qkv, const void* d_tok, int N, int tokens_per_split, int n_split, float* dctx,
                        int n_img, cudaStream_t st) {
  cudaError_t le = launch_pdl(sla_dctx_mma_kernel, dim3(n_split, n_img), dim3(256), (size_t)0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(d_tok), N,
                              tokens_per_split, dctx);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "sla_dctx launch: %s", cudaGetErrorString(le));
  return check_launch("sla_dctx_mma");
}

// x -> out fused forward (C = 32). The merge of the split partials (sla_ctx_merge_kernel, attn.cu) is launched by the
// caller between the two passes.
int sla_ctx_fused_launch(const void* x, const void* w_qkv, int N, int tokens_per_split, int n_split, float* ctx_part,
                         float* ms_part, int n_img, cudaStream_t st) {
  cudaError_t le = launch_pdl(sla_ctx_fused_kernel, dim3(n_split, n_img), dim3(256), (size_t)0, st, 1,
                              reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(w_qkv), N,
                              tokens_per_split, ctx_part, ms_part);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "sla_ctx_fused launch: %s", cudaGetErrorString(le));
  return check_launch("sla_ctx_fused");
}

int sla_apply_fused_launch(const void* x, const void* w_qkv, const void* w_out, const float* ctx, void* out, int n_img,
                           int N, cudaStream_t st) {
  const size_t smem = (size_t)8 * 3 * 16 * 32 * sizeof(uint32_t);
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(sla_apply_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int n_groups = (N + 15) / 16;
  const int gx = std::max(1, std::min(n_groups, std::max(1, 148 * 4 / n_img)));
  cudaError_t le = launch_pdl(sla_apply_fused_kernel, dim3(gx, n_img), dim3(256), smem, st, 1,
                              reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(w_qkv),
                              reinterpret_cast<const bf16*>(w_out), ctx, reinterpret_cast<bf16*>(out), N);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "sla_apply_fused launch: %s", cudaGetErrorString(le));
  return check_launch("sla_apply_fused");
}

}  // namespace vdn
