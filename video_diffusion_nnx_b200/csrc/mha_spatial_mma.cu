// Mid-block spatial MultiheadAttention core on warp-level tensor-core MMAs, forward and backward
// (modules.py:285-324 in the 'b f (h w) c' arrangement of unet3d.py:196-205; backward = what jax.value_and_grad
// derives at trainer.py:361). Sequences are the H*W pixels of one frame (64 at config_v2_2, 256 at v2_3x), 8 heads
// of 32 features, q | k | v head-major in the fused projection [P][768].
//
// The CUDA-core kernels this replaces (attn.cu, lane pair per (token, head), keys streamed from global memory) took
// 66 + 45 + 49 us for the 168 MFLOP of config_v2_2's single instance. Here a CTA owns a 64-token block of one
// (sequence, head): four warps of 16 rows each, the "other side" of every product (keys / values in the forward and
// the dQ pass, queries / dO in the dK / dV pass) is staged in shared memory 64 tokens at a time and read with
// ldmatrix (plain for row-major B operands, .trans where the contraction runs over tokens); softmax is the exact
// online form over key blocks. The backward needs no atomics and no D workspace: a warp computes dQ for its 16 query
// rows (S, dP with rows = queries) and dK / dV for its 16 key rows (S^T, dP^T recomputed with rows = keys), and
// D = rowsum(dO * O) is recomputed per staged block.
#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {
namespace {

constexpr int kSpHeads = 8, kSpDh = 32, kSpHD = 256, kSpQKV = 768;
constexpr int kSpBlk = 64;    // tokens per staged block = rows per CTA
constexpr int kSpPitch = 40;  // bf16 per staged row: 64 B of payload + 16 B pad (conflict-free ldmatrix)

__device__ __forceinline__ void sp_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ float quad_sum_f(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float quad_max_f(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// 64 x 32 head slice (rows of `ld` elements in global memory) -> shared [64][kSpPitch]; 128 threads, 2 chunks each
__device__ __forceinline__ void stage64(bf16* s, const bf16* g, int ld) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = threadIdx.x + u * 128;
    const int row = idx >> 2, ch = idx & 3;
    *reinterpret_cast<uint4*>(s + row * kSpPitch + ch * 8) = __ldg(reinterpret_cast<const uint4*>(g + (long)row * ld + ch * 8));
  }
}
// A-operand fragments (m16 x k32 = two k steps) of rows g / g+8 of a row-major head slice in global memory
__device__ __forceinline__ void load_afrag(uint32_t (&a)[2][4], const bf16* rows, int ld, int g, int t) {
  const uint32_t* lo = reinterpret_cast<const uint32_t*>(rows + (long)g * ld);
  const uint32_t* hi = reinterpret_cast<const uint32_t*>(rows + (long)(g + 8) * ld);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    a[ks][0] = __ldg(lo + 8 * ks + t);
    a[ks][1] = __ldg(hi + 8 * ks + t);
    a[ks][2] = __ldg(lo + 8 * ks + 4 + t);
    a[ks][3] = __ldg(hi + 8 * ks + 4 + t);
  }
}
// c[nt] (16 x 64, n-tile nt = staged rows 8nt..8nt+7) = A[16 x 32] * X^T, X = staged [64][32] (contraction over features)
__device__ __forceinline__ void rows_times_staged_t(const uint32_t (&a)[2][4], const bf16* sX, int lane, float (&c)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b[4];
    ldsm_x4(b, sX + (nt * 8 + (lane & 7)) * kSpPitch + (lane >> 3) * 8);
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    sp_mma(c[nt], a[0], b[0], b[1]);
    sp_mma(c[nt], a[1], b[2], b[3]);
  }
}
// acc[16 x 32] += P[16 x 64] * X, P given as accumulator-layout fp32 (packed to bf16 A fragments), X = staged [64][32]
// (contraction over the 64 staged tokens)
__device__ __forceinline__ void probs_times_staged(const float (&p)[8][4], const bf16* sX, int lane, float (&acc)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * ks][0], p[2 * ks][1]);
    a[1] = pack_bf16x2(p[2 * ks][2], p[2 * ks][3]);
    a[2] = pack_bf16x2(p[2 * ks + 1][0], p[2 * ks + 1][1]);
    a[3] = pack_bf16x2(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, sX + (16 * ks + (lane & 7) + 8 * ((lane >> 3) & 1)) * kSpPitch + 16 * np + 8 * (lane >> 4));
      sp_mma(acc[2 * np], a, b[0], b[1]);
      sp_mma(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}
// rows g / g+8 of a 16 x 32 accumulator -> bf16 head slice in global memory
__device__ __forceinline__ void store_rows32(const float (&acc)[4][4], bf16* rows, int ld, int g, int t, float s_lo, float s_hi) {
  uint32_t* lo = reinterpret_cast<uint32_t*>(rows + (long)g * ld);
  uint32_t* hi = reinterpret_cast<uint32_t*>(rows + (long)(g + 8) * ld);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    lo[4 * nt + t] = pack_bf16x2(acc[nt][0] * s_lo, acc[nt][1] * s_lo);
    hi[4 * nt + t] = pack_bf16x2(acc[nt][2] * s_hi, acc[nt][3] * s_hi);
  }
}

__global__ void __launch_bounds__(128) mha_spatial_mma_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o,
                                                                  float* __restrict__ lse, int S) {
  __shared__ __align__(16) bf16 sK[kSpBlk * kSpPitch];
  __shared__ __align__(16) bf16 sV[kSpBlk * kSpPitch];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y;
  const long row0 = (long)blockIdx.z * S;                  // first token row of the sequence
  const long own = row0 + blockIdx.x * kSpBlk + 16 * w;    // first of this warp's 16 query rows
  const float scale = rsqrtf((float)kSpDh);
  pdl_trigger();
  pdl_wait();
  uint32_t qa[2][4];
  load_afrag(qa, qkv + own * kSpQKV + h * kSpDh, kSpQKV, g, t);
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int kb = 0; kb < S; kb += kSpBlk) {
    __syncthreads();
    stage64(sK, qkv + (row0 + kb) * kSpQKV + kSpHD + h * kSpDh, kSpQKV);
    stage64(sV, qkv + (row0 + kb) * kSpQKV + 2 * kSpHD + h * kSpDh, kSpQKV);
    __syncthreads();
    float s[8][4];
    rows_times_staged_t(qa, sK, lane, s);
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[nt][i] *= scale;
      mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
    }
    const float mn_lo = fmaxf(m_lo, quad_max_f(mx_lo)), mn_hi = fmaxf(m_hi, quad_max_f(mx_hi));
    const float c_lo = __expf(m_lo - mn_lo), c_hi = __expf(m_hi - mn_hi);
    m_lo = mn_lo;
    m_hi = mn_hi;
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = __expf(s[nt][0] - mn_lo);
      s[nt][1] = __expf(s[nt][1] - mn_lo);
      s[nt][2] = __expf(s[nt][2] - mn_hi);
      s[nt][3] = __expf(s[nt][3] - mn_hi);
      sum_lo += s[nt][0] + s[nt][1];
      sum_hi += s[nt][2] + s[nt][3];
    }
    l_lo = l_lo * c_lo + sum_lo;
    l_hi = l_hi * c_hi + sum_hi;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      acc[nt][0] *= c_lo;
      acc[nt][1] *= c_lo;
      acc[nt][2] *= c_hi;
      acc[nt][3] *= c_hi;
    }
    probs_times_staged(s, sV, lane, acc);
  }
  l_lo = quad_sum_f(l_lo);
  l_hi = quad_sum_f(l_hi);
  store_rows32(acc, o + own * kSpHD + h * kSpDh, kSpHD, g, t, 1.f / l_lo, 1.f / l_hi);
  if (t == 0) {
    lse[(own + g) * kSpHeads + h] = m_lo + __logf(l_lo);
    lse[(own + g + 8) * kSpHeads + h] = m_hi + __logf(l_hi);
  }
}

__global__ void __launch_bounds__(128) mha_spatial_mma_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                                  const bf16* __restrict__ d_o,
                                                                  const float* __restrict__ lse, bf16* __restrict__ dqkv,
                                                                  int S) {
  __shared__ __align__(16) bf16 sA[kSpBlk * kSpPitch];
  __shared__ __align__(16) bf16 sB[kSpBlk * kSpPitch];
  __shared__ float sL[kSpBlk], sD[kSpBlk];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y;
  const long row0 = (long)blockIdx.z * S;
  const long own = row0 + blockIdx.x * kSpBlk + 16 * w;
  const float scale = rsqrtf((float)kSpDh);
  pdl_trigger();
  pdl_wait();

  // ---------------- dQ for this warp's 16 query rows: loop over key blocks ----------------
  {
    uint32_t qa[2][4], da[2][4], oa[2][4];
    load_afrag(qa, qkv + own * kSpQKV + h * kSpDh, kSpQKV, g, t);
    load_afrag(da, d_o + own * kSpHD + h * kSpDh, kSpHD, g, t);
    load_afrag(oa, o + own * kSpHD + h * kSpDh, kSpHD, g, t);
    float D_lo = 0.f, D_hi = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 a = unpack_bf16x2(da[ks][i]), b = unpack_bf16x2(oa[ks][i]);
        const float v = a.x * b.x + a.y * b.y;
        if (i & 1) D_hi += v; else D_lo += v;
      }
    D_lo = quad_sum_f(D_lo);
    D_hi = quad_sum_f(D_hi);
    const float L_lo = lse[(own + g) * kSpHeads + h], L_hi = lse[(own + g + 8) * kSpHeads + h];
    float dq[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int kb = 0; kb < S; kb += kSpBlk) {
      __syncthreads();
      stage64(sA, qkv + (row0 + kb) * kSpQKV + kSpHD + h * kSpDh, kSpQKV);      // K block
      stage64(sB, qkv + (row0 + kb) * kSpQKV + 2 * kSpHD + h * kSpDh, kSpQKV);  // V block
      __syncthreads();
      float s[8][4], dp[8][4];
      rows_times_staged_t(qa, sA, lane, s);
      rows_times_staged_t(da, sB, lane, dp);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = __expf(s[nt][0] * scale - L_lo) * (dp[nt][0] - D_lo) * scale;
        s[nt][1] = __expf(s[nt][1] * scale - L_lo) * (dp[nt][1] - D_lo) * scale;
        s[nt][2] = __expf(s[nt][2] * scale - L_hi) * (dp[nt][2] - D_hi) * scale;
        s[nt][3] = __expf(s[nt][3] * scale - L_hi) * (dp[nt][3] - D_hi) * scale;
      }
      probs_times_staged(s, sA, lane, dq);  // dQ += dS K
    }
    store_rows32(dq, dqkv + own * kSpQKV + h * kSpDh, kSpQKV, g, t, 1.f, 1.f);
  }

  // ---------------- dK, dV for this warp's 16 key rows: loop over query blocks ----------------
  {
    uint32_t ka[2][4], va[2][4];
    load_afrag(ka, qkv + own * kSpQKV + kSpHD + h * kSpDh, kSpQKV, g, t);
    load_afrag(va, qkv + own * kSpQKV + 2 * kSpHD + h * kSpDh, kSpQKV, g, t);
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qb = 0; qb < S; qb += kSpBlk) {
      __syncthreads();
      stage64(sA, qkv + (row0 + qb) * kSpQKV + h * kSpDh, kSpQKV);  // Q block
      stage64(sB, d_o + (row0 + qb) * kSpHD + h * kSpDh, kSpHD);    // dO block
      if (threadIdx.x < kSpBlk) {
        const long r = row0 + qb + threadIdx.x;
        const uint4* po = reinterpret_cast<const uint4*>(o + r * kSpHD + h * kSpDh);
        const uint4* pd = reinterpret_cast<const uint4*>(d_o + r * kSpHD + h * kSpDh);
        float D = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 x = __ldg(po + c), y = __ldg(pd + c);
          float2 a, b;
          a = unpack_bf16x2(x.x); b = unpack_bf16x2(y.x); D += a.x * b.x + a.y * b.y;
          a = unpack_bf16x2(x.y); b = unpack_bf16x2(y.y); D += a.x * b.x + a.y * b.y;
          a = unpack_bf16x2(x.z); b = unpack_bf16x2(y.z); D += a.x * b.x + a.y * b.y;
          a = unpack_bf16x2(x.w); b = unpack_bf16x2(y.w); D += a.x * b.x + a.y * b.y;
        }
        sD[threadIdx.x] = D;
        sL[threadIdx.x] = lse[r * kSpHeads + h];
      }
      __syncthreads();
      float st[8][4], dpt[8][4];
      rows_times_staged_t(ka, sA, lane, st);   // S^T = K Q^T   (rows = keys, columns = queries)
      rows_times_staged_t(va, sB, lane, dpt);  // dP^T = V dO^T
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c0 = 8 * nt + 2 * t;
        const float L0 = sL[c0], L1 = sL[c0 + 1], D0 = sD[c0], D1 = sD[c0 + 1];
        const float p0 = __expf(st[nt][0] * scale - L0), p1 = __expf(st[nt][1] * scale - L1);
        const float p2 = __expf(st[nt][2] * scale - L0), p3 = __expf(st[nt][3] * scale - L1);
        st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
        dpt[nt][0] = p0 * (dpt[nt][0] - D0) * scale;
        dpt[nt][1] = p1 * (dpt[nt][1] - D1) * scale;
        dpt[nt][2] = p2 * (dpt[nt][2] - D0) * scale;
        dpt[nt][3] = p3 * (dpt[nt][3] - D1) * scale;
      }
      probs_times_staged(st, sB, lane, dv);   // dV += P^T dO
      probs_times_staged(dpt, sA, lane, dk);  // dK += dS^T Q
    }
    store_rows32(dk, dqkv + own * kSpQKV + kSpHD + h * kSpDh, kSpQKV, g, t, 1.f, 1.f);
    store_rows32(dv, dqkv + own * kSpQKV + 2 * kSpHD + h * kSpDh, kSpQKV, g, t, 1.f, 1.f);
  }
}

}  // namespace

bool mha_spatial_mma_applicable(int HW) { return HW >= kSpBlk && HW % kSpBlk == 0; }

int mha_spatial_mma_fwd_launch(const void* qkv, void* o, float* lse, int n_seq, int S, cudaStream_t st) {
  cudaError_t le = launch_pdl(mha_spatial_mma_fwd_kernel, dim3(S / kSpBlk, kSpHeads, n_seq), dim3(128), 0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(o), lse, S);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_spatial_mma_fwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_spatial_mma_fwd");
}

int mha_spatial_mma_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int n_seq,
                               int S, cudaStream_t st) {
  cudaError_t le = launch_pdl(mha_spatial_mma_bwd_kernel, dim3(S / kSpBlk, kSpHeads, n_seq), dim3(128), 0, st, 1,
                              reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(o),
                              reinterpret_cast<const bf16*>(d_o), lse, reinterpret_cast<bf16*>(dqkv), S);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "mha_spatial_mma_bwd launch: %s", cudaGetErrorString(le));
  return check_launch("mha_spatial_mma_bwd");
}

}  // namespace vdn
