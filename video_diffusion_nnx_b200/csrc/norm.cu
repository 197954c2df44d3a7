// GroupNorm / LayerNorm / SiLU kernels of Block and ResnetBlock (modules.py:150-243), forward
// and backward. HBM/L2-bound elementwise + reduction work: 16-byte vector accesses, per-sample
// affine coefficients staged in shared memory, warp-shuffle reductions.
//
// GroupNorm statistics (per sample, per group, over F*H*W*C/G elements - flax semantics, eps 1e-6,
// fast variance) arrive as raw sums [B][G][2] = (sum x, sum x^2) accumulated by the producing
// conv's epilogue (tapgemm.cu).
#include <algorithm>
#include <cstdlib>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr float kEps = 1e-6f;
constexpr int kNormThreads = 256;

struct GnArgs {
  const bf16* x;        // raw conv output [B][rows][C]
  const float* sums;    // [kGnReplicas][B][G][2]
  const float* gamma;   // [C]
  const float* beta;    // [C]
  const float* ss;      // [B][ss_ld] scale = [0,C), shift = [C,2C); or null
  int ss_ld;
  int B, rows, C, G;
};

constexpr int kMaxGroups = 16;

// (mean, rstd) of every group of sample b into shared memory: thread (g, r) reads replica r of group g, the 16
// replicas are added up with shuffles (one coalesced round trip instead of 16 dependent loads per CHANNEL - at
// C = 1024 the per-block statistics prologue used to cost more than the block's payload). Ends with a barrier.
// Must be called by every thread of the block (blockDim >= 16 * G).
__device__ __forceinline__ const float2* gn_group_stats(const GnArgs& a, int b) {
  __shared__ float2 s_stat[kMaxGroups];
  static_assert(kGnReplicas == 16, "one 16-lane segment per group");
  const int g = threadIdx.x >> 4, r = threadIdx.x & 15;
  float s1 = 0.f, s2 = 0.f;
  if (g < a.G) {
    const float2 v = *reinterpret_cast<const float2*>(a.sums + ((long)(r * a.B + b) * a.G + g) * 2);
    s1 = v.x;
    s2 = v.y;
  }
  if (threadIdx.x < 16 * kMaxGroups) {  // whole warps: the shuffles below are convergent
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (r == 0 && g < a.G) {
      const float inv_n = 1.f / ((float)a.rows * (float)(a.C / a.G));
      const float mean = s1 * inv_n;
      const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
      s_stat[g] = make_float2(mean, rsqrtf(var + kEps));
    }
  }
  __syncthreads();
  return s_stat;
}

// Per-sample affine in smem: y = x * A[c] + Bc[c]  (GN * gamma + beta, then *(scale+1)+shift). Returns the
// per-group (mean, rstd) table (shared memory).
__device__ __forceinline__ const float2* gn_affine_to_smem(const GnArgs& a, int b, float* sA, float* sB) {
  // the parameters of this thread's first channel are requested BEFORE the statistics round trip (they overlap)
  const int c_first = threadIdx.x;
  float ga = 0.f, be = 0.f, sc = 1.f, sh = 0.f;
  if (c_first < a.C) {
    ga = a.gamma[c_first];
    be = a.beta[c_first];
    if (a.ss) {
      sc = a.ss[(long)b * a.ss_ld + c_first] + 1.f;
      sh = a.ss[(long)b * a.ss_ld + a.C + c_first];
    }
  }
  const float2* st = gn_group_stats(a, b);
  const int cpg = a.C / a.G;
  for (int c = c_first; c < a.C; c += blockDim.x) {
    if (c != c_first) {
      ga = a.gamma[c];
      be = a.beta[c];
      if (a.ss) {
        sc = a.ss[(long)b * a.ss_ld + c] + 1.f;
        sh = a.ss[(long)b * a.ss_ld + a.C + c];
      }
    }
    const float2 mr = st[c / cpg];
    const float A = mr.y * ga;
    const float Bc = be - mr.x * A;
    sA[c] = A * sc;
    sB[c] = Bc * sc + sh;
  }
  return st;
}

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(q.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(q.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(q.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(q.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = q;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(q.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(q.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(q.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(q.w); v[6] = f.x; v[7] = f.y;
}
// 8 consecutive per-channel coefficients from shared memory (two LDS.128)
__device__ __forceinline__ void load_coef8(const float* sp, int c0, float (&v)[8]) {
  const float4 lo = *reinterpret_cast<const float4*>(sp + c0), hi = *reinterpret_cast<const float4*>(sp + c0 + 4);
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
  v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}
__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// All kernels below follow one latency plan: a thread first ISSUES the 16-byte loads of every vector it
// will process (fully unrolled, addresses do not depend on the statistics), then builds the per-sample
// affine coefficients in shared memory (a second, independent chain of global loads), synchronises, and
// only then consumes the data. The two DRAM round trips overlap, and a launch is one wave of short CTAs.
constexpr int kVecPerThread = 4;

// ---------------------------------------------------------------------------------------
// out = silu(GN(x) [* (scale+1) + shift])            (Block, modules.py:171-179)
// grid (ceil(nvec / (256*4)), B); each thread owns 4 8-channel vectors of its sample.
// ---------------------------------------------------------------------------------------
template <int V>  // 8-channel vectors per thread
__global__ void __launch_bounds__(kNormThreads) gn_silu_fwd_kernel(const GnArgs a, bf16* __restrict__ out) {
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  float* sA = sm;
  float* sB = sm + a.C;
  const int b = blockIdx.y;
  const int c8n = a.C / 8;
  const long nvec = (long)a.rows * c8n;
  const bf16* xb = a.x + (long)b * a.rows * a.C;
  bf16* ob = out + (long)b * a.rows * a.C;
  const long i0 = (long)blockIdx.x * (kNormThreads * V) + threadIdx.x;
  uint4 raw[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const long i = i0 + k * kNormThreads;
    raw[k] = i < nvec ? ldg16(xb + i * 8) : make_uint4(0, 0, 0, 0);
  }
  gn_affine_to_smem(a, b, sA, sB);
  __syncthreads();
  // when the block size is a multiple of the vectors per pixel a thread always owns the same 8 channels:
  // their coefficients are read from shared memory once
  const bool fixed_c = (kNormThreads % c8n) == 0;
  float cA[8], cB[8];
  load_coef8(sA, (int)(i0 % c8n) * 8, cA);
  load_coef8(sB, (int)(i0 % c8n) * 8, cB);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const long i = i0 + k * kNormThreads;
    if (i >= nvec) break;
    if (!fixed_c) {
      load_coef8(sA, (int)(i % c8n) * 8, cA);
      load_coef8(sB, (int)(i % c8n) * 8, cB);
    }
    float v[8];
    unpack8(raw[k], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = silu_f(fmaf(v[j], cA[j], cB[j]));
    store8(ob + i * 8, v);
  }
}

// ---------------------------------------------------------------------------------------
// ResnetBlock tail (modules.py:241-242): out = silu(GN(b_raw)) + LayerNorm_C(s)
// One lane group of LG = min(32, C/8) lanes per pixel; each lane owns 8*VPL channels; R pixels per thread.
// ---------------------------------------------------------------------------------------
template <int VPL, int R>
__global__ void __launch_bounds__(kNormThreads) resblock_tail_fwd_kernel(const GnArgs a, const bf16* __restrict__ s,
                                                                         const float* __restrict__ ln_g,
                                                                         const float* __restrict__ ln_b,
                                                                         bf16* __restrict__ out) {
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  float* sA = sm;
  float* sB = sm + a.C;
  float* sG = sm + 2 * a.C;
  float* sBt = sm + 3 * a.C;
  const int b = blockIdx.y;
  const int lg = a.C / (8 * VPL);  // lanes per pixel (power of two <= 32)
  const int sub = threadIdx.x % lg;
  const int ppb = blockDim.x / lg;  // pixels per block pass
  const long base = (long)b * a.rows;
  uint4 sraw[R][VPL], xraw[R][VPL];
  long roff[R];
  bool valid[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long p_raw = ((long)blockIdx.x * R + r) * ppb + threadIdx.x / lg;
    valid[r] = p_raw < a.rows;
    roff[r] = (base + (valid[r] ? p_raw : (long)a.rows - 1)) * a.C;  // clamped: the shuffles need whole warps
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      sraw[r][k] = ldg16(s + roff[r] + (k * lg + sub) * 8);
      xraw[r][k] = ldg16(a.x + roff[r] + (k * lg + sub) * 8);
    }
  }
  gn_affine_to_smem(a, b, sA, sB);
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    sG[c] = ln_g[c];
    sBt[c] = ln_b[c];
  }
  __syncthreads();
  const float inv_c = 1.f / (float)a.C;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float sv[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      unpack8(sraw[r][k], sv[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1 += sv[k][j];
        s2 += sv[k][j] * sv[k][j];
      }
    }
    for (int o = lg >> 1; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mean = s1 * inv_c;
    const float rstd = rsqrtf(fmaxf(s2 * inv_c - mean * mean, 0.f) + kEps);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c0 = (k * lg + sub) * 8;
      float xv[8], o8[8], cA[8], cB[8], cG[8], cBt[8];
      load_coef8(sA, c0, cA);
      load_coef8(sB, c0, cB);
      load_coef8(sG, c0, cG);
      load_coef8(sBt, c0, cBt);
      unpack8(xraw[r][k], xv);
      const float nmr = -mean * rstd;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = silu_f(fmaf(xv[j], cA[j], cB[j]));
        o8[j] = y + fmaf(fmaf(sv[k][j], rstd, nmr), cG[j], cBt[j]);
      }
      if (valid[r]) store8(out + roff[r] + c0, o8);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Backward.
//   z = x*A + Bc ; y = silu(z) ; dz = dy * silu'(z)
//   T1[b,c] = sum_pix dz ; T2[b,c] = sum_pix dz * xhat      (xhat = (x - mean_g) * rstd_g)
// grid (ceil(rows / (pixel lanes * 4)), B) in clusters along x: 4 pixels per thread, block partials are summed
// across the cluster through DSMEM before the atomics into T.
// ---------------------------------------------------------------------------------------
template <bool kMulti>
__global__ void __launch_bounds__(kNormThreads, 2) gn_bwd_reduce_kernel(const GnArgs a, const bf16* __restrict__ dy,
                                                                        float* __restrict__ T /*[B][C][2]*/, int iters_arg) {
  const int iters = kMulti ? iters_arg : 1;
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  float* sA = sm;
  float* sB = sm + a.C;
  float* sMean = sm + 2 * a.C;   // per channel (group value replicated)
  float* sRstd = sm + 3 * a.C;
  float* vals = sm + 4 * a.C;    // [C][2] block totals
  float* red = sm + 6 * a.C;     // [blockDim][16]
  const int b = blockIdx.y;
  const int c8n = a.C / 8;              // vectors per pixel
  const int pl_n = blockDim.x / c8n;    // pixel lanes (>= 1 when C <= 2048)
  const int ci = threadIdx.x % c8n;
  const int pl = threadIdx.x / c8n;
  const int c0 = ci * 8;
  const long base = (long)b * a.rows;
  // A block owns `iters` consecutive chunks of kVecPerThread * pl_n pixels (large samples: the statistics
  // prologue and the block-level reduction + atomics are amortised over `iters` times the data); two resident
  // blocks per SM overlap one block's loads with the other's arithmetic.
  uint4 xraw[kVecPerThread], draw[kVecPerThread];
  bool valid[kVecPerThread];
  auto issue = [&](int it, uint4 (&xr)[kVecPerThread], uint4 (&dr)[kVecPerThread], bool (&vl)[kVecPerThread]) {
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const long p = (((long)blockIdx.x * iters + it) * kVecPerThread + k) * pl_n + pl;
      vl[k] = pl < pl_n && p < a.rows;
      if (vl[k]) {
        const long off = (base + p) * a.C + c0;
        xr[k] = ldg16(a.x + off);
        dr[k] = ldg16(dy + off);
      }
    }
  };
  issue(0, xraw, draw, valid);
  {
    const float2* st = gn_affine_to_smem(a, b, sA, sB);
    const int cpg = a.C / a.G;
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      const float2 mr = st[c / cpg];
      sRstd[c] = mr.y;
      sMean[c] = -mr.x * mr.y;  // xhat = x * rstd + (-mean * rstd)
    }
  }
  __syncthreads();
  float t1[8], t2[8], cA[8], cB[8], cR[8], cM[8];
  load_coef8(sA, c0, cA);
  load_coef8(sB, c0, cB);
  load_coef8(sRstd, c0, cR);
  load_coef8(sMean, c0, cM);
#pragma unroll
  for (int j = 0; j < 8; ++j) t1[j] = t2[j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      if (!valid[k]) continue;
      float xv[8], dv[8];
      unpack8(xraw[k], xv);
      unpack8(draw[k], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(xv[j], cA[j], cB[j]);
        const float dz = dv[j] * silu_grad_f(z);
        const float xh = fmaf(xv[j], cR[j], cM[j]);
        t1[j] += dz;
        t2[j] = fmaf(dz, xh, t2[j]);
      }
    }
    if (kMulti && it + 1 < iters) issue(it + 1, xraw, draw, valid);
  }
  // reduce over pixel lanes through smem, over the cluster through DSMEM, then one atomic per (c, {T1,T2})
  float* my = red + (long)threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    my[j] = t1[j];
    my[8 + j] = t2[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c8n * 16; i += blockDim.x) {
    const int cc = i / 16, j = i % 16;
    float acc = 0.f;
    for (int q = 0; q < pl_n; ++q) acc += red[(long)(q * c8n + cc) * 16 + j];
    vals[(cc * 8 + (j & 7)) * 2 + (j >> 3)] = acc;
  }
  cluster_reduce_atomic_add(vals, 2 * a.C, T + (long)b * a.C * 2);
}

// dx = rstd_g * (gamma_c*(s+1)*dz - m1_g - xhat*m2_g),  m1_g = mean_g(dxhat), m2_g = mean_g(dxhat*xhat)
// The x == 0 block of each sample also finalises dgamma / dbeta (atomics over samples) and (dscale | dshift);
// every block accumulates the column sums of dx = the gradient of the producing conv's bias (cluster-reduced).
template <bool kMulti>
__global__ void __launch_bounds__(kNormThreads, 2) gn_bwd_apply_kernel(const GnArgs a, const bf16* __restrict__ dy,
                                                                       const float* __restrict__ T,
                                                                       bf16* __restrict__ dx, float* __restrict__ dgamma,
                                                                       float* __restrict__ dbeta, float* __restrict__ dss,
                                                                       int dss_ld, float* __restrict__ dconv_bias,
                                                                       int iters_arg) {
  const int iters = kMulti ? iters_arg : 1;
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  float* sA = sm;
  float* sB = sm + a.C;
  float* sMean = sm + 2 * a.C;
  float* sRstd = sm + 3 * a.C;
  float* sK = sm + 4 * a.C;    // gamma*(s+1)
  float* sM1 = sm + 5 * a.C;   // per channel copy of m1_g
  float* sM2 = sm + 6 * a.C;
  float* vals = sm + 7 * a.C;  // [C] block column sums of dx
  float* red = sm + 8 * a.C;   // [blockDim][8]
  const int b = blockIdx.y;
  const int cpg = a.C / a.G;
  const int c8n = a.C / 8;
  const long nvec = (long)a.rows * c8n;
  const long boff = (long)b * a.rows * a.C;
  // a block owns `iters` consecutive chunks of kNormThreads * kVecPerThread vectors (see gn_bwd_reduce_kernel)
  const long i0 = (long)blockIdx.x * iters * (kNormThreads * kVecPerThread) + threadIdx.x;
  uint4 xraw[kVecPerThread], draw[kVecPerThread];
  auto issue = [&](int it, uint4 (&xr)[kVecPerThread], uint4 (&dr)[kVecPerThread]) {
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const long i = i0 + ((long)it * kVecPerThread + k) * kNormThreads;
      if (i < nvec) {
        xr[k] = ldg16(a.x + boff + i * 8);
        dr[k] = ldg16(dy + boff + i * 8);
      }
    }
  };
  issue(0, xraw, draw);
  const float2* st = gn_affine_to_smem(a, b, sA, sB);
  const float inv_n = 1.f / ((float)a.rows * (float)cpg);
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const float2 mr = st[c / cpg];
    const float rstd = mr.y;
    sRstd[c] = rstd;
    sMean[c] = -mr.x * rstd;  // xhat = x * rstd + (-mean * rstd)
    const float sc = a.ss ? a.ss[(long)b * a.ss_ld + c] + 1.f : 1.f;
    sK[c] = a.gamma[c] * sc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    float m1 = 0.f, m2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      m1 += sK[g0 + k] * T[((long)b * a.C + g0 + k) * 2];
      m2 += sK[g0 + k] * T[((long)b * a.C + g0 + k) * 2 + 1];
    }
    sM1[c] = -m1 * inv_n * sRstd[c];  // pre-multiplied by rstd (and negated) for the FMA chain below
    sM2[c] = -m2 * inv_n * sRstd[c];
    if (blockIdx.x == 0) {  // finalize (once per sample)
      const float t1 = T[((long)b * a.C + c) * 2], t2 = T[((long)b * a.C + c) * 2 + 1];
      const float sc = a.ss ? a.ss[(long)b * a.ss_ld + c] + 1.f : 1.f;
      atomicAdd(&dgamma[c], sc * t2);
      atomicAdd(&dbeta[c], sc * t1);
      if (dss) {
        dss[(long)b * dss_ld + c] = a.gamma[c] * t2 + a.beta[c] * t1;  // dscale = sum dz * (xhat*gamma + beta)
        dss[(long)b * dss_ld + a.C + c] = t1;                          // dshift
      }
    }
  }
  __syncthreads();
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = 0.f;
  // blockDim (256) is a multiple of c8n, so a thread always owns the same 8 channels
  const int c0 = (int)(i0 % c8n) * 8;
  float cA[8], cB[8], cR[8], cM[8], cK[8], cM1[8], cM2[8];
  load_coef8(sA, c0, cA);
  load_coef8(sB, c0, cB);
  load_coef8(sRstd, c0, cR);
  load_coef8(sMean, c0, cM);
  load_coef8(sK, c0, cK);
  load_coef8(sM1, c0, cM1);
  load_coef8(sM2, c0, cM2);
#pragma unroll
  for (int j = 0; j < 8; ++j) cK[j] *= cR[j];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kVecPerThread; ++k) {
      const long i = i0 + ((long)it * kVecPerThread + k) * kNormThreads;
      if (i >= nvec) break;
      float xv[8], dv[8], o[8];
      unpack8(xraw[k], xv);
      unpack8(draw[k], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(xv[j], cA[j], cB[j]);
        const float dz = dv[j] * silu_grad_f(z);
        const float xh = fmaf(xv[j], cR[j], cM[j]);
        o[j] = fmaf(xh, cM2[j], fmaf(dz, cK[j], cM1[j]));  // rstd * (K dz - m1 - xhat m2)
        bs[j] += o[j];
      }
      store8(dx + boff + i * 8, o);
    }
    if (kMulti && it + 1 < iters) issue(it + 1, xraw, draw);
  }
  if (dconv_bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = bs[j];
    __syncthreads();
    const int per = blockDim.x / c8n;  // threads sharing a channel vector
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      float acc = 0.f;
      for (int q = 0; q < per; ++q) acc += red[(q * c8n + c / 8) * 8 + (c & 7)];
      vals[c] = acc;
    }
    cluster_reduce_atomic_add(vals, a.C, dconv_bias);
  }
}

// ---------------------------------------------------------------------------------------
// GroupNorm + SiLU backward in ONE launch without any grid-level exchange: the sums a backward needs couple only the
// channels of one GROUP of one sample, so a thread-block cluster that owns a whole (sample, channel slice) - max(8, cpg)
// channels = one or two groups, all rows, split by rows over the CTAs of the cluster - can do both passes by itself:
// pass 1 reads x and dy once (dz = dy * silu'(z) goes to shared memory as fp32; x stays in registers as raw bf16 when a
// thread owns at most 3 vectors, else it is read again in pass 2 - an L2 hit), accumulates T1 / T2 per channel, the block
// reduces them (shuffles over the lanes that own the same 8 channels, warp partials through shared memory), the CTAs of
// the cluster add each other's 64 floats through DSMEM, and pass 2 turns dz into dx with three FMAs per element.
// No T round trip through L2, no second launch, no barrier between clusters; 128 CTAs at every level of config_v2_2.
// ---------------------------------------------------------------------------------------
constexpr int kGrpThreads = 512;
constexpr int kGrpMaxSmem = 200 * 1024;

template <int kMaxVec, bool kXRegs>
__global__ void __launch_bounds__(kGrpThreads, 1) gn_bwd_group_kernel(const GnArgs a, const bf16* __restrict__ dy,
                                                                      float* __restrict__ T, bf16* __restrict__ dx,
                                                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                      float* __restrict__ dss, int dss_ld,
                                                                      float* __restrict__ dconv_bias) {
  extern __shared__ __align__(16) float s_dz[];  // [nvec][8]
  __shared__ float2 s_stat[2];
  __shared__ __align__(16) float sA[32], sB[32], sK[32], sSc[32], sR[32], sNm[32], sM1[32], sM2[32];
  __shared__ float sT[2][32];
  __shared__ __align__(16) float sRed[kGrpThreads / 32][4][16];
  const int b = blockIdx.y;
  const int cpg = a.C / a.G;
  const int cw = cpg < 8 ? 8 : cpg;        // channels of this slice: one group, or two groups of 4
  const int vpr = cw >> 3, c0s = blockIdx.x * cw;
  const int n_cl = (int)gridDim.z;         // CTAs per (sample, slice) = cluster size: row ranges
  const int rows_cta = (a.rows + n_cl - 1) / n_cl;
  const int row_lo = blockIdx.z * rows_cta;
  const int my_rows = max(0, min(a.rows, row_lo + rows_cta) - row_lo);
  const long nvec = (long)my_rows * vpr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int v = tid % vpr;  // kGrpThreads is a multiple of vpr: a thread always owns the same 8 channels
  const int c0 = c0s + v * 8;
  const long base = ((long)b * a.rows + row_lo) * a.C;
  pdl_trigger();
  pdl_wait();
  uint4 xraw[kXRegs ? kMaxVec : 1];
  // pass 1 in batches of up to 4 vectors per thread: all loads of a batch are in flight before the first is consumed
  if (warp == 0) {  // (mean, rstd) of the slice's group(s): lanes 0-15 / 16-31 add up the 16 replicas of group 0 / 1
    const int gi = lane >> 4, g = c0s / cpg + gi;
    float s1 = 0.f, s2 = 0.f;
    if (gi * cpg < cw) {
      const float2 q = *reinterpret_cast<const float2*>(a.sums + ((long)((lane & 15) * a.B + b) * a.G + g) * 2);
      s1 = q.x;
      s2 = q.y;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((lane & 15) == 0) {
      const float inv_n = 1.f / ((float)a.rows * (float)cpg);
      const float mean = s1 * inv_n;
      s_stat[gi] = make_float2(mean, rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + kEps));
    }
  }
  __syncthreads();
  if (tid < cw) {
    const int c = c0s + tid;
    const float2 mr = s_stat[tid / cpg];
    const float ga = a.gamma[c], be = a.beta[c];
    const float sc = a.ss ? a.ss[(long)b * a.ss_ld + c] + 1.f : 1.f;
    const float sh = a.ss ? a.ss[(long)b * a.ss_ld + a.C + c] : 0.f;
    const float A = mr.y * ga;
    sA[tid] = A * sc;
    sB[tid] = (be - mr.x * A) * sc + sh;
    sK[tid] = ga * sc;
    sSc[tid] = sc;
    sR[tid] = mr.y;
    sNm[tid] = -mr.x * mr.y;
  }
  __syncthreads();
  float cA[8], cB[8], cR[8], cN[8], t1[8], t2[8];
  load_coef8(sA, v * 8, cA);
  load_coef8(sB, v * 8, cB);
  load_coef8(sR, v * 8, cR);
  load_coef8(sNm, v * 8, cN);
#pragma unroll
  for (int j = 0; j < 8; ++j) t1[j] = t2[j] = 0.f;
  constexpr int kBatch = kMaxVec < 4 ? kMaxVec : 4;
#pragma unroll
  for (int k0 = 0; k0 < kMaxVec; k0 += kBatch) {
    uint4 xb[kBatch], db[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const long i = tid + (long)(k0 + q) * kGrpThreads;
      if (k0 + q < kMaxVec && i < nvec) {
        const long off = base + (i / vpr) * a.C + c0;
        xb[q] = ldg16(a.x + off);
        db[q] = ldg16(dy + off);
      }
    }
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const long i = tid + (long)(k0 + q) * kGrpThreads;
      if (k0 + q < kMaxVec && i < nvec) {
        if (kXRegs) xraw[k0 + q] = xb[q];
        float xv[8], dv[8], dz[8];
        unpack8(xb[q], xv);
        unpack8(db[q], dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(xv[j], cA[j], cB[j]);
          dz[j] = dv[j] * silu_grad_f(z);
          const float xh = fmaf(xv[j], cR[j], cN[j]);
          t1[j] += dz[j];
          t2[j] = fmaf(dz[j], xh, t2[j]);
        }
        float4* dp = reinterpret_cast<float4*>(s_dz + i * 8);
        dp[0] = make_float4(dz[0], dz[1], dz[2], dz[3]);
        dp[1] = make_float4(dz[4], dz[5], dz[6], dz[7]);
      }
    }
  }
  // lanes with the same v hold partials of the same 8 channels: xor-reduce over them, then over the warps
  for (int o = vpr; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      t1[j] += __shfl_xor_sync(0xffffffffu, t1[j], o);
      t2[j] += __shfl_xor_sync(0xffffffffu, t2[j], o);
    }
  }
  if (lane < vpr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sRed[warp][lane][j] = t1[j];
      sRed[warp][lane][8 + j] = t2[j];
    }
  }
  __syncthreads();
  if (tid < 2 * cw) {
    const int which = tid / cw, cc = tid - which * cw;
    float acc = 0.f;
#pragma unroll 8
    for (int w = 0; w < kGrpThreads / 32; ++w) acc += sRed[w][cc >> 3][which * 8 + (cc & 7)];
    sT[which][cc] = acc;
  }
  // the row ranges of the cluster's CTAs meet here: every CTA adds its peers' per-channel partials (<= 64 floats through
  // DSMEM) and ends up with the sums of the whole slice
  cluster_sync_all();
  float tot = 0.f;
  if (tid < 2 * cw) {
    const float* mine = &sT[0][0] + (tid / cw) * 32 + (tid % cw);
    for (uint32_t r = 0; r < (uint32_t)n_cl; ++r) tot += dsmem_ld_f32(mine, r);
  }
  cluster_sync_all();  // every peer has read this CTA's partials
  if (tid < 2 * cw) sT[tid / cw][tid % cw] = tot;
  __syncthreads();
  if (warp == 0) {
    // group means: segmented sum over the cpg lanes of each group (cpg is a power of two >= 4)
    float m1 = lane < cw ? sK[lane] * sT[0][lane] : 0.f;
    float m2 = lane < cw ? sK[lane] * sT[1][lane] : 0.f;
    for (int o = 1; o < cpg && o < 32; o <<= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    if (lane < cw) {
      const float inv_n = 1.f / ((float)a.rows * (float)cpg);
      sM1[lane] = -m1 * inv_n * sR[lane];  // pre-multiplied by rstd (and negated) for the FMA chain below
      sM2[lane] = -m2 * inv_n * sR[lane];
      if (blockIdx.z == 0) {  // per-channel results of this (sample, slice): once per cluster
        const int c = c0s + lane;
        const float u1 = sT[0][lane], u2 = sT[1][lane], sc = sSc[lane];
        T[((long)b * a.C + c) * 2] = u1;
        T[((long)b * a.C + c) * 2 + 1] = u2;
        atomicAdd(&dgamma[c], sc * u2);
        atomicAdd(&dbeta[c], sc * u1);
        if (dss) {
          dss[(long)b * dss_ld + c] = a.gamma[c] * u2 + a.beta[c] * u1;  // dscale = sum dz * (xhat*gamma + beta)
          dss[(long)b * dss_ld + a.C + c] = u1;                          // dshift
        }
      }
    }
  }
  __syncthreads();
  float cK[8], cM1[8], cM2[8], bs[8];
  load_coef8(sK, v * 8, cK);
  load_coef8(sM1, v * 8, cM1);
  load_coef8(sM2, v * 8, cM2);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    cK[j] *= cR[j];
    bs[j] = 0.f;
  }
#pragma unroll
  for (int k0 = 0; k0 < kMaxVec; k0 += kBatch) {
    uint4 xb[kBatch];
    if (!kXRegs) {
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const long i = tid + (long)(k0 + q) * kGrpThreads;
        if (k0 + q < kMaxVec && i < nvec) xb[q] = ldg16(a.x + base + (i / vpr) * a.C + c0);
      }
    }
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const long i = tid + (long)(k0 + q) * kGrpThreads;
      if (k0 + q < kMaxVec && i < nvec) {
        float xv[8], o[8];
        unpack8(kXRegs ? xraw[k0 + q] : xb[q], xv);
        const float4* dp = reinterpret_cast<const float4*>(s_dz + i * 8);
        const float4 d0 = dp[0], d1 = dp[1];
        const float dz[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = fmaf(xv[j], cR[j], cN[j]);
          o[j] = fmaf(xh, cM2[j], fmaf(dz[j], cK[j], cM1[j]));  // rstd * (K dz - m1 - xhat m2)
          bs[j] += o[j];
        }
        store8(dx + base + (i / vpr) * a.C + c0, o);
      }
    }
  }
  if (dconv_bias) {  // column sums of dx = the gradient of the producing conv's bias
    for (int o = vpr; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], o);
    }
    __syncthreads();  // sRed is reused
    if (lane < vpr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sRed[warp][lane][j] = bs[j];
    }
    __syncthreads();
    if (tid < cw) {
      float acc = 0.f;
#pragma unroll 8
      for (int w = 0; w < kGrpThreads / 32; ++w) acc += sRed[w][tid >> 3][tid & 7];
      atomicAdd(&dconv_bias[c0s + tid], acc);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Single-launch GroupNorm+SiLU backward for samples whose (x, dy) slices fit in the shared memory of the GPU's SMs
// (every level of config_v2_2): one CTA per SM, gridDim.x CTAs per sample. A CTA streams its slice of x and dy from
// HBM ONCE into shared memory while accumulating T1/T2, the CTAs of a sample meet at a per-sample barrier (global
// counters; all CTAs are co-resident by construction: grid <= number of SMs, one CTA per SM), then dx is computed from
// the shared-memory copy. Against the two-kernel version: 3 instead of 5 tensor passes over HBM/L2, one launch, one
// statistics prologue. pdl_trigger() is issued only AFTER the barrier: a dependent grid that started early could
// otherwise take the SM a not-yet-scheduled CTA of this grid needs, and the barrier would never complete.
// ---------------------------------------------------------------------------------------
constexpr int kFusedThreads = 512;
constexpr int kFusedSlots = 64;
constexpr int kFusedMaxB = 148;
__device__ unsigned g_gn_barrier[kFusedSlots][kFusedMaxB][2];  // [launch slot][sample][arrived, left]; self-resetting

__global__ void __launch_bounds__(kFusedThreads, 1) gn_bwd_fused_kernel(const GnArgs a, const bf16* __restrict__ dy,
                                                                       float* __restrict__ T, bf16* __restrict__ dx,
                                                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                       float* __restrict__ dss, int dss_ld,
                                                                       float* __restrict__ dconv_bias, int rows_per_cta,
                                                                       int slot) {
  extern __shared__ __align__(16) float sm[];
  pdl_wait();
  float* sA = sm;
  float* sB = sm + a.C;
  float* sMean = sm + 2 * a.C;
  float* sRstd = sm + 3 * a.C;
  float* sK = sm + 4 * a.C;
  float* sM1 = sm + 5 * a.C;
  float* sM2 = sm + 6 * a.C;
  float* vals = sm + 7 * a.C;                 // [2C]
  float* red = sm + 9 * a.C;                  // [blockDim][16]
  uint4* sx = reinterpret_cast<uint4*>(red + kFusedThreads * 16);
  const int b = blockIdx.y;
  const int cpg = a.C / a.G;
  const int c8n = a.C / 8;
  const int row0 = blockIdx.x * rows_per_cta;
  const int nrows = max(0, min(a.rows - row0, rows_per_cta));
  const int nvec = nrows * c8n;
  uint4* sd = sx + (long)rows_per_cta * c8n;
  const long v0 = ((long)b * a.rows + row0) * c8n;  // first vector of the slice
  const int tid = threadIdx.x;
  const int c0 = (tid % c8n) * 8;               // blockDim is a multiple of c8n: a thread always owns the same 8 channels

  const float2* st = gn_affine_to_smem(a, b, sA, sB);
  for (int c = tid; c < a.C; c += blockDim.x) {
    const float2 mr = st[c / cpg];
    sRstd[c] = mr.y;
    sMean[c] = -mr.x * mr.y;  // xhat = x * rstd + (-mean * rstd)
    const float sc = a.ss ? a.ss[(long)b * a.ss_ld + c] + 1.f : 1.f;
    sK[c] = a.gamma[c] * sc;
  }
  __syncthreads();
  float cA[8], cB[8], cR[8], cM[8];
  load_coef8(sA, c0, cA);
  load_coef8(sB, c0, cB);
  load_coef8(sRstd, c0, cR);
  load_coef8(sMean, c0, cM);
  // ---- phase 1: HBM -> shared memory, T1 / T2 partial sums ----
  float t1[8], t2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t1[j] = t2[j] = 0.f;
  for (int i0 = tid; i0 < nvec; i0 += kFusedThreads * 4) {
    uint4 xr[4], dr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k * kFusedThreads;
      if (i < nvec) {
        xr[k] = ldg16(a.x + (v0 + i) * 8);
        dr[k] = ldg16(dy + (v0 + i) * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k * kFusedThreads;
      if (i >= nvec) break;
      sx[i] = xr[k];
      sd[i] = dr[k];
      float xv[8], dv[8];
      unpack8(xr[k], xv);
      unpack8(dr[k], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(xv[j], cA[j], cB[j]);
        const float dz = dv[j] * silu_grad_f(z);
        const float xh = fmaf(xv[j], cR[j], cM[j]);
        t1[j] += dz;
        t2[j] = fmaf(dz, xh, t2[j]);
      }
    }
  }
  {
    float* my = red + (long)tid * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      my[j] = t1[j];
      my[8 + j] = t2[j];
    }
  }
  __syncthreads();
  const int per = blockDim.x / c8n;  // threads sharing a channel vector
  for (int i = tid; i < c8n * 16; i += blockDim.x) {
    const int cc = i / 16, j = i % 16;
    float acc = 0.f;
    for (int q = 0; q < per; ++q) acc += red[(long)(q * c8n + cc) * 16 + j];
    atomicAdd(T + ((long)b * a.C + cc * 8 + (j & 7)) * 2 + (j >> 3), acc);
  }
  // ---- barrier over the CTAs of this sample ----
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned* bar = g_gn_barrier[slot][b];
    atomicAdd(&bar[0], 1u);
    while (*reinterpret_cast<volatile unsigned*>(&bar[0]) < gridDim.x) {
    }
    __threadfence();
    if (atomicAdd(&bar[1], 1u) == gridDim.x - 1) {  // last one out resets the counters for the slot's next use
      bar[0] = 0u;
      bar[1] = 0u;
      __threadfence();
    }
  }
  __syncthreads();
  pdl_trigger();
  // ---- phase 2: dx from the shared-memory copy ----
  const float inv_n = 1.f / ((float)a.rows * (float)cpg);
  for (int c = tid; c < a.C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    float m1 = 0.f, m2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      m1 += sK[g0 + k] * __ldcg(T + ((long)b * a.C + g0 + k) * 2);
      m2 += sK[g0 + k] * __ldcg(T + ((long)b * a.C + g0 + k) * 2 + 1);
    }
    sM1[c] = -m1 * inv_n * sRstd[c];
    sM2[c] = -m2 * inv_n * sRstd[c];
    if (blockIdx.x == 0) {  // finalize (once per sample)
      const float tt1 = __ldcg(T + ((long)b * a.C + c) * 2), tt2 = __ldcg(T + ((long)b * a.C + c) * 2 + 1);
      const float sc = a.ss ? a.ss[(long)b * a.ss_ld + c] + 1.f : 1.f;
      atomicAdd(&dgamma[c], sc * tt2);
      atomicAdd(&dbeta[c], sc * tt1);
      if (dss) {
        dss[(long)b * dss_ld + c] = a.gamma[c] * tt2 + a.beta[c] * tt1;
        dss[(long)b * dss_ld + a.C + c] = tt1;
      }
    }
  }
  __syncthreads();
  float cK[8], cM1[8], cM2[8], bs[8];
  load_coef8(sK, c0, cK);
  load_coef8(sM1, c0, cM1);
  load_coef8(sM2, c0, cM2);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    cK[j] *= cR[j];
    bs[j] = 0.f;
  }
#pragma unroll 2
  for (int i = tid; i < nvec; i += kFusedThreads) {
    float xv[8], dv[8], o[8];
    unpack8(sx[i], xv);
    unpack8(sd[i], dv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(xv[j], cA[j], cB[j]);
      const float dz = dv[j] * silu_grad_f(z);
      const float xh = fmaf(xv[j], cR[j], cM[j]);
      o[j] = fmaf(xh, cM2[j], fmaf(dz, cK[j], cM1[j]));  // rstd * (K dz - m1 - xhat m2)
      bs[j] += o[j];
    }
    store8(dx + (v0 + i) * 8, o);
  }
  if (dconv_bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[tid * 8 + j] = bs[j];
    __syncthreads();
    for (int c = tid; c < a.C; c += blockDim.x) {
      float acc = 0.f;
      for (int q = 0; q < per; ++q) acc += red[(q * c8n + c / 8) * 8 + (c & 7)];
      atomicAdd(dconv_bias + c, acc);
    }
  }
}

// LayerNorm over channels, backward (norm_2 of ResnetBlock, modules.py:223,242):
//   y = shat*g + b ; ds = rstd * (g*dy - mean_c(g*dy) - shat*mean_c(g*dy*shat)) ; dg += dy*shat ; db += dy
// R pixels per thread; block column sums (dg | db) are cluster-reduced before the global atomics.
template <int VPL, int R>
__global__ void __launch_bounds__(kNormThreads) ln_bwd_kernel(const bf16* __restrict__ s, const bf16* __restrict__ dy,
                                                              const float* __restrict__ ln_g, bf16* __restrict__ ds,
                                                              float* __restrict__ dg, float* __restrict__ db, long P,
                                                              int C, int iters) {
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  float* sG = sm;            // [C]
  float* acc = sm + C;       // [2C] block-level accumulators: dg | db
  const int lg = C / (8 * VPL);
  const int sub = threadIdx.x % lg;
  const int ppb = blockDim.x / lg;
  uint4 sraw[R][VPL], draw[R][VPL];
  long roff[R];
  bool valid[R];
  // a block owns `iters` consecutive chunks of R * ppb pixels (large tensors: the column-sum reduction and its atomics
  // are amortised over `iters` times the data)
  auto issue = [&](int it) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long p_raw = (((long)blockIdx.x * iters + it) * R + r) * ppb + threadIdx.x / lg;
      valid[r] = p_raw < P;
      roff[r] = (valid[r] ? p_raw : P - 1) * C;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        sraw[r][k] = ldg16(s + roff[r] + (k * lg + sub) * 8);
        draw[r][k] = ldg16(dy + roff[r] + (k * lg + sub) * 8);
      }
    }
  };
  issue(0);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sG[c] = ln_g[c];
    acc[c] = 0.f;
    acc[C + c] = 0.f;
  }
  __syncthreads();
  const float inv_c = 1.f / (float)C;
  float pg[VPL][8], pb[VPL][8];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) pg[k][j] = pb[k][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float sv[VPL][8], dv[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      unpack8(sraw[r][k], sv[k]);
      unpack8(draw[r][k], dv[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!valid[r]) dv[k][j] = 0.f;
        s1 += sv[k][j];
        s2 += sv[k][j] * sv[k][j];
      }
    }
    for (int o = lg >> 1; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mean = s1 * inv_c;
    const float rstd = rsqrtf(fmaxf(s2 * inv_c - mean * mean, 0.f) + kEps);
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c0 = (k * lg + sub) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sh = (sv[k][j] - mean) * rstd;
        const float gd = sG[c0 + j] * dv[k][j];
        a1 += gd;
        a2 += gd * sh;
        pg[k][j] += dv[k][j] * sh;
        pb[k][j] += dv[k][j];
        sv[k][j] = sh;  // keep shat
      }
    }
    for (int o = lg >> 1; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    a1 *= inv_c;
    a2 *= inv_c;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c0 = (k * lg + sub) * 8;
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = rstd * (sG[c0 + j] * dv[k][j] - a1 - sv[k][j] * a2);
      if (valid[r]) store8(ds + roff[r] + c0, o8);
    }
  }
    if (it + 1 < iters) issue(it + 1);
  }
  // lanes that own the same channels (same `sub`, other pixels of the warp) are summed with shuffles first
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c0 = (k * lg + sub) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = pg[k][j], bb = pb[k][j];
      for (int o = 16; o >= lg; o >>= 1) {
        g += __shfl_xor_sync(0xffffffffu, g, o);
        bb += __shfl_xor_sync(0xffffffffu, bb, o);
      }
      if ((threadIdx.x & 31) < lg) {
        atomicAdd(&acc[c0 + j], g);
        atomicAdd(&acc[C + c0 + j], bb);
      }
    }
  }
  __syncthreads();
  // acc = (dg | db): cluster-level sum, then one atomic per channel
  const uint32_t nrank = cluster_nctarank();
  if (nrank > 1) cluster_sync_all();
  if (nrank == 1 || cluster_ctarank() == 0) {
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
      float v = acc[c];
      for (uint32_t r = 1; r < nrank; ++r) v += dsmem_ld_f32(acc + c, r);
      atomicAdd(c < C ? &dg[c] : &db[c - C], v);
    }
  }
  if (nrank > 1) cluster_sync_all();
}

// cluster width for a grid.x of `want` CTAs (power of two <= 8) and the padded grid.x
static int cluster_for(int want, int* grid_x) {
  // Measured on B200 (profiles/r1_microbench_sweep.txt): launching these short kernels as thread-block clusters costs
  // more (co-scheduling of 8 CTAs per GPC) than the DSMEM pre-reduction saves in L2 atomics - gn_silu_bwd 38.8 us with
  // clusters of 8, 30.5 us without. The cluster path stays selectable (VDN_NORM_CLUSTER=2|4|8) but is off by default.
  const int cmax = tune_int("VDN_NORM_CLUSTER", 1);
  int c = 1;
  while (c < cmax && c * 2 <= want) c *= 2;
  *grid_x = (want + c - 1) / c * c;
  return c;
}

static int vpl_for(int C) {
  // lanes per pixel = C / (8*VPL) must be a power of two <= 32
  int vpl = 1;
  while (C / (8 * vpl) > 32) vpl *= 2;
  return vpl;
}
static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static int check_gn(const char* who, int B, int rows, int C, int G) {
  VDN_REQUIRE(B > 0 && rows > 0 && C >= 8 && C % 8 == 0 && G > 0 && C % G == 0, VDN_E_SHAPE,
              "%s: bad shape B=%d rows=%d C=%d G=%d", who, B, rows, C, G);
  VDN_REQUIRE(C <= 2048, VDN_E_SHAPE, "%s: C=%d > 2048 unsupported", who, C);
  VDN_REQUIRE(G <= kMaxGroups, VDN_E_SHAPE, "%s: G=%d > %d groups unsupported", who, G, kMaxGroups);
  return VDN_OK;
}

static int grid_x_for(long work_items, int per_block) {
  return (int)std::max<long>(1, (work_items + per_block - 1) / per_block);
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_gn_silu_fwd(const void* x_raw, const float* gn_sums, const float* gamma, const float* beta,
                               const float* scale_shift, int ss_ld, void* out, int B, int rows_per_sample, int C,
                               int G, void* stream) {
  int rc = check_gn("gn_silu_fwd", B, rows_per_sample, C, G);
  if (rc) return rc;
  GnArgs a{reinterpret_cast<const bf16*>(x_raw), gn_sums, gamma, beta, scale_shift, ss_ld, B, rows_per_sample, C, G};
  const long nvec = (long)rows_per_sample * (C / 8);
  // large samples: 8 vectors per thread (the per-block statistics prologue is amortised over twice the data)
  const bool big = nvec >= 64 * 1024;
  dim3 grid(grid_x_for(nvec, kNormThreads * (big ? 8 : kVecPerThread)), B);
  cudaError_t le = big ? launch_pdl(gn_silu_fwd_kernel<8>, grid, dim3(kNormThreads), 2 * C * sizeof(float),
                                    reinterpret_cast<cudaStream_t>(stream), 1, a, reinterpret_cast<bf16*>(out))
                       : launch_pdl(gn_silu_fwd_kernel<kVecPerThread>, grid, dim3(kNormThreads), 2 * C * sizeof(float),
                                    reinterpret_cast<cudaStream_t>(stream), 1, a, reinterpret_cast<bf16*>(out));
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "gn_silu_fwd launch: %s", cudaGetErrorString(le));
  return check_launch("gn_silu_fwd");
}

extern "C" int vdn_resblock_tail_fwd(const void* b_raw, const float* gn_sums, const float* gamma, const float* beta,
                                     const void* s, const float* ln_gamma, const float* ln_beta, void* out, int B,
                                     int rows_per_sample, int C, int G, void* stream) {
  int rc = check_gn("resblock_tail_fwd", B, rows_per_sample, C, G);
  if (rc) return rc;
  const int vpl = vpl_for(C);
  VDN_REQUIRE(pow2(C / (8 * vpl)), VDN_E_SHAPE, "resblock_tail_fwd: C=%d must be 8 * power of two", C);
  GnArgs a{reinterpret_cast<const bf16*>(b_raw), gn_sums, gamma, beta, nullptr, 0, B, rows_per_sample, C, G};
  const int lg = C / (8 * vpl);
  const int ppb = kNormThreads / lg;
  const size_t smem = 4 * C * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bf16* sp = reinterpret_cast<const bf16*>(s);
  bf16* op = reinterpret_cast<bf16*>(out);
  cudaError_t le;
  switch (vpl) {  // R pixels per thread: 8 16-byte loads in flight per thread
    case 1: le = launch_pdl(resblock_tail_fwd_kernel<1, 4>, dim3(grid_x_for(rows_per_sample, ppb * 4), B), dim3(kNormThreads), smem, st, 1, a, sp, ln_gamma, ln_beta, op); break;
    case 2: le = launch_pdl(resblock_tail_fwd_kernel<2, 2>, dim3(grid_x_for(rows_per_sample, ppb * 2), B), dim3(kNormThreads), smem, st, 1, a, sp, ln_gamma, ln_beta, op); break;
    case 4: le = launch_pdl(resblock_tail_fwd_kernel<4, 1>, dim3(grid_x_for(rows_per_sample, ppb), B), dim3(kNormThreads), smem, st, 1, a, sp, ln_gamma, ln_beta, op); break;
    default: le = launch_pdl(resblock_tail_fwd_kernel<8, 1>, dim3(grid_x_for(rows_per_sample, ppb), B), dim3(kNormThreads), smem, st, 1, a, sp, ln_gamma, ln_beta, op); break;
  }
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "resblock_tail_fwd launch: %s", cudaGetErrorString(le));
  return check_launch("resblock_tail_fwd");
}

static int gn_silu_bwd_impl(const void* dy, const void* x_raw, const float* gn_sums, const float* gamma, const float* beta,
                            const float* scale_shift, int ss_ld, float* T_ws, void* dx_raw, float* dgamma, float* dbeta,
                            float* dss, int dss_ld, float* dconv_bias, int B, int rows_per_sample, int C, int G, void* stream,
                            bool zero_ws);

extern "C" int vdn_gn_silu_bwd(const void* dy, const void* x_raw, const float* gn_sums, const float* gamma,
                               const float* beta, const float* scale_shift, int ss_ld, float* T_ws, void* dx_raw,
                               float* dgamma, float* dbeta, float* dss, int dss_ld, float* dconv_bias, int B,
                               int rows_per_sample, int C, int G, void* stream) {
  return gn_silu_bwd_impl(dy, x_raw, gn_sums, gamma, beta, scale_shift, ss_ld, T_ws, dx_raw, dgamma, dbeta, dss, dss_ld,
                          dconv_bias, B, rows_per_sample, C, G, stream, true);
}
// Same with T_ws zeroed by the CALLER (one memset for all the layers of a step instead of a memset node per call).
extern "C" int vdn_gn_silu_bwd_acc(const void* dy, const void* x_raw, const float* gn_sums, const float* gamma,
                                   const float* beta, const float* scale_shift, int ss_ld, float* T_ws, void* dx_raw,
                                   float* dgamma, float* dbeta, float* dss, int dss_ld, float* dconv_bias, int B,
                                   int rows_per_sample, int C, int G, void* stream) {
  return gn_silu_bwd_impl(dy, x_raw, gn_sums, gamma, beta, scale_shift, ss_ld, T_ws, dx_raw, dgamma, dbeta, dss, dss_ld,
                          dconv_bias, B, rows_per_sample, C, G, stream, false);
}

static int gn_silu_bwd_impl(const void* dy, const void* x_raw, const float* gn_sums, const float* gamma, const float* beta,
                            const float* scale_shift, int ss_ld, float* T_ws, void* dx_raw, float* dgamma, float* dbeta,
                            float* dss, int dss_ld, float* dconv_bias, int B, int rows_per_sample, int C, int G, void* stream,
                            bool zero_ws) {
  int rc = check_gn("gn_silu_bwd", B, rows_per_sample, C, G);
  if (rc) return rc;
  VDN_REQUIRE(pow2(C / 8) && C / 8 <= kNormThreads, VDN_E_SHAPE, "gn_silu_bwd: C=%d must be 8 * power of two <= 2048", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GnArgs a{reinterpret_cast<const bf16*>(x_raw), gn_sums, gamma, beta, scale_shift, ss_ld, B, rows_per_sample, C, G};
  cudaError_t e = cudaSuccess;
  {
    // one cluster per (sample, channel slice) does both passes by itself (gn_bwd_group_kernel); T_ws receives the
    // per-channel sums with plain stores, so it needs no zeroing on this path. Cluster size 4 or 8 (row ranges): the
    // smallest that fits a CTA's slice into 10 vectors per thread and 200 KB of shared memory, as long as the launch
    // stays within one wave of one CTA per SM.
    const int cpg = C / G;
    const int cw = cpg < 8 ? 8 : cpg;
    const int slices = C / cw;
    int n_cl = 0, max_vec = 0;
    if ((cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) && !tune_on("VDN_GN_NO_GROUP")) {
      for (int cl = 4; cl <= 8; cl *= 2) {
        const long rows_cta = (rows_per_sample + cl - 1) / cl;
        const long nvec = rows_cta * (cw / 8);
        const long ctas = (long)slices * B * cl;
        if (nvec <= 10L * kGrpThreads && rows_cta * cw * 4 <= kGrpMaxSmem && ctas <= num_sms() && slices <= 65535) {
          n_cl = cl;
          max_vec = (int)((nvec + kGrpThreads - 1) / kGrpThreads);
          break;
        }
      }
    }
    // Measured on the training step (tools/ab_step.py): at the 8x8 / 16x16 levels (<= 2560 rows per sample) the single
    // launch is neutral to slightly faster (5.70 vs 5.72 ms) and saves 24 launches; at the 32x32 / 64x64 levels it is
    // SLOWER than the two wide kernels (5.74 / 5.89 ms): 128 CTAs of 512 threads do not stream a 10 MB tensor as fast as
    // 216 CTAs of 256, and the arithmetic of both passes sits in one kernel. Hence the row limit.
    if (rows_per_sample > tune_int("VDN_GN_GROUP_MAXROWS", 2560)) n_cl = 0;
    if (n_cl > 0) {
      const long rows_cta = (rows_per_sample + n_cl - 1) / n_cl;
      const size_t smem_g = (size_t)rows_cta * cw * sizeof(float);
      auto kern = max_vec <= 3 ? gn_bwd_group_kernel<3, true> : max_vec <= 6 ? gn_bwd_group_kernel<6, false>
                                                                             : gn_bwd_group_kernel<10, false>;
      static bool cfg_g = false;
      if (!cfg_g) {
        e = cudaFuncSetAttribute(gn_bwd_group_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGrpMaxSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_bwd_group_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGrpMaxSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_bwd_group_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGrpMaxSmem);
        VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "gn_bwd_group cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        cfg_g = true;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(slices, B, n_cl);
      cfg.blockDim = dim3(kGrpThreads);
      cfg.dynamicSmemBytes = smem_g;
      cfg.stream = st;
      cudaLaunchAttribute at[2];
      int na = 0;
      at[na].id = cudaLaunchAttributeClusterDimension;
      at[na].val.clusterDim.x = 1;
      at[na].val.clusterDim.y = 1;
      at[na].val.clusterDim.z = n_cl;
      ++na;
      if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
      cfg.attrs = at;
      cfg.numAttrs = na;
      cudaError_t lg = cudaLaunchKernelEx(&cfg, kern, a, reinterpret_cast<const bf16*>(dy), T_ws,
                                          reinterpret_cast<bf16*>(dx_raw), dgamma, dbeta, dss, dss_ld, dconv_bias);
      VDN_REQUIRE(lg == cudaSuccess, VDN_E_CUDA, "gn_bwd_group launch: %s", cudaGetErrorString(lg));
      return check_launch("gn_bwd_group");
    }
  }
  if (zero_ws) {
    e = cudaMemsetAsync(T_ws, 0, (size_t)B * C * 2 * sizeof(float), st);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "gn_silu_bwd memset: %s", cudaGetErrorString(e));
  }
  {
    // single-launch version when the (x, dy) slices of all samples fit in the SMs' shared memory. Opt-in
    // (VDN_GN_FUSED=1): measured 23.7 us against 32.9 us for the two kernels at the 64x64 level of config_v2_2 in
    // isolation, but no gain on the training step (6.87 vs 6.89 ms) - a grid that needs every SM cannot overlap with
    // the weight-gradient GEMMs of the side streams, which is where the two-kernel version hides its latency.
    const bool fe = tune_on("VDN_GN_FUSED");
    // VDN_GN_FUSED=1: every sample that fits; VDN_GN_FUSED_MAX=<elements per sample>: only samples up to that size
    const long fused_max = tune_int("VDN_GN_FUSED_MAX", 0);
    const bool fused_off = !fe && !((long)rows_per_sample * C <= fused_max);
    const int cps = B <= kFusedMaxB ? num_sms() / B : 0;  // CTAs per sample; cps * B <= number of SMs
    if (!fused_off && cps >= 1 && kFusedThreads % (C / 8) == 0) {
      const int rpc = (rows_per_sample + cps - 1) / cps;
      const size_t smem_f = (size_t)(9 * C + kFusedThreads * 16) * sizeof(float) + (size_t)rpc * C * 2 * 2;
      if (smem_f <= 220 * 1024) {
        static bool cfg = false;
        if (!cfg) {
          e = cudaFuncSetAttribute(gn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
          VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "gn_bwd_fused cudaFuncSetAttribute: %s", cudaGetErrorString(e));
          cfg = true;
        }
        static int next_slot = 0;
        const int slot = next_slot;
        next_slot = (next_slot + 1) % kFusedSlots;
        const int gx_f = (rows_per_sample + rpc - 1) / rpc;
        cudaError_t lf = launch_pdl(gn_bwd_fused_kernel, dim3(gx_f, B), dim3(kFusedThreads), smem_f, st, 1, a,
                                    reinterpret_cast<const bf16*>(dy), T_ws, reinterpret_cast<bf16*>(dx_raw), dgamma, dbeta, dss,
                                    dss_ld, dconv_bias, rpc, slot);
        VDN_REQUIRE(lf == cudaSuccess, VDN_E_CUDA, "gn_bwd_fused launch: %s", cudaGetErrorString(lf));
        return check_launch("gn_bwd_fused");
      }
    }
  }
  const int pl_n = kNormThreads / (C / 8);
  int gx, cl;
  // Chunks per block. These kernels hold two blocks per SM (register budget), i.e. 296 blocks in flight, and a block
  // is three dependent global round trips (statistics, data, atomics): more than ~1.4 waves of one-chunk blocks costs
  // more than folding the extra waves into the blocks (measured, tools/tune_gn.py, 64x64 / 32x32 levels of
  // config_v2_2: 640 blocks 32.8 us with 1 chunk, 26.6 us with 3; 320 blocks 24.6 -> 20.8 us; at <= 160 blocks extra
  // chunks only lengthen the one wave). Up to 8 for large samples.
  const long blocks1 = (long)grid_x_for(rows_per_sample, pl_n * kVecPerThread) * B;
  int iters = (int)std::max<long>(1, std::min<long>(8, (2 * blocks1 + 213) / (2 * 213)));
  if (tune_is_set("VDN_GN_ITERS")) iters = std::max(1, std::min(16, tune_int("VDN_GN_ITERS", iters)));  // experiments
  cl = cluster_for(grid_x_for(rows_per_sample, pl_n * kVecPerThread * iters), &gx);
  const size_t smem_r = (6 * C + kNormThreads * 16) * sizeof(float);
  cudaError_t le = iters > 1 ? launch_pdl(gn_bwd_reduce_kernel<true>, dim3(gx, B), dim3(kNormThreads), (size_t)(smem_r), st, cl, a, reinterpret_cast<const bf16*>(dy), T_ws, iters)
                             : launch_pdl(gn_bwd_reduce_kernel<false>, dim3(gx, B), dim3(kNormThreads), (size_t)(smem_r), st, cl, a, reinterpret_cast<const bf16*>(dy), T_ws, 1);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "gn_bwd_reduce launch: %s", cudaGetErrorString(le));
  rc = check_launch("gn_bwd_reduce");
  if (rc) return rc;
  const long nvec = (long)rows_per_sample * (C / 8);
  cl = cluster_for(grid_x_for(nvec, kNormThreads * kVecPerThread * iters), &gx);
  const size_t smem_a = (size_t)((8 * C + kNormThreads * 8) * sizeof(float));
  le = iters > 1 ? launch_pdl(gn_bwd_apply_kernel<true>, dim3(gx, B), dim3(kNormThreads), smem_a, st, cl, a,
                              reinterpret_cast<const bf16*>(dy), (const float*)T_ws, reinterpret_cast<bf16*>(dx_raw), dgamma,
                              dbeta, dss, dss_ld, dconv_bias, iters)
                 : launch_pdl(gn_bwd_apply_kernel<false>, dim3(gx, B), dim3(kNormThreads), smem_a, st, cl, a,
                              reinterpret_cast<const bf16*>(dy), (const float*)T_ws, reinterpret_cast<bf16*>(dx_raw), dgamma,
                              dbeta, dss, dss_ld, dconv_bias, 1);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "gn_bwd_apply launch: %s", cudaGetErrorString(le));
  return check_launch("gn_bwd_apply");
}

extern "C" int vdn_ln_bwd(const void* s, const void* dy, const float* ln_gamma, void* ds, float* dgamma, float* dbeta,
                          long P, int C, void* stream) {
  VDN_REQUIRE(P > 0 && C >= 8 && C % 8 == 0 && C <= 2048, VDN_E_SHAPE, "ln_bwd: bad shape P=%ld C=%d", P, C);
  const int vpl = vpl_for(C);
  VDN_REQUIRE(pow2(C / (8 * vpl)), VDN_E_SHAPE, "ln_bwd: C=%d must be 8 * power of two", C);
  const int lg = C / (8 * vpl);
  const int ppb = kNormThreads / lg;
  const size_t smem = 3 * C * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bf16* sp = reinterpret_cast<const bf16*>(s);
  const bf16* dp = reinterpret_cast<const bf16*>(dy);
  bf16* op = reinterpret_cast<bf16*>(ds);
  const int R = vpl == 1 ? 4 : vpl == 2 ? 2 : 1;
  int gx;
  const int iters = (int)std::max<long>(1, std::min<long>(8, (long)grid_x_for(P, ppb * R) / (2 * 8 * num_sms())));
  const int cl = cluster_for(grid_x_for(P, ppb * R * iters), &gx);
  cudaError_t le;
  switch (vpl) {
    case 1: le = launch_pdl(ln_bwd_kernel<1, 4>, dim3(gx), dim3(kNormThreads), smem, st, cl, sp, dp, ln_gamma, op, dgamma, dbeta, P, C, iters); break;
    case 2: le = launch_pdl(ln_bwd_kernel<2, 2>, dim3(gx), dim3(kNormThreads), smem, st, cl, sp, dp, ln_gamma, op, dgamma, dbeta, P, C, iters); break;
    case 4: le = launch_pdl(ln_bwd_kernel<4, 1>, dim3(gx), dim3(kNormThreads), smem, st, cl, sp, dp, ln_gamma, op, dgamma, dbeta, P, C, iters); break;
    default: le = launch_pdl(ln_bwd_kernel<8, 1>, dim3(gx), dim3(kNormThreads), smem, st, cl, sp, dp, ln_gamma, op, dgamma, dbeta, P, C, iters); break;
  }
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "ln_bwd launch: %s", cudaGetErrorString(le));
  return check_launch("ln_bwd");
}
