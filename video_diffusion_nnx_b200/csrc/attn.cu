// Attention cores of the Unet3D hot path, forward and backward.
//
//  * MultiheadAttention core (modules.py:285-324) for the temporal ('b f h w c -> b (h w) f c',
//    unet3d.py:86-96) and mid spatial ('b f (h w) c', unet3d.py:196-205) arrangements. The einops
//    transposes are never materialised: a sequence is addressed through strides into the fused
//    qkv projection [P][768] (q | k | v, head-major, 8 heads x 32).
//  * SpatialLinearAttention core (modules.py:105-123): q softmax over the 32 features (NOT scaled:
//    the scaled copy is dead code in the reference), k softmax over the N tokens, ctx = k~^T v,
//    out = q~ ctx.
//
// The projections around these cores run on the tensor cores (tapgemm.cu); the cores themselves
// are ~10 % of the FLOPs and run on CUDA cores with fp32 accumulation.
#include <algorithm>

#include <cstdlib>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kHeads = 8;
constexpr int kDh = 32;
constexpr int kHD = kHeads * kDh;   // 256
constexpr int kQKV = 3 * kHD;       // 768

struct SeqMap {
  long n_seq;
  int S;            // tokens per sequence
  long inner;       // row(seq,i) = (seq/inner)*outer_stride + (seq%inner)*inner_stride + i*tok_stride
  long outer_stride, inner_stride, tok_stride;
};
__device__ __forceinline__ long seq_row(const SeqMap& m, long seq, int i) {
  return (seq / m.inner) * m.outer_stride + (seq % m.inner) * m.inner_stride + (long)i * m.tok_stride;
}

__device__ __forceinline__ void load32(const bf16* p, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 u = reinterpret_cast<const uint4*>(p)[q];
    float2 f;
    f = unpack_bf16x2(u.x); v[8 * q + 0] = f.x; v[8 * q + 1] = f.y;
    f = unpack_bf16x2(u.y); v[8 * q + 2] = f.x; v[8 * q + 3] = f.y;
    f = unpack_bf16x2(u.z); v[8 * q + 4] = f.x; v[8 * q + 5] = f.y;
    f = unpack_bf16x2(u.w); v[8 * q + 6] = f.x; v[8 * q + 7] = f.y;
  }
}
__device__ __forceinline__ void store32(bf16* p, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
    u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
    u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
    reinterpret_cast<uint4*>(p)[q] = u;
  }
}
__device__ __forceinline__ float dot32(const float (&a)[32], const float (&b)[32]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // 4 independent chains (ILP)
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    s0 = fmaf(a[e], b[e], s0);
    s1 = fmaf(a[e + 1], b[e + 1], s1);
    s2 = fmaf(a[e + 2], b[e + 2], s2);
    s3 = fmaf(a[e + 3], b[e + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------------------
// MHA core. Two adjacent lanes share one (sequence, token, head): each owns 16 of the 32 head
// features (32 B loads, ~64 live registers), dot products are completed with one shuffle.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void load16(const bf16* p, float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint4 u = reinterpret_cast<const uint4*>(p)[q];
    float2 f;
    f = unpack_bf16x2(u.x); v[8 * q + 0] = f.x; v[8 * q + 1] = f.y;
    f = unpack_bf16x2(u.y); v[8 * q + 2] = f.x; v[8 * q + 3] = f.y;
    f = unpack_bf16x2(u.z); v[8 * q + 4] = f.x; v[8 * q + 5] = f.y;
    f = unpack_bf16x2(u.w); v[8 * q + 6] = f.x; v[8 * q + 7] = f.y;
  }
}
__device__ __forceinline__ void store16(bf16* p, const float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
    u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
    u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
    reinterpret_cast<uint4*>(p)[q] = u;
  }
}
// full 32-wide dot product of the lane pair
__device__ __forceinline__ float dot_pair(const float (&a)[16], const float (&b)[16]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int e = 0; e < 16; e += 4) {
    s0 = fmaf(a[e], b[e], s0);
    s1 = fmaf(a[e + 1], b[e + 1], s1);
    s2 = fmaf(a[e + 2], b[e + 2], s2);
    s3 = fmaf(a[e + 3], b[e + 3], s3);
  }
  const float s = (s0 + s1) + (s2 + s3);
  return s + __shfl_xor_sync(0xffffffffu, s, 1);
}

struct MhaIdx {
  bool valid;
  int h, tok, half;
  long seq;
};
__device__ __forceinline__ MhaIdx mha_index(const SeqMap& m) {
  const long total = m.n_seq * m.S * kHeads * 2;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  MhaIdx r;
  r.valid = idx < total;
  if (!r.valid) idx = total - 2 + (idx & 1);  // keep whole warps alive for the shuffles
  r.half = (int)(idx & 1);
  idx >>= 1;
  r.h = (int)(idx % kHeads);
  r.tok = (int)((idx / kHeads) % m.S);
  r.seq = idx / ((long)kHeads * m.S);
  return r;
}

__global__ void __launch_bounds__(256) mha_core_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o,
                                                           float* __restrict__ lse, const SeqMap m) {
  const MhaIdx t = mha_index(m);
  const int col = t.h * kDh + t.half * 16;
  const long row_i = seq_row(m, t.seq, t.tok);
  const float scale = rsqrtf((float)kDh);
  float q[16], acc[16];
  load16(qkv + row_i * kQKV + col, q);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    q[e] *= scale;
    acc[e] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
  for (int j = 0; j < m.S; ++j) {
    const long row_j = seq_row(m, t.seq, j);
    float kv[16];
    load16(qkv + row_j * kQKV + kHD + col, kv);
    const float sc = dot_pair(q, kv);
    const float mn = fmaxf(mx, sc);
    const float corr = __expf(mx - mn);
    const float p = __expf(sc - mn);
    load16(qkv + row_j * kQKV + 2 * kHD + col, kv);
    l = l * corr + p;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = fmaf(acc[e], corr, p * kv[e]);
    mx = mn;
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] *= inv;
  if (t.valid) {
    store16(o + row_i * kHD + col, acc);
    if (t.half == 0) lse[row_i * kHeads + t.h] = mx + __logf(l);
  }
}

// Backward, phase 1: dq (and D_i = do_i . o_i for phase 2).
__global__ void __launch_bounds__(256) mha_core_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                              const bf16* __restrict__ d_o,
                                                              const float* __restrict__ lse, float* __restrict__ Dws,
                                                              bf16* __restrict__ dqkv, const SeqMap m) {
  const MhaIdx t = mha_index(m);
  const int col = t.h * kDh + t.half * 16;
  const long row_i = seq_row(m, t.seq, t.tok);
  const float scale = rsqrtf((float)kDh);
  float q[16], dov[16], tmp[16], dq[16];
  load16(qkv + row_i * kQKV + col, q);
  load16(d_o + row_i * kHD + col, dov);
  load16(o + row_i * kHD + col, tmp);
  const float D = dot_pair(dov, tmp);
  if (t.valid && t.half == 0) Dws[row_i * kHeads + t.h] = D;
  const float L = lse[row_i * kHeads + t.h];
#pragma unroll
  for (int e = 0; e < 16; ++e) dq[e] = 0.f;
  for (int j = 0; j < m.S; ++j) {
    const long row_j = seq_row(m, t.seq, j);
    float kv[16];
    load16(qkv + row_j * kQKV + kHD + col, kv);
    const float p = __expf(dot_pair(q, kv) * scale - L);
    load16(qkv + row_j * kQKV + 2 * kHD + col, tmp);
    const float ds = p * (dot_pair(dov, tmp) - D) * scale;
#pragma unroll
    for (int e = 0; e < 16; ++e) dq[e] = fmaf(ds, kv[e], dq[e]);
  }
  if (t.valid) store16(dqkv + row_i * kQKV + col, dq);
}

// Backward, phase 2: dk_j, dv_j; lane pair per (sequence, key token, head).
__global__ void __launch_bounds__(256) mha_core_bwd_dkv_kernel(const bf16* __restrict__ qkv,
                                                               const bf16* __restrict__ d_o,
                                                               const float* __restrict__ lse,
                                                               const float* __restrict__ Dws,
                                                               bf16* __restrict__ dqkv, const SeqMap m) {
  const MhaIdx t = mha_index(m);
  const int col = t.h * kDh + t.half * 16;
  const long row_j = seq_row(m, t.seq, t.tok);
  const float scale = rsqrtf((float)kDh);
  float kj[16], vj[16], dk[16], dv[16];
  load16(qkv + row_j * kQKV + kHD + col, kj);
  load16(qkv + row_j * kQKV + 2 * kHD + col, vj);
#pragma unroll
  for (int e = 0; e < 16; ++e) dk[e] = dv[e] = 0.f;
  for (int i = 0; i < m.S; ++i) {
    const long row_i = seq_row(m, t.seq, i);
    float qi[16], doi[16];
    load16(qkv + row_i * kQKV + col, qi);
    load16(d_o + row_i * kHD + col, doi);
    const float p = __expf(dot_pair(qi, kj) * scale - lse[row_i * kHeads + t.h]);
    const float ds = p * (dot_pair(doi, vj) - Dws[row_i * kHeads + t.h]) * scale;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      dv[e] = fmaf(p, doi[e], dv[e]);
      dk[e] = fmaf(ds, qi[e], dk[e]);
    }
  }
  if (t.valid) {
    store16(dqkv + row_j * kQKV + kHD + col, dk);
    store16(dqkv + row_j * kQKV + 2 * kHD + col, dv);
  }
}

// ---------------------------------------------------------------------------------------
// SpatialLinearAttention
// ---------------------------------------------------------------------------------------
constexpr int kSlaTile = 64;  // tokens per smem tile

// Partial context over a token range with online softmax over tokens (per feature column d).
// grid (n_split, heads, n_img), 256 threads = 4 token sub-groups x 64 threads; a thread owns the 4x4
// register tile ctx[d4..d4+3][e4..e4+3] for the tokens n = sub (mod 4) of each 64-token smem tile:
// 2 LDS.128 feed 16 FMAs. The 4 sub-group partials are summed through smem at the end.
__global__ void __launch_bounds__(256) sla_ctx_partial_kernel(const bf16* __restrict__ qkv, int N, int tokens_per_split,
                                                              float* __restrict__ ctx_part /*[img][h][split][32][32]*/,
                                                              float* __restrict__ ms_part /*[img][h][split][2][32]*/) {
  __shared__ __align__(16) float kt[kSlaTile][36];   // k, then p = exp(k - m); pitch 36 keeps LDS.128 aligned
  __shared__ __align__(16) float vt[kSlaTile][32];
  __shared__ float pm[8][32];
  __shared__ float sm_m[32], sm_corr[32];
  __shared__ __align__(16) float red[4][1024 + 32];
  const int split = blockIdx.x, h = blockIdx.y, img = blockIdx.z;
  const int n_split = gridDim.x;
  const int tid = threadIdx.x;
  const int sub = tid >> 6, t64 = tid & 63;
  const int d4 = (t64 >> 3) * 4, e4 = (t64 & 7) * 4;
  const int n_begin = split * tokens_per_split;
  const int n_end = min(N, n_begin + tokens_per_split);
  float c[4][4];
  float ss[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ss[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  }
  if (tid < 32) sm_m[tid] = -INFINITY;
  const bf16* base = qkv + (long)img * N * kQKV;
  for (int n0 = n_begin; n0 < n_end; n0 += kSlaTile) {
    const int nt = min(kSlaTile, n_end - n0);
    {  // stage k, v tile as fp32; thread -> (token tid/4, 8 features (tid%4)*8)
      const int t = tid >> 2, part = (tid & 3) * 8;
      float kv8[8], vv8[8];
      if (t < nt) {
        const bf16* rp = base + (long)(n0 + t) * kQKV + h * kDh + part;
        const uint4 uk = *reinterpret_cast<const uint4*>(rp + kHD);
        const uint4 uv = *reinterpret_cast<const uint4*>(rp + 2 * kHD);
        float2 f;
        f = unpack_bf16x2(uk.x); kv8[0] = f.x; kv8[1] = f.y;
        f = unpack_bf16x2(uk.y); kv8[2] = f.x; kv8[3] = f.y;
        f = unpack_bf16x2(uk.z); kv8[4] = f.x; kv8[5] = f.y;
        f = unpack_bf16x2(uk.w); kv8[6] = f.x; kv8[7] = f.y;
        f = unpack_bf16x2(uv.x); vv8[0] = f.x; vv8[1] = f.y;
        f = unpack_bf16x2(uv.y); vv8[2] = f.x; vv8[3] = f.y;
        f = unpack_bf16x2(uv.z); vv8[4] = f.x; vv8[5] = f.y;
        f = unpack_bf16x2(uv.w); vv8[6] = f.x; vv8[7] = f.y;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          kv8[j] = -INFINITY;  // exp() -> 0: padded tokens contribute nothing
          vv8[j] = 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        kt[t][part + j] = kv8[j];
        vt[t][part + j] = vv8[j];
      }
    }
    __syncthreads();
    {  // tile column max: thread -> (column tid%32, 8-token segment tid/32)
      const int col = tid & 31, seg = tid >> 5;
      float mxv = -INFINITY;
#pragma unroll
      for (int r = 0; r < 8; ++r) mxv = fmaxf(mxv, kt[seg * 8 + r][col]);
      pm[seg][col] = mxv;
    }
    __syncthreads();
    if (tid < 32) {
      float tmax = pm[0][tid];
#pragma unroll
      for (int sgm = 1; sgm < 8; ++sgm) tmax = fmaxf(tmax, pm[sgm][tid]);
      const float m_old = sm_m[tid];
      const float m_new = fmaxf(m_old, tmax);
      sm_corr[tid] = __expf(m_old - m_new);  // m_old = -inf on the first tile -> 0
      sm_m[tid] = m_new;
    }
    __syncthreads();
    {  // p = exp(k - m) in place; thread -> (column tid%32, tokens seg*8..)
      const int col = tid & 31, seg = tid >> 5;
      const float mc = sm_m[col];
#pragma unroll
      for (int r = 0; r < 8; ++r) kt[seg * 8 + r][col] = __expf(kt[seg * 8 + r][col] - mc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float cr = sm_corr[d4 + i];
      ss[i] *= cr;
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] *= cr;
    }
    __syncthreads();
#pragma unroll 4
    for (int n = sub; n < kSlaTile; n += 4) {
      const float4 p4 = *reinterpret_cast<const float4*>(&kt[n][d4]);
      const float4 v4 = *reinterpret_cast<const float4*>(&vt[n][e4]);
      const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        c[i][0] = fmaf(pp[i], v4.x, c[i][0]); c[i][1] = fmaf(pp[i], v4.y, c[i][1]);
        c[i][2] = fmaf(pp[i], v4.z, c[i][2]); c[i][3] = fmaf(pp[i], v4.w, c[i][3]);
        ss[i] += pp[i];
      }
    }
    __syncthreads();
  }
  // reduce the 4 token sub-groups
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    *reinterpret_cast<float4*>(&red[sub][(d4 + i) * 32 + e4]) = make_float4(c[i][0], c[i][1], c[i][2], c[i][3]);
    if (e4 == 0) red[sub][1024 + d4 + i] = ss[i];
  }
  __syncthreads();
  const long blk = ((long)img * kHeads + h) * n_split + split;
  for (int i = tid; i < 1024; i += 256)
    ctx_part[blk * 1024 + i] = red[0][i] + red[1][i] + red[2][i] + red[3][i];
  if (tid < 32) {
    ms_part[blk * 64 + tid] = sm_m[tid];
    ms_part[blk * 64 + 32 + tid] = red[0][1024 + tid] + red[1][1024 + tid] + red[2][1024 + tid] + red[3][1024 + tid];
  }
}

// Merge split partials: ctx = sum_s w_s ctx_s / sum_s w_s s_s, w_s = exp(m_s - m). grid (img*heads), 256 thr.
__global__ void __launch_bounds__(256) sla_ctx_merge_kernel(const float* __restrict__ ctx_part,
                                                            const float* __restrict__ ms_part, int n_split,
                                                            float* __restrict__ ctx /*[img][h][32][32]*/,
                                                            float* __restrict__ kstat /*[img][h][2][32] (m, S)*/) {
  const long ih = blockIdx.x;
  const int tid = threadIdx.x;
  const int d = tid >> 3, e4 = (tid & 7) * 4;
  float m = -INFINITY;
  for (int s = 0; s < n_split; ++s) m = fmaxf(m, ms_part[(ih * n_split + s) * 64 + d]);
  float S = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const long blk = ih * n_split + s;
    const float w = __expf(ms_part[blk * 64 + d] - m);
    S += w * ms_part[blk * 64 + 32 + d];
    const float* cp = ctx_part + blk * 1024 + d * 32 + e4;
    c0 += w * cp[0]; c1 += w * cp[1]; c2 += w * cp[2]; c3 += w * cp[3];
  }
  const float inv = 1.f / S;
  float* op = ctx + ih * 1024 + d * 32 + e4;
  op[0] = c0 * inv; op[1] = c1 * inv; op[2] = c2 * inv; op[3] = c3 * inv;
  if ((tid & 7) == 0) {
    kstat[ih * 64 + d] = m;
    kstat[ih * 64 + 32 + d] = S;
  }
}

__device__ __forceinline__ void softmax32(float (&q)[32]) {
  float mx = q[0];
#pragma unroll
  for (int e = 1; e < 32; ++e) mx = fmaxf(mx, q[e]);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    q[e] = __expf(q[e] - mx);
    s += q[e];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int e = 0; e < 32; ++e) q[e] *= inv;
}

// out[n, h*32+e] = sum_d softmax_D(q[n,h,:])[d] * ctx[h][d][e].
// grid (chunks, n_img), 256 threads = 8 warps = 8 heads; a warp's lanes are 32 consecutive tokens,
// so ctx reads are smem broadcasts.
__global__ void __launch_bounds__(256) sla_apply_kernel(const bf16* __restrict__ qkv, const float* __restrict__ ctx,
                                                        bf16* __restrict__ out, int N) {
  extern __shared__ float sctx[];  // [8][32][32]
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) sctx[i] = ctx[(long)img * 8 * 1024 + i];
  __syncthreads();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* ch = sctx + h * 1024;
  for (int n0 = blockIdx.x * 32; n0 < N; n0 += gridDim.x * 32) {
    const int n = n0 + lane;
    if (n >= N) continue;
    const long row = (long)img * N + n;
    float q[32], o[32];
    load32(qkv + row * kQKV + h * kDh, q);
    softmax32(q);
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e] = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float qd = q[d];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(&ch[d * 32 + e]);
        o[e] = fmaf(qd, c4.x, o[e]); o[e + 1] = fmaf(qd, c4.y, o[e + 1]);
        o[e + 2] = fmaf(qd, c4.z, o[e + 2]); o[e + 3] = fmaf(qd, c4.w, o[e + 3]);
      }
    }
    store32(out + row * kHD + h * kDh, o);
  }
}

// dctx[h][d][e] += sum_n q~[n,d] * dtok[n,e] over a token range. grid (n_split, heads, n_img); same
// 4 sub-groups x (4x4 register tile) scheme as sla_ctx_partial_kernel.
__global__ void __launch_bounds__(256) sla_dctx_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dtok,
                                                       int N, int tokens_per_split, float* __restrict__ dctx) {
  __shared__ __align__(16) float qt[kSlaTile][36];
  __shared__ __align__(16) float gt[kSlaTile][32];
  __shared__ __align__(16) float red[4][1024];
  const int split = blockIdx.x, h = blockIdx.y, img = blockIdx.z;
  const int tid = threadIdx.x;
  const int sub = tid >> 6, t64 = tid & 63;
  const int d4 = (t64 >> 3) * 4, e4 = (t64 & 7) * 4;
  const int n_begin = split * tokens_per_split;
  const int n_end = min(N, n_begin + tokens_per_split);
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int n0 = n_begin; n0 < n_end; n0 += kSlaTile) {
    const int nt = min(kSlaTile, n_end - n0);
    // q~ rows: one thread per token computes the 32-wide softmax (tokens 0..63 -> threads 0..63)
    if (tid < kSlaTile) {
      float q[32];
      if (tid < nt) {
        load32(qkv + ((long)img * N + n0 + tid) * kQKV + h * kDh, q);
        softmax32(q);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) q[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(&qt[tid][e]) = make_float4(q[e], q[e + 1], q[e + 2], q[e + 3]);
    } else if (tid < 2 * kSlaTile) {
      const int t = tid - kSlaTile;
      float g[32];
      if (t < nt) {
        load32(dtok + ((long)img * N + n0 + t) * kHD + h * kDh, g);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) g[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(&gt[t][e]) = make_float4(g[e], g[e + 1], g[e + 2], g[e + 3]);
    }
    __syncthreads();
#pragma unroll 4
    for (int n = sub; n < kSlaTile; n += 4) {
      const float4 p4 = *reinterpret_cast<const float4*>(&qt[n][d4]);
      const float4 v4 = *reinterpret_cast<const float4*>(&gt[n][e4]);
      const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        c[i][0] = fmaf(pp[i], v4.x, c[i][0]); c[i][1] = fmaf(pp[i], v4.y, c[i][1]);
        c[i][2] = fmaf(pp[i], v4.z, c[i][2]); c[i][3] = fmaf(pp[i], v4.w, c[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(&red[sub][(d4 + i) * 32 + e4]) = make_float4(c[i][0], c[i][1], c[i][2], c[i][3]);
  __syncthreads();
  float* op = dctx + ((long)img * kHeads + h) * 1024;
  for (int i = tid; i < 1024; i += 256) atomicAdd(op + i, red[0][i] + red[1][i] + red[2][i] + red[3][i]);
}

// Per-token backward: dq, dk, dv from ctx, dctx, k statistics. Same thread mapping as sla_apply.
__global__ void __launch_bounds__(256) sla_bwd_tokens_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dtok,
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
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) {
    sctx[i] = ctx[(long)img * 8 * 1024 + i];
    sdctx[i] = dctx[(long)img * 8 * 1024 + i];
  }
  __syncthreads();
  {
    const int hh = threadIdx.x >> 5, dd = threadIdx.x & 31;
    sm_m[threadIdx.x] = kstat[((long)img * kHeads + hh) * 64 + dd];
    sm_is[threadIdx.x] = 1.f / kstat[((long)img * kHeads + hh) * 64 + 32 + dd];
    float r = 0.f;
    for (int e = 0; e < 32; ++e) r += sdctx[hh * 1024 + dd * 32 + e] * sctx[hh * 1024 + dd * 32 + e];
    sm_r[threadIdx.x] = r;
  }
  __syncthreads();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* ch = sctx + h * 1024;
  const float* dch = sdctx + h * 1024;
  for (int n0 = blockIdx.x * 32; n0 < N; n0 += gridDim.x * 32) {
    const int n = n0 + lane;
    if (n >= N) continue;
    const long row = (long)img * N + n;
    float a[32], g[32], r[32];
    // ---- dq ----
    load32(qkv + row * kQKV + h * kDh, a);   // q
    softmax32(a);                            // q~
    load32(dtok + row * kHD + h * kDh, g);   // d_tok
    float dotq = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(&ch[d * 32 + e]);
        t0 = fmaf(c4.x, g[e], t0); t1 = fmaf(c4.y, g[e + 1], t1);
        t2 = fmaf(c4.z, g[e + 2], t2); t3 = fmaf(c4.w, g[e + 3], t3);
      }
      const float t = (t0 + t1) + (t2 + t3);
      r[d] = t;  // dq~[d]
      dotq = fmaf(a[d], t, dotq);
    }
#pragma unroll
    for (int d = 0; d < 32; ++d) r[d] = a[d] * (r[d] - dotq);
    store32(dqkv + row * kQKV + h * kDh, r);
    // ---- dk ----
    load32(qkv + row * kQKV + kHD + h * kDh, a);      // k
    load32(qkv + row * kQKV + 2 * kHD + h * kDh, g);  // v
#pragma unroll
    for (int d = 0; d < 32; ++d) a[d] = __expf(a[d] - sm_m[h * 32 + d]) * sm_is[h * 32 + d];  // k~
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(&dch[d * 32 + e]);
        t0 = fmaf(c4.x, g[e], t0); t1 = fmaf(c4.y, g[e + 1], t1);
        t2 = fmaf(c4.z, g[e + 2], t2); t3 = fmaf(c4.w, g[e + 3], t3);
      }
      r[d] = a[d] * (((t0 + t1) + (t2 + t3)) - sm_r[h * 32 + d]);
    }
    store32(dqkv + row * kQKV + kHD + h * kDh, r);
    // ---- dv[e] = sum_d k~[d] dctx[d][e] ----
#pragma unroll
    for (int e = 0; e < 32; ++e) r[e] = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float kd = a[d];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(&dch[d * 32 + e]);
        r[e] = fmaf(kd, c4.x, r[e]); r[e + 1] = fmaf(kd, c4.y, r[e + 1]);
        r[e + 2] = fmaf(kd, c4.z, r[e + 2]); r[e + 3] = fmaf(kd, c4.w, r[e + 3]);
      }
    }
    store32(dqkv + row * kQKV + 2 * kHD + h * kDh, r);
  }
}

// Token splits per frame for the context / dcontext passes: at most 32, and no more than needed to give every SM
// its resident blocks (n_img frames x splits): every split writes a 32 KB partial per frame that the merge pass re-reads,
// which at large batch (n_img = 160) was more traffic than the input itself.
static int sla_splits(int N, int n_img) {
  int per = std::max(kSlaTile, ((N + 31) / 32 + kSlaTile - 1) / kSlaTile * kSlaTile);  // <= 32 splits
  const int max_splits = std::max(1, (N + per - 1) / per);
  // The context kernels keep two blocks of 256 threads resident per SM (92 - 124 registers): take the smallest split
  // count whose n_img x splits blocks fill whole waves of those 296 slots to >= 85 % (40 frames: 7 splits = 280 blocks;
  // 160 frames: 5 splits = 800 blocks = 2.7 waves of 3). Four blocks per SM with a ceiling gave 600 blocks at 40 frames -
  // a wave and a half; 5.84 vs 6.04 ms per training step.
  const int slots = 148 * tune_int("VDN_SLA_BLOCKS_PER_SM", 2);
  int want = 1;
  double best = 0.0;
  for (int sp = 1; sp <= max_splits; ++sp) {
    const long blocks = (long)n_img * sp;
    const long waves = (blocks + slots - 1) / slots;
    const double eff = (double)blocks / (double)(waves * slots);
    if (eff > best + 1e-9) {
      best = eff;
      want = sp;
    }
    if (eff >= 0.85) break;
  }
  return std::min(max_splits, want);
}

}  // namespace vdn

using namespace vdn;

static SeqMap make_seqmap(int mode, int B, int F, int HW) {
  SeqMap m;
  if (mode == 0) {  // temporal: sequences (b, pixel), tokens = frames
    m.n_seq = (long)B * HW; m.S = F; m.inner = HW; m.outer_stride = (long)F * HW; m.inner_stride = 1; m.tok_stride = HW;
  } else {          // spatial: sequences (b, f), tokens = pixels
    m.n_seq = (long)B * F; m.S = HW; m.inner = 1; m.outer_stride = HW; m.inner_stride = 0; m.tok_stride = 1;
  }
  return m;
}

extern "C" int vdn_mha_core_fwd(const void* qkv, void* o, float* lse, int mode, int B, int F, int HW, void* stream) {
  VDN_REQUIRE(qkv && o && lse && B > 0 && F > 0 && HW > 0 && (mode == 0 || mode == 1), VDN_E_SHAPE, "mha_core_fwd: bad args");
  if (mode == 1 && mha_spatial_mma_applicable(HW) && !tune_on("VDN_MHA_SPATIAL_SCALAR"))
    return mha_spatial_mma_fwd_launch(qkv, o, lse, B * F, HW, reinterpret_cast<cudaStream_t>(stream));
  const SeqMap m = make_seqmap(mode, B, F, HW);
  const long total = m.n_seq * m.S * kHeads * 2;
  mha_core_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(o), lse, m);
  return check_launch("mha_core_fwd");
}

extern "C" int vdn_mha_core_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* D_ws,
                                void* dqkv, int mode, int B, int F, int HW, void* stream) {
  VDN_REQUIRE(qkv && o && d_o && lse && D_ws && dqkv && (mode == 0 || mode == 1), VDN_E_SHAPE, "mha_core_bwd: bad args");
  if (mode == 1 && mha_spatial_mma_applicable(HW) && !tune_on("VDN_MHA_SPATIAL_SCALAR"))
    return mha_spatial_mma_bwd_launch(qkv, o, d_o, lse, dqkv, B * F, HW, reinterpret_cast<cudaStream_t>(stream));
  const SeqMap m = make_seqmap(mode, B, F, HW);
  const long total = m.n_seq * m.S * kHeads * 2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((total + 255) / 256);
  mha_core_bwd_dq_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(o),
                                               reinterpret_cast<const bf16*>(d_o), lse, D_ws,
                                               reinterpret_cast<bf16*>(dqkv), m);
  int rc = check_launch("mha_core_bwd_dq");
  if (rc) return rc;
  mha_core_bwd_dkv_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(d_o),
                                                lse, D_ws, reinterpret_cast<bf16*>(dqkv), m);
  return check_launch("mha_core_bwd_dkv");
}

extern "C" size_t vdn_sla_workspace_floats(int n_img, int N) {
  const int ns = sla_splits(N, n_img);
  return (size_t)n_img * kHeads * ns * (1024 + 64);
}

extern "C" int vdn_sla_core_fwd(const void* qkv, void* tok_out, float* ctx, float* kstat, float* ws, int n_img, int N,
                                void* stream) {
  VDN_REQUIRE(qkv && tok_out && ctx && kstat && ws && n_img > 0 && N > 0, VDN_E_SHAPE, "sla_core_fwd: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ns = sla_splits(N, n_img);
  const int per = (N + ns - 1) / ns;
  const int per_al = (per + kSlaTile - 1) / kSlaTile * kSlaTile;
  float* ctx_part = ws;
  float* ms_part = ws + (size_t)n_img * kHeads * ns * 1024;
  const bool scalar = tune_on("VDN_SLA_SCALAR");  // CUDA-core kernels (A/B comparison only)
  int rc;
  if (scalar) {
    sla_ctx_partial_kernel<<<dim3(ns, kHeads, n_img), 256, 0, st>>>(reinterpret_cast<const bf16*>(qkv), N, per_al,
                                                                    ctx_part, ms_part);
    rc = check_launch("sla_ctx_partial");
  } else {
    rc = sla_ctx_partial_mma_launch(qkv, N, per_al, ns, ctx_part, ms_part, n_img, st);
  }
  if (rc) return rc;
  sla_ctx_merge_kernel<<<n_img * kHeads, 256, 0, st>>>(ctx_part, ms_part, ns, ctx, kstat);
  rc = check_launch("sla_ctx_merge");
  if (rc) return rc;
  if (!scalar) return sla_apply_mma_launch(qkv, ctx, tok_out, n_img, N, st);
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(sla_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 4);
    cfg = true;
  }
  const int gx = std::max(1, std::min((N + 31) / 32, std::max(1, 148 * 8 / n_img)));
  sla_apply_kernel<<<dim3(gx, n_img), 256, 8 * 1024 * sizeof(float), st>>>(reinterpret_cast<const bf16*>(qkv), ctx,
                                                                           reinterpret_cast<bf16*>(tok_out), N);
  return check_launch("sla_apply");
}

// Fused SpatialLinearAttention forward, C = 32: out = x + to_out(SLA(to_qkv(x))) without materialising q/k/v/tok.
// w_qkv: packed bf16 [768][32] (rows q | k | v, head-major), w_out: packed bf16 [32][256]; ctx / kstat are outputs
// like vdn_sla_core_fwd, ws the same scratch.
extern "C" int vdn_sla_fused_fwd(const void* x, const void* w_qkv, const void* w_out, void* out, float* ctx,
                                 float* kstat, float* ws, int n_img, int N, int C, void* stream) {
  VDN_REQUIRE(x && w_qkv && w_out && out && ctx && kstat && ws && n_img > 0 && N > 0, VDN_E_SHAPE, "sla_fused_fwd: bad args");
  VDN_REQUIRE(C == 32, VDN_E_SHAPE, "sla_fused_fwd: C=%d (only the 32-channel level is fused)", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ns = sla_splits(N, n_img);
  const int per = (N + ns - 1) / ns;
  const int per_al = (per + kSlaTile - 1) / kSlaTile * kSlaTile;
  float* ctx_part = ws;
  float* ms_part = ws + (size_t)n_img * kHeads * ns * 1024;
  int rc = sla_ctx_fused_launch(x, w_qkv, N, per_al, ns, ctx_part, ms_part, n_img, st);
  if (rc) return rc;
  sla_ctx_merge_kernel<<<n_img * kHeads, 256, 0, st>>>(ctx_part, ms_part, ns, ctx, kstat);
  rc = check_launch("sla_ctx_merge");
  if (rc) return rc;
  // apply pass as two tcgen05 GEMMs around a thread-local feature softmax (sla_apply_tc.cu); the split partials in ws
  // are dead after the merge, so the folded per-image matrices (16 KB each) take their place.
  // VDN_SLA_APPLY_MMA=1 keeps the all-mma.sync kernel.
  if (sla_apply_tc_applicable(N) && !tune_on("VDN_SLA_APPLY_MMA"))
    return sla_apply_tc_launch(x, w_qkv, w_out, ctx, ws, out, n_img, N, st);
  return sla_apply_fused_launch(x, w_qkv, w_out, ctx, out, n_img, N, st);
}

static int sla_core_bwd_impl(const void* qkv, const void* d_tok, const float* ctx, const float* kstat, float* dctx,
                             void* dqkv, int n_img, int N, void* stream, bool zero_ws);
extern "C" int vdn_sla_core_bwd(const void* qkv, const void* d_tok, const float* ctx, const float* kstat, float* dctx,
                                void* dqkv, int n_img, int N, void* stream) {
  return sla_core_bwd_impl(qkv, d_tok, ctx, kstat, dctx, dqkv, n_img, N, stream, true);
}
// Same with dctx zeroed by the CALLER (one memset for all the blocks of a step instead of a memset node per call).
extern "C" int vdn_sla_core_bwd_acc(const void* qkv, const void* d_tok, const float* ctx, const float* kstat, float* dctx,
                                    void* dqkv, int n_img, int N, void* stream) {
  return sla_core_bwd_impl(qkv, d_tok, ctx, kstat, dctx, dqkv, n_img, N, stream, false);
}
static int sla_core_bwd_impl(const void* qkv, const void* d_tok, const float* ctx, const float* kstat, float* dctx,
                             void* dqkv, int n_img, int N, void* stream, bool zero_ws) {
  VDN_REQUIRE(qkv && d_tok && ctx && kstat && dctx && dqkv && n_img > 0 && N > 0, VDN_E_SHAPE, "sla_core_bwd: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (zero_ws) {
    cudaError_t e = cudaMemsetAsync(dctx, 0, (size_t)n_img * kHeads * 1024 * sizeof(float), st);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "sla_core_bwd memset: %s", cudaGetErrorString(e));
  }
  const int ns = sla_splits(N, n_img);
  const int per = (N + ns - 1) / ns;
  const int per_al = (per + kSlaTile - 1) / kSlaTile * kSlaTile;
  const bool scalar = tune_on("VDN_SLA_SCALAR");
  int rc;
  if (scalar) {
    sla_dctx_kernel<<<dim3(ns, kHeads, n_img), 256, 0, st>>>(reinterpret_cast<const bf16*>(qkv),
                                                             reinterpret_cast<const bf16*>(d_tok), N, per_al, dctx);
    rc = check_launch("sla_dctx");
  } else {
    rc = sla_dctx_mma_launch(qkv, d_tok, N, per_al, ns, dctx, n_img, st);
  }
  if (rc) return rc;
  if (!scalar) return sla_bwd_tokens_mma_launch(qkv, d_tok, ctx, dctx, kstat, dqkv, n_img, N, st);
  const size_t smem = (16 * 1024 + 3 * 256) * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(sla_bwd_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int gx = std::max(1, std::min((N + 31) / 32, std::max(1, 148 * 6 / n_img)));
  sla_bwd_tokens_kernel<<<dim3(gx, n_img), 256, smem, st>>>(reinterpret_cast<const bf16*>(qkv),
                                                            reinterpret_cast<const bf16*>(d_tok), ctx, dctx, kstat,
                                                            reinterpret_cast<bf16*>(dqkv), N);
  return check_launch("sla_bwd_tokens");
}
