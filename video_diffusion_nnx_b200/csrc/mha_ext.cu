// MultiheadAttention core with the reference's OPTIONAL inputs, and RelativePositionBias on the device.
//
//   modules.py:285-326  MultiheadAttention.__call__(x, focus_present_mask=None, pos_bias=None)
//   modules.py:330-390  RelativePositionBias (bucket ids :351-378, __call__ :380-390)
//
// Inside Unet3D neither input ever reaches an attention (PreNorm drops kwargs, modules.py:146-148), so the hot
// path's fused kernels (mha_mma.cu, mha_tc.cu, mha_fused.cu) do not carry them. This file serves the module when
// it is called directly (test_modules.py:242-271) and reproduces the reference literally:
//   * attn = softmax_j(q k^T / sqrt(dim))                                  (:294-304)
//   * focus_present_mask (per batch element): entries off the diagonal are REPLACED, after the softmax, by
//     finfo(float32).min (:307-316)
//   * pos_bias is ADDED after the softmax (:320-321)
//   * all batch elements focusing on the present: the module returns out(v) (:291-292) - copy_v below.
// Templated on the q|k|v element type: bf16 (tensor-core projection output) and fp32 (the fp32-grade path, which
// also uses this kernel - without mask / bias - as its temporal and spatial attention core).
#include <cfloat>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kExtThreads = 128;

template <typename T>
__device__ __forceinline__ float ext_ld(const T* p);
template <>
__device__ __forceinline__ float ext_ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ext_ld<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void ext_st(T* p, float v);
template <>
__device__ __forceinline__ void ext_st<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void ext_st<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

struct ExtArgs {
  int n_seq, S;        // sequences, tokens per sequence
  int inner;           // token t of sequence s lives at row (s / inner) * S * inner + (s % inner) + t * inner
  int seqs_per_batch;  // mask_b[s / seqs_per_batch]
  int heads;           // q|k|v row = [3][heads][D], o row = [heads][D]
  float scale;         // 1 / sqrt(D), modules.py:294
  const unsigned char* mask_b;  // [n_batch] or NULL; non-zero = this batch element attends to itself only
  const float* pos_bias;        // [heads][S][S] or NULL
  int copy_v;                   // 1: o = v (the all-focus early return)
};

// One block per (sequence, head). K and V of the head are staged in shared memory as fp32; a thread owns a query.
template <typename T, int kExtDim>
__global__ void __launch_bounds__(kExtThreads) mha_core_ext_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ o,
                                                                        const ExtArgs a) {
  extern __shared__ float ext_smem[];
  float* ks = ext_smem;                    // [S][32]
  float* vs = ext_smem + (size_t)a.S * kExtDim;
  const int s = blockIdx.x, h = blockIdx.y;
  const long row0 = (long)(s / a.inner) * a.S * a.inner + (s % a.inner);
  const int hd = a.heads * kExtDim;
  const int ld_qkv = 3 * hd, ld_o = hd;
  for (int idx = threadIdx.x; idx < a.S * kExtDim; idx += kExtThreads) {
    const int j = idx / kExtDim, d = idx % kExtDim;
    const T* rowp = qkv + (row0 + (long)j * a.inner) * ld_qkv + h * kExtDim + d;
    ks[idx] = ext_ld(rowp + hd);
    vs[idx] = ext_ld(rowp + 2 * hd);
  }
  __syncthreads();
  const bool focus = a.mask_b != nullptr && a.mask_b[s / a.seqs_per_batch] != 0;
  const float scale = a.scale;
  for (int i = threadIdx.x; i < a.S; i += kExtThreads) {
    T* op = o + (row0 + (long)i * a.inner) * ld_o + h * kExtDim;
    if (a.copy_v) {
#pragma unroll
      for (int d = 0; d < kExtDim; ++d) ext_st(op + d, vs[i * kExtDim + d]);
      continue;
    }
    float q[kExtDim];
    const T* qp = qkv + (row0 + (long)i * a.inner) * ld_qkv + h * kExtDim;
#pragma unroll
    for (int d = 0; d < kExtDim; ++d) q[d] = ext_ld(qp + d) * scale;
    // pass 1: row maximum; pass 2: normaliser (jax.nn.softmax: exp(x - max) / sum)
    float mx = -FLT_MAX;
    for (int j = 0; j < a.S; ++j) {
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < kExtDim; ++d) sc = fmaf(q[d], ks[j * kExtDim + d], sc);
      mx = fmaxf(mx, sc);
    }
    float den = 0.f;
    for (int j = 0; j < a.S; ++j) {
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < kExtDim; ++d) sc = fmaf(q[d], ks[j * kExtDim + d], sc);
      den += expf(sc - mx);
    }
    const float inv = 1.f / den;
    float acc[kExtDim];
#pragma unroll
    for (int d = 0; d < kExtDim; ++d) acc[d] = 0.f;
    const float* pb = a.pos_bias ? a.pos_bias + ((long)h * a.S + i) * a.S : nullptr;
    for (int j = 0; j < a.S; ++j) {
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < kExtDim; ++d) sc = fmaf(q[d], ks[j * kExtDim + d], sc);
      float p = expf(sc - mx) * inv;
      if (focus && j != i) p = -FLT_MAX;  // jnp.where(mask, attn, finfo(float32).min), AFTER the softmax
      if (pb) p += pb[j];                 // attn += pos_bias, AFTER the softmax
#pragma unroll
      for (int d = 0; d < kExtDim; ++d) acc[d] = fmaf(p, vs[j * kExtDim + d], acc[d]);
    }
#pragma unroll
    for (int d = 0; d < kExtDim; ++d) ext_st(op + d, acc[d]);
  }
}

// RelativePositionBias.__call__(n) -> (heads, n, n): bucket ids (int32) + gather from the (32, heads) embedding.
// Bucket of rel = i - j with the STATIC defaults num_buckets 32 / max_distance 128 (modules.py:386 ignores the
// constructor arguments):  m = -rel; ret = (m < 0) * 16; m = |m|;
//   ret += m < 8 ? m : min(15, 8 + int(log(m / 8) / log(16) * 8)).
// The logarithmic branch is evaluated in exact integer arithmetic: 8 * log16(m / 8) = log2(m^2 / 64), so
// int(...) = floor(log2(m * m)) - 6. That equals the float32 formula wherever the float32 log is correctly rounded
// (checked against the oracle for |rel| < 300) and does not depend on a device libm.
__global__ void rel_pos_bias_kernel(const float* __restrict__ emb, int n, int heads, float* __restrict__ out,
                                    int* __restrict__ buckets) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int i = idx / n, j = idx - i * n;
  int m = -(i - j);
  int ret = m < 0 ? 16 : 0;
  m = m < 0 ? -m : m;
  if (m < 8) {
    ret += m;
  } else {
    const long long sq = (long long)m * m;
    const int k = 63 - __clzll(sq) - 6;
    ret += min(15, 8 + k);
  }
  if (buckets) buckets[idx] = ret;
  for (int h = 0; h < heads; ++h) out[((long)h * n + i) * n + j] = emb[ret * heads + h];
}

template <typename T, int D>
static int launch_ext(const void* qkv, void* o, const ExtArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)a.S * D * 2 * sizeof(float);
  VDN_REQUIRE(smem <= 200 * 1024, VDN_E_SHAPE, "mha_core_ext: sequence length %d too long for shared memory", a.S);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mha_core_ext_fwd_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "mha_core_ext cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  mha_core_ext_fwd_kernel<T, D><<<dim3(a.n_seq, a.heads), kExtThreads, smem, st>>>(reinterpret_cast<const T*>(qkv),
                                                                                  reinterpret_cast<T*>(o), a);
  return check_launch("mha_core_ext_fwd");
}
template <typename T>
static int launch_ext_d(const void* qkv, void* o, const ExtArgs& a, int D, cudaStream_t st) {
  switch (D) {
    case 8: return launch_ext<T, 8>(qkv, o, a, st);
    case 16: return launch_ext<T, 16>(qkv, o, a, st);
    case 32: return launch_ext<T, 32>(qkv, o, a, st);
    case 64: return launch_ext<T, 64>(qkv, o, a, st);
  }
  set_last_error("mha_core_ext: head dimension %d not in {8, 16, 32, 64}", D);
  return VDN_E_SHAPE;
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_mha_core_ext_fwd(const void* qkv, void* o, int dtype, int heads, int dim, int n_seq, int S,
                                    int inner, const unsigned char* mask_b, int seqs_per_batch,
                                    const float* pos_bias, int copy_v, void* stream) {
  VDN_REQUIRE(qkv && o && n_seq > 0 && S > 0 && inner > 0 && heads > 0, VDN_E_SHAPE, "mha_core_ext_fwd: bad arguments");
  VDN_REQUIRE(n_seq % inner == 0, VDN_E_SHAPE, "mha_core_ext_fwd: n_seq must be a multiple of inner");
  VDN_REQUIRE(!mask_b || seqs_per_batch > 0, VDN_E_SHAPE, "mha_core_ext_fwd: seqs_per_batch missing");
  VDN_REQUIRE(dtype == VDN_BF16 || dtype == VDN_F32, VDN_E_SHAPE, "mha_core_ext_fwd: dtype");
  ExtArgs a;
  a.n_seq = n_seq; a.S = S; a.inner = inner; a.seqs_per_batch = seqs_per_batch > 0 ? seqs_per_batch : 1;
  a.heads = heads; a.scale = 1.0f / sqrtf((float)dim);
  a.mask_b = mask_b; a.pos_bias = pos_bias; a.copy_v = copy_v;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == VDN_F32 ? launch_ext_d<float>(qkv, o, a, dim, st) : launch_ext_d<bf16>(qkv, o, a, dim, st);
}

extern "C" int vdn_rel_pos_bias(const float* embedding, int n, int heads, float* out, int* buckets_out, void* stream) {
  VDN_REQUIRE(embedding && out && n > 0 && heads > 0, VDN_E_SHAPE, "rel_pos_bias: bad arguments");
  const int total = n * n;
  rel_pos_bias_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(embedding, n, heads, out,
                                                                                             buckets_out);
  return check_launch("rel_pos_bias");
}
