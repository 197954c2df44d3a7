// Small / bandwidth-bound kernels of the hot path:
//   init_conv (1,k,k) with tiny Cin (unet3d.py:110-115) fwd + wgrad, final 1x1 conv to `channels`
//   (unet3d.py:251) fwd + bwd, time-embedding MLP and the per-ResnetBlock time heads
//   (unet3d.py:128-133, modules.py:202-208,233-238) fwd + bwd, q_sample / loss / p_sample
//   (gaussian_diffusion.py:401-470,120-261), bias-gradient column sums, fused Adam + EMA
//   (trainer.py:367-382).
#include <algorithm>
#include <cstdlib>

#include <curand_kernel.h>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

__device__ __forceinline__ float block_sum(float v, float* red /*>=32 floats*/) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  r = red[0];
  return r;
}

// ---------------------------------------------------------------------------------------
// init conv: x fp32 (B,Cin,F,H,W) -> out bf16 (B*F,H,W,Cout); 16x16 pixel tile per block.
// ---------------------------------------------------------------------------------------
constexpr int kIT = 16;
__global__ void __launch_bounds__(256) init_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, bf16* __restrict__ out,
                                                            int B, int Cin, int F, int H, int W, int Cout, int ks) {
  extern __shared__ float sm[];
  const int pad = ks / 2, hs = kIT + ks - 1;
  float* sW = sm;                             // [ks*ks*Cin][Cout]
  float* sX = sm + ks * ks * Cin * Cout;      // [Cin][hs][hs]
  const int tiles_x = W / kIT;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
  const int img = blockIdx.y, b = img / F, f = img % F;
  for (int i = threadIdx.x; i < ks * ks * Cin * Cout; i += blockDim.x) sW[i] = w[i];
  for (int i = threadIdx.x; i < Cin * hs * hs; i += blockDim.x) {
    const int ci = i / (hs * hs), r = i % (hs * hs);
    const int yy = ty * kIT + r / hs - pad, xx = tx * kIT + r % hs - pad;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((((long)b * Cin + ci) * F + f) * H + yy) * W + xx];
    sX[i] = v;
  }
  __syncthreads();
  const int py = threadIdx.x / kIT, px = threadIdx.x % kIT;
  const long opix = ((long)img * H + ty * kIT + py) * W + tx * kIT + px;
  for (int co0 = 0; co0 < Cout; co0 += 32) {
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = bias[co0 + j];
    for (int ci = 0; ci < Cin; ++ci)
      for (int t = 0; t < ks * ks; ++t) {
        const float xv = sX[(ci * hs + py + t / ks) * hs + px + t % ks];
        const float* wp = sW + (t * Cin + ci) * Cout + co0;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fmaf(xv, wp[j], acc[j]);
      }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      u.x = pack_bf16x2(acc[8 * q + 0], acc[8 * q + 1]);
      u.y = pack_bf16x2(acc[8 * q + 2], acc[8 * q + 3]);
      u.z = pack_bf16x2(acc[8 * q + 4], acc[8 * q + 5]);
      u.w = pack_bf16x2(acc[8 * q + 6], acc[8 * q + 7]);
      reinterpret_cast<uint4*>(out + opix * Cout + co0)[q] = u;
    }
  }
}

// wgrad + bias grad of the init conv. grid (tile groups, n_img); thread -> (co lane = tid%32, tap group tid/32).
__global__ void __launch_bounds__(256) init_conv_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                              float* __restrict__ dw, float* __restrict__ dbias, int B,
                                                              int Cin, int F, int H, int W, int Cout, int ks,
                                                              int tiles_per_block) {
  extern __shared__ float sm[];
  const int pad = ks / 2, hs = kIT + ks - 1;
  float* sX = sm;                       // [Cin][hs][hs]
  float* sD = sm + Cin * hs * hs;       // [256][33]
  const int tiles_x = W / kIT, n_tiles = tiles_x * (H / kIT);
  const int img = blockIdx.y, b = img / F, f = img % F;
  const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int n_tc = ks * ks * Cin;             // (tap, ci) pairs
  const int per_thr = (n_tc + 7) / 8;         // pairs per thread (<= 19 for ks=7,Cin=3)
  for (int co0 = 0; co0 < Cout; co0 += 32) {
    float acc[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) acc[j] = 0.f;
    float bsum = 0.f;
    for (int tt = 0; tt < tiles_per_block; ++tt) {
      const int tile = blockIdx.x * tiles_per_block + tt;
      if (tile >= n_tiles) break;
      const int ty = tile / tiles_x, tx = tile % tiles_x;
      __syncthreads();
      for (int i = threadIdx.x; i < Cin * hs * hs; i += blockDim.x) {
        const int ci = i / (hs * hs), r = i % (hs * hs);
        const int yy = ty * kIT + r / hs - pad, xx = tx * kIT + r % hs - pad;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((((long)b * Cin + ci) * F + f) * H + yy) * W + xx];
        sX[i] = v;
      }
      for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
        const int p = i >> 5, c = i & 31;
        const long opix = ((long)img * H + ty * kIT + p / kIT) * W + tx * kIT + p % kIT;
        sD[p * 33 + c] = __bfloat162float(dy[opix * Cout + co0 + c]);
      }
      __syncthreads();
      if (grp == 0)
        for (int p = 0; p < 256; ++p) bsum += sD[p * 33 + col];
#pragma unroll
      for (int j = 0; j < 20; ++j) {
        const int tc = grp + 8 * j;
        if (j < per_thr && tc < n_tc) {
          const int t = tc / Cin, ci = tc % Cin;
          const float* xr = sX + (ci * hs + t / ks) * hs + t % ks;
          // four independent chains: the loop is bound by the smem load latency, not by the FMAs
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
          for (int p = 0; p < 256; p += 4) {
            const float* xp = xr + (p / kIT) * hs + p % kIT;  // kIT is a multiple of 4: p..p+3 share a tile row
            a0 = fmaf(xp[0], sD[p * 33 + col], a0);
            a1 = fmaf(xp[1], sD[(p + 1) * 33 + col], a1);
            a2 = fmaf(xp[2], sD[(p + 2) * 33 + col], a2);
            a3 = fmaf(xp[3], sD[(p + 3) * 33 + col], a3);
          }
          acc[j] += (a0 + a1) + (a2 + a3);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      const int tc = grp + 8 * j;
      if (j < per_thr && tc < n_tc) atomicAdd(&dw[(long)tc * Cout + co0 + col], acc[j]);
    }
    if (grp == 0) atomicAdd(&dbias[co0 + col], bsum);
  }
}

// ---------------------------------------------------------------------------------------
// init conv, 7x7 specialisation (the reference default init_kernel_size, unet3d.py:64): register-blocked and with the
// tap loops resolved at compile time. The generic kernels above spend most of their issue slots on index arithmetic
// (t / ks, t % ks with a runtime ks) and shared-memory loads (9 LDS per 32 FMA forward, 2 LDS per FMA in the wgrad).
//   forward: 16 x 32 pixel tile per block, a thread owns 2 pixels x 32 output channels (10 LDS per 64 FMA).
//   wgrad:   a thread owns one kernel ROW (7 taps) x 4 output channels = 28 accumulators and slides a 7-wide window of
//            x along the tile row (1 LDS + 1 LDS.64 per 28 FMA); blocks are persistent over tiles, partial sums are
//            combined in shared memory and leave with ONE atomic per (tap, ci, co) per block.
// ---------------------------------------------------------------------------------------
constexpr int kI7 = 7, kI7TH = 16, kI7TW = 32;  // kernel size, forward tile
__global__ void __launch_bounds__(256) init_conv7_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, bf16* __restrict__ out,
                                                             int B, int Cin, int F, int H, int W, int Cout) {
  extern __shared__ float sm[];
  constexpr int hh = kI7TH + kI7 - 1, hwv = kI7TW + kI7 - 1;  // 22 x 38 halo tile
  constexpr int hw = 48;  // row pitch: the two pixel rows of a warp fall into disjoint bank halves
  float* sW = sm;                              // [49*Cin][Cout]
  float* sX = sm + kI7 * kI7 * Cin * Cout;     // [Cin][hh][hw]
  const int tiles_x = W / kI7TW;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
  const int img = blockIdx.y, b = img / F, f = img % F;
  for (int i = threadIdx.x; i < kI7 * kI7 * Cin * Cout; i += blockDim.x) sW[i] = w[i];
  for (int i = threadIdx.x; i < Cin * hh * hwv; i += blockDim.x) {
    const int ci = i / (hh * hwv), r = i % (hh * hwv);
    const int ry = r / hwv, rx = r % hwv;
    const int yy = ty * kI7TH + ry - 3, xx = tx * kI7TW + rx - 3;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((((long)b * Cin + ci) * F + f) * H + yy) * W + xx];
    sX[(ci * hh + ry) * hw + rx] = v;
  }
  __syncthreads();
  const int py = threadIdx.x >> 4, px = threadIdx.x & 15;  // pixels (py, px) and (py, px + 16)
  const long opix = ((long)img * H + ty * kI7TH + py) * W + tx * kI7TW + px;
  for (int co0 = 0; co0 < Cout; co0 += 32) {
    float a0[32], a1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a0[j] = a1[j] = __ldg(bias + co0 + j);
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xb = sX + (ci * hh + py) * hw + px;
#pragma unroll
      for (int dy = 0; dy < kI7; ++dy)
#pragma unroll
        for (int dx = 0; dx < kI7; ++dx) {
          const float x0 = xb[dy * hw + dx], x1 = xb[dy * hw + dx + 16];
          const float4* wp = reinterpret_cast<const float4*>(sW + ((dy * kI7 + dx) * Cin + ci) * Cout + co0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 wv = wp[q];
            a0[4 * q + 0] = fmaf(x0, wv.x, a0[4 * q + 0]); a1[4 * q + 0] = fmaf(x1, wv.x, a1[4 * q + 0]);
            a0[4 * q + 1] = fmaf(x0, wv.y, a0[4 * q + 1]); a1[4 * q + 1] = fmaf(x1, wv.y, a1[4 * q + 1]);
            a0[4 * q + 2] = fmaf(x0, wv.z, a0[4 * q + 2]); a1[4 * q + 2] = fmaf(x1, wv.z, a1[4 * q + 2]);
            a0[4 * q + 3] = fmaf(x0, wv.w, a0[4 * q + 3]); a1[4 * q + 3] = fmaf(x1, wv.w, a1[4 * q + 3]);
          }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u, v;
      u.x = pack_bf16x2(a0[8 * q + 0], a0[8 * q + 1]); v.x = pack_bf16x2(a1[8 * q + 0], a1[8 * q + 1]);
      u.y = pack_bf16x2(a0[8 * q + 2], a0[8 * q + 3]); v.y = pack_bf16x2(a1[8 * q + 2], a1[8 * q + 3]);
      u.z = pack_bf16x2(a0[8 * q + 4], a0[8 * q + 5]); v.z = pack_bf16x2(a1[8 * q + 4], a1[8 * q + 5]);
      u.w = pack_bf16x2(a0[8 * q + 6], a0[8 * q + 7]); v.w = pack_bf16x2(a1[8 * q + 6], a1[8 * q + 7]);
      reinterpret_cast<uint4*>(out + opix * Cout + co0)[q] = u;
      reinterpret_cast<uint4*>(out + (opix + 16) * Cout + co0)[q] = v;
    }
  }
}

// block = 7 kernel rows x 32 lanes = 224 threads; lane -> (pixel range, group of 4 output channels):
// cg_n = min(Cout, 128) / 4 channel groups, 32 / cg_n pixel ranges (each a band of rows of the 16 x 16 tile).
constexpr int kW7Threads = 224;
__global__ void __launch_bounds__(kW7Threads) init_conv7_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                                      float* __restrict__ dw, float* __restrict__ dbias,
                                                                      int B, int Cin, int F, int H, int W, int Cout,
                                                                      int n_tiles_total) {
  extern __shared__ float sm[];
  constexpr int hs = kIT + kI7 - 1;  // 22
  const int cb = min(Cout, 128);     // output channels per pass
  constexpr int xp = hs + 3;         // odd row pitch: the pixel ranges of a warp read different banks
  float* sX = sm;                                                      // [hs][xp] (one input channel at a time)
  bf16* sD = reinterpret_cast<bf16*>(sm + 552);                        // [256 px][cb]   (552 >= hs * xp, 16-byte aligned)
  float* sR = reinterpret_cast<float*>(sD + 256 * cb);                 // [7 rows][7 taps][cb] cross-range reduction
  const int krow = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg_n = cb >> 2, n_rng = 32 / cg_n, rows_per_rng = kIT / n_rng;
  const int cg = lane % cg_n, rng = lane / cg_n;
  const int tiles_x = W / kIT, tiles_img = tiles_x * (H / kIT);
  for (int co0 = 0; co0 < Cout; co0 += cb)
    for (int ci = 0; ci < Cin; ++ci) {
      float acc[kI7][4];
      float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < kI7; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int img = tile / tiles_img, tr = tile % tiles_img;
        const int ty = tr / tiles_x, tx = tr % tiles_x;
        const int b = img / F, f = img % F;
        __syncthreads();
        for (int i = threadIdx.x; i < hs * hs; i += blockDim.x) {
          const int ry = i / hs, rx = i % hs;
          const int yy = ty * kIT + ry - 3, xx = tx * kIT + rx - 3;
          float v = 0.f;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((((long)b * Cin + ci) * F + f) * H + yy) * W + xx];
          sX[ry * xp + rx] = v;
        }
        {
          const int segs = cb >> 3;  // 16-byte segments per pixel
          for (int i = threadIdx.x; i < 256 * segs; i += blockDim.x) {
            const int p = i / segs, sg = i % segs;
            const long opix = ((long)img * H + ty * kIT + p / kIT) * W + tx * kIT + p % kIT;
            reinterpret_cast<uint4*>(sD)[i] = __ldg(reinterpret_cast<const uint4*>(dy + opix * Cout + co0) + sg);
          }
        }
        __syncthreads();
        for (int r = 0; r < rows_per_rng; ++r) {
          const int py = rng * rows_per_rng + r;
          const float* xr = sX + (py + krow) * xp;
          float win[kI7];
#pragma unroll
          for (int t = 0; t < kI7 - 1; ++t) win[t + 1] = xr[t];
#pragma unroll
          for (int px = 0; px < kIT; ++px) {
#pragma unroll
            for (int t = 0; t < kI7 - 1; ++t) win[t] = win[t + 1];
            win[kI7 - 1] = xr[px + kI7 - 1];
            const uint2 d2 = *reinterpret_cast<const uint2*>(sD + (py * kIT + px) * cb + cg * 4);
            const float2 d01 = unpack_bf16x2(d2.x), d23 = unpack_bf16x2(d2.y);
#pragma unroll
            for (int t = 0; t < kI7; ++t) {
              acc[t][0] = fmaf(win[t], d01.x, acc[t][0]);
              acc[t][1] = fmaf(win[t], d01.y, acc[t][1]);
              acc[t][2] = fmaf(win[t], d23.x, acc[t][2]);
              acc[t][3] = fmaf(win[t], d23.y, acc[t][3]);
            }
            if (krow == 0 && ci == 0) {
              bsum[0] += d01.x; bsum[1] += d01.y; bsum[2] += d23.x; bsum[3] += d23.y;
            }
          }
        }
      }
      // combine the pixel ranges of the block, then one atomic per (tap, ci, co)
      for (int q = 0; q < n_rng; ++q) {
        __syncthreads();
        if (rng == q) {
#pragma unroll
          for (int t = 0; t < kI7; ++t)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float* dst = sR + (krow * kI7 + t) * cb + cg * 4 + c;
              *dst = (q == 0 ? 0.f : *dst) + acc[t][c];
            }
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < kI7 * kI7 * cb; i += blockDim.x) {
        const int tap = i / cb, c = i % cb;
        atomicAdd(&dw[((long)tap * Cin + ci) * Cout + co0 + c], sR[i]);
      }
      if (ci == 0) {
        for (int q = 0; q < n_rng; ++q) {
          __syncthreads();
          if (krow == 0 && rng == q) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float* dst = sR + cg * 4 + c;
              *dst = (q == 0 ? 0.f : *dst) + bsum[c];
            }
          }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < cb; c += blockDim.x) atomicAdd(&dbias[co0 + c], sR[c]);
      }
    }
}

// ---------------------------------------------------------------------------------------
// final 1x1 conv: h bf16 [P][C] -> out fp32 [P][Co] (Co = image channels, <= 4)
// ---------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(256) final_conv_fwd_kernel(const bf16* __restrict__ h, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             long P, int C) {
  extern __shared__ float sW[];  // [C][CO]
  for (int i = threadIdx.x; i < C * CO; i += blockDim.x) sW[i] = w[i];
  __syncthreads();
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = bias[o];
    for (int c0 = 0; c0 < C; c0 += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(h + p * C + c0);
      float v[8];
      float2 f2;
      f2 = unpack_bf16x2(u.x); v[0] = f2.x; v[1] = f2.y;
      f2 = unpack_bf16x2(u.y); v[2] = f2.x; v[3] = f2.y;
      f2 = unpack_bf16x2(u.z); v[4] = f2.x; v[5] = f2.y;
      f2 = unpack_bf16x2(u.w); v[6] = f2.x; v[7] = f2.y;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = fmaf(v[j], sW[(c0 + j) * CO + o], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) out[p * CO + o] = acc[o];
  }
}

// dh = dout W^T (bf16), dW[c][o] += sum_p h[p][c] dout[p][o], db[o] += sum_p dout[p][o].
// Threads are organised as (pixel lane, channel-vector): tid = pl * (C/8) + ci.
template <int CO>
__global__ void __launch_bounds__(256) final_conv_bwd_kernel(const bf16* __restrict__ h, const float* __restrict__ dout,
                                                             const float* __restrict__ w, bf16* __restrict__ dh,
                                                             float* __restrict__ dw, float* __restrict__ db, long P,
                                                             int C) {
  extern __shared__ float sm[];
  float* sW = sm;            // [C][CO]
  float* red = sm + C * CO;  // [256][8*CO]
  for (int i = threadIdx.x; i < C * CO; i += blockDim.x) sW[i] = w[i];
  __syncthreads();
  const int c8n = C / 8, pl_n = blockDim.x / c8n;
  const int ci = threadIdx.x % c8n, pl = threadIdx.x / c8n, c0 = ci * 8;
  float gw[8][CO], gb[CO];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 0; o < CO; ++o) gw[j][o] = 0.f;
#pragma unroll
  for (int o = 0; o < CO; ++o) gb[o] = 0.f;
  if (pl < pl_n) {
    for (long p = (long)blockIdx.x * pl_n + pl; p < P; p += (long)gridDim.x * pl_n) {
      float d[CO];
#pragma unroll
      for (int o = 0; o < CO; ++o) d[o] = dout[p * CO + o];
      const uint4 u = *reinterpret_cast<const uint4*>(h + p * C + c0);
      float v[8], g[8];
      float2 f2;
      f2 = unpack_bf16x2(u.x); v[0] = f2.x; v[1] = f2.y;
      f2 = unpack_bf16x2(u.y); v[2] = f2.x; v[3] = f2.y;
      f2 = unpack_bf16x2(u.z); v[4] = f2.x; v[5] = f2.y;
      f2 = unpack_bf16x2(u.w); v[6] = f2.x; v[7] = f2.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          a = fmaf(d[o], sW[(c0 + j) * CO + o], a);
          gw[j][o] = fmaf(v[j], d[o], gw[j][o]);
        }
        g[j] = a;
      }
      uint4 q;
      q.x = pack_bf16x2(g[0], g[1]); q.y = pack_bf16x2(g[2], g[3]);
      q.z = pack_bf16x2(g[4], g[5]); q.w = pack_bf16x2(g[6], g[7]);
      *reinterpret_cast<uint4*>(dh + p * C + c0) = q;
      if (ci == 0) {
#pragma unroll
        for (int o = 0; o < CO; ++o) gb[o] += d[o];
      }
    }
  }
  float* my = red + (long)threadIdx.x * 8 * CO;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 0; o < CO; ++o) my[j * CO + o] = gw[j][o];
  __syncthreads();
  for (int i = threadIdx.x; i < C * CO; i += blockDim.x) {
    const int c = i / CO, o = i % CO;
    float a = 0.f;
    for (int q = 0; q < pl_n; ++q) a += red[(long)(q * c8n + c / 8) * 8 * CO + (c % 8) * CO + o];
    atomicAdd(&dw[i], a);
  }
  __syncthreads();
  if (ci == 0 && pl < pl_n) {
#pragma unroll
    for (int o = 0; o < CO; ++o) red[pl * CO + o] = gb[o];
  }
  __syncthreads();
  if (threadIdx.x < CO) {
    float a = 0.f;
    for (int q = 0; q < pl_n; ++q) a += red[q * CO + threadIdx.x];
    atomicAdd(&db[threadIdx.x], a);
  }
}

// ---------------------------------------------------------------------------------------
// time embedding MLP: sinusoid(dim) -> Linear(dim,4dim) -> gelu_tanh -> Linear(4dim,4dim)
// One block per batch row. Saves emb, h1 (pre-GELU) for the backward.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k = 0.7978845608028654f;
  return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float gelu_tanh_grad_f(float x) {
  const float k = 0.7978845608028654f;
  const float u = k * (x + 0.044715f * x * x * x);
  const float th = tanhf(u);
  const float du = k * (1.f + 3.f * 0.044715f * x * x);
  return 0.5f * (1.f + th) + 0.5f * x * (1.f - th * th) * du;
}

__global__ void __launch_bounds__(256) time_mlp_fwd_kernel(const int* __restrict__ time, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, float* __restrict__ emb_out,
                                                           float* __restrict__ h1_out, float* __restrict__ t_out,
                                                           int dim) {
  extern __shared__ float sm[];
  const int td = 4 * dim;
  float* emb = sm;        // [dim]
  float* g = sm + dim;    // [4dim]
  const int b = blockIdx.x;
  const float tv = (float)time[b];
  const int half = dim / 2;
  const float ef = logf(10000.f) / (float)(half - 1);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float ang = tv * expf((float)i * -ef);
    emb[i] = sinf(ang);
    emb[half + i] = cosf(ang);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dim; i += blockDim.x) emb_out[(long)b * dim + i] = emb[i];
  for (int j = threadIdx.x; j < td; j += blockDim.x) {
    float a = b1[j], a1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < dim; k += 2) {
      a = fmaf(emb[k], w1[(long)k * td + j], a);
      a1 = fmaf(emb[k + 1], w1[(long)(k + 1) * td + j], a1);
    }
    a += a1;
    h1_out[(long)b * td + j] = a;
    g[j] = gelu_tanh_f(a);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < td; j += blockDim.x) {
    float a = b2[j], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < td; k += 4) {
      a = fmaf(g[k], w2[(long)k * td + j], a);
      a1 = fmaf(g[k + 1], w2[(long)(k + 1) * td + j], a1);
      a2 = fmaf(g[k + 2], w2[(long)(k + 2) * td + j], a2);
      a3 = fmaf(g[k + 3], w2[(long)(k + 3) * td + j], a3);
    }
    a = (a + a1) + (a2 + a3);
    t_out[(long)b * td + j] = a;
  }
}

// dt [B][4dim] -> grads of w1,b1,w2,b2. Single block (B is the per-GPU batch, tiny).
__global__ void __launch_bounds__(256) time_mlp_bwd_kernel(const float* __restrict__ dt, const float* __restrict__ emb,
                                                           const float* __restrict__ h1, const float* __restrict__ w2,
                                                           float* __restrict__ dw1, float* __restrict__ db1,
                                                           float* __restrict__ dw2, float* __restrict__ db2,
                                                           float* __restrict__ dh1_ws /*[B][4dim]*/, int B, int dim) {
  const int td = 4 * dim;
  // Every block first rebuilds dh1 (B x td, tiny) and gelu(h1), dt in shared memory, then owns a slice of the
  // weight-gradient elements: the work is a few hundred kFLOP, so the kernel is pure latency and one block of
  // 256 threads (the previous shape) took 120 us.
  extern __shared__ float smt[];
  float* s_dt = smt;                 // [B][td]
  float* s_g = smt + B * td;         // gelu(h1) [B][td]
  float* s_dh = smt + 2 * B * td;    // dh1 [B][td]
  for (int i = threadIdx.x; i < B * td; i += blockDim.x) {
    s_dt[i] = dt[i];
    s_g[i] = gelu_tanh_f(h1[i]);
  }
  __syncthreads();
  // dh1[b][k] = gelu'(h1) * sum_j w2[k][j] dt[b][j]: one warp per (b, k), lanes over j
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int i = warp; i < B * td; i += n_warps) {
    const int b = i / td, k = i % td;
    float a = 0.f;
    for (int j = lane; j < td; j += 32) a = fmaf(w2[(long)k * td + j], s_dt[b * td + j], a);
    a = warp_sum(a);
    if (lane == 0) {
      const float v = a * gelu_tanh_grad_f(h1[i]);
      s_dh[i] = v;
      if (blockIdx.x == 0) dh1_ws[i] = v;
    }
  }
  __syncthreads();
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long)gridDim.x * blockDim.x;
  // dw2[k][j] = sum_b gelu(h1[b][k]) dt[b][j] ; db2[j] = sum_b dt[b][j]
  for (long i = gtid; i < (long)td * td; i += gsz) {
    const int k = (int)(i / td), j = (int)(i % td);
    float a = 0.f;
    for (int b = 0; b < B; ++b) a = fmaf(s_g[b * td + k], s_dt[b * td + j], a);
    dw2[i] += a;
  }
  // dw1[k][j] = sum_b emb[b][k] dh1[b][j] ; db1[j] = sum_b dh1[b][j]
  for (long i = gtid; i < (long)dim * td; i += gsz) {
    const int k = (int)(i / td), j = (int)(i % td);
    float a = 0.f;
    for (int b = 0; b < B; ++b) a = fmaf(emb[(long)b * dim + k], s_dh[b * td + j], a);
    dw1[i] += a;
  }
  for (long j = gtid; j < td; j += gsz) {
    float a = 0.f, c = 0.f;
    for (int b = 0; b < B; ++b) {
      a += s_dt[b * td + j];
      c += s_dh[b * td + j];
    }
    db2[j] += a;
    db1[j] += c;
  }
}

}  // namespace vdn

namespace vdn {

// grid (n_heads, B)
__global__ void __launch_bounds__(256) time_heads_fwd_kernel(const float* __restrict__ t, const vdn_time_head* __restrict__ heads,
                                                             float* __restrict__ e_pre, float* __restrict__ ss, int ss_ld,
                                                             int td) {
  extern __shared__ float sm[];
  float* st = sm;          // silu(t_b) [td]
  float* red = sm + td;    // [32]
  const vdn_time_head hd = heads[blockIdx.x];
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < td; k += blockDim.x) st[k] = silu_f(t[(long)b * td + k]);
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int j = threadIdx.x; j < hd.n_out; j += blockDim.x) {
    // four independent chains and unrolled loads: the loop is a chain of L2-latency loads otherwise
    float a = hd.b[j], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < td; k += 4) {
      a = fmaf(st[k], hd.w[(long)k * hd.n_out + j], a);
      a1 = fmaf(st[k + 1], hd.w[(long)(k + 1) * hd.n_out + j], a1);
      a2 = fmaf(st[k + 2], hd.w[(long)(k + 2) * hd.n_out + j], a2);
      a3 = fmaf(st[k + 3], hd.w[(long)(k + 3) * hd.n_out + j], a3);
    }
    a = (a + a1) + (a2 + a3);
    e_pre[(long)b * ss_ld + hd.off + j] = a;
    s1 += a;
    s2 += a * a;
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  const float mean = s1 / (float)hd.n_out;
  const float rstd = rsqrtf(fmaxf(s2 / (float)hd.n_out - mean * mean, 0.f) + 1e-6f);
  for (int j = threadIdx.x; j < hd.n_out; j += blockDim.x) {
    const float a = e_pre[(long)b * ss_ld + hd.off + j];
    ss[(long)b * ss_ld + hd.off + j] = (a - mean) * rstd * hd.ln_g[j] + hd.ln_b[j];
  }
}

// Backward of the time heads, two kernels.
// A: grid (n_heads, B): LayerNorm backward of one (head, sample) row -> de; db / dln accumulate (atomics over B).
__global__ void __launch_bounds__(256) time_heads_bwd_ln_kernel(const vdn_time_head* __restrict__ heads,
                                                                const float* __restrict__ e_pre,
                                                                const float* __restrict__ dss, int ss_ld,
                                                                float* __restrict__ de_ws /*[B][ss_ld]*/) {
  __shared__ float red[32];
  const vdn_time_head hd = heads[blockIdx.x];
  const int b = blockIdx.y, n = hd.n_out;
  const float* e = e_pre + (long)b * ss_ld + hd.off;
  const float* dy = dss + (long)b * ss_ld + hd.off;
  float s1 = 0.f, s2 = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    s1 += e[j];
    s2 += e[j] * e[j];
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  const float mean = s1 / (float)n;
  const float rstd = rsqrtf(fmaxf(s2 / (float)n - mean * mean, 0.f) + 1e-6f);
  float a1 = 0.f, a2 = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float eh = (e[j] - mean) * rstd;
    const float gd = hd.ln_g[j] * dy[j];
    a1 += gd;
    a2 += gd * eh;
    atomicAdd(&hd.dln_g[j], dy[j] * eh);
    atomicAdd(&hd.dln_b[j], dy[j]);
  }
  a1 = block_sum(a1, red) / (float)n;
  a2 = block_sum(a2, red) / (float)n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float eh = (e[j] - mean) * rstd;
    const float v = rstd * (hd.ln_g[j] * dy[j] - a1 - eh * a2);
    de_ws[(long)b * ss_ld + hd.off + j] = v;
    atomicAdd(&hd.db[j], v);
  }
}

// B: grid (n_heads, td/8): 8 rows k of the head's weight: dW[k][:] += sum_b silu(t[b][k]) de[b][:],
//    dt[b][k] += silu'(t[b][k]) * sum_j W[k][j] de[b][j] (one warp per k, atomics over heads).
constexpr int kTHB = 8;  // max batch rows held in registers per pass
__global__ void __launch_bounds__(256) time_heads_bwd_w_kernel(const float* __restrict__ t,
                                                               const vdn_time_head* __restrict__ heads,
                                                               const float* __restrict__ de_ws, int ss_ld,
                                                               float* __restrict__ dt, int B, int td) {
  const vdn_time_head hd = heads[blockIdx.x];
  const int n = hd.n_out;
  const int k0 = blockIdx.y * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < B; b0 += kTHB) {
    const int nb = min(kTHB, B - b0);
    // dW rows k0..k0+7
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      float de[kTHB];
#pragma unroll
      for (int q = 0; q < kTHB; ++q) de[q] = q < nb ? de_ws[(long)(b0 + q) * ss_ld + hd.off + j] : 0.f;
      for (int kk = 0; kk < 8 && k0 + kk < td; ++kk) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < kTHB; ++q)
          if (q < nb) a = fmaf(silu_f(t[(long)(b0 + q) * td + k0 + kk]), de[q], a);
        hd.dw[(long)(k0 + kk) * n + j] += a;
      }
    }
    // dt: warp `warp` owns row k0 + warp
    const int k = k0 + warp;
    if (k < td) {
      float acc[kTHB];
#pragma unroll
      for (int q = 0; q < kTHB; ++q) acc[q] = 0.f;
      for (int j = lane; j < n; j += 32) {
        const float w = hd.w[(long)k * n + j];
#pragma unroll
        for (int q = 0; q < kTHB; ++q)
          if (q < nb) acc[q] = fmaf(w, de_ws[(long)(b0 + q) * ss_ld + hd.off + j], acc[q]);
      }
#pragma unroll
      for (int q = 0; q < kTHB; ++q) {
        const float v = warp_sum(acc[q]);
        if (lane == 0 && q < nb) atomicAdd(&dt[(long)(b0 + q) * td + k], v * silu_grad_f(t[(long)(b0 + q) * td + k]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// diffusion elementwise math. Images are fp32 (B,C,F,H,W); the Unet output is (B,F,H,W,C).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long bfhwc_index(long i, int C, long FHW) {
  // i indexes (b,c,r) with r in [0,FHW); returns the index of the same element in (b,r,c)
  const long r = i % FHW;
  const long bc = i / FHW;
  const long b = bc / C, c = bc % C;
  return (b * FHW + r) * C + c;
}

__global__ void q_sample_kernel(const float* __restrict__ x, const float* __restrict__ noise, const int* __restrict__ t,
                                const float* __restrict__ sa, const float* __restrict__ sb, float* __restrict__ out,
                                long n, long per_sample, int normalize) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int tb = t[i / per_sample];
    float xv = x[i];
    if (normalize) xv = xv * 2.f - 1.f;
    out[i] = sa[tb] * xv + sb[tb] * noise[i];
  }
}

// loss += mean(|pred-noise|^p); dpred = d loss / d pred (fp32, Unet output layout)
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ pred, const float* __restrict__ noise,
                                                   float* __restrict__ loss, float* __restrict__ dpred, long n, int C,
                                                   long FHW, int l1) {
  __shared__ float red[32];
  float acc = 0.f;
  const float inv_n = 1.f / (float)n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long j = (C == 1) ? i : bfhwc_index(i, C, FHW);
    const float d = pred[j] - noise[i];
    if (l1) {
      acc += fabsf(d);
      if (dpred) dpred[j] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_n;
    } else {
      acc += d * d;
      if (dpred) dpred[j] = 2.f * d * inv_n;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_n);
}

// x_{t-1} = c1*clip(rc*x - rm1*eps) + c2*x + (t!=0)*exp(0.5*logvar)*z   (gaussian_diffusion.py:120-261)
__global__ void p_sample_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z,
                                const int* __restrict__ t, const float* __restrict__ rc, const float* __restrict__ rm1,
                                const float* __restrict__ c1, const float* __restrict__ c2,
                                const float* __restrict__ logvar, float* __restrict__ out, long n, long per_sample,
                                int C, long FHW, int clip) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int tb = t[i / per_sample];
    const long j = (C == 1) ? i : bfhwc_index(i, C, FHW);
    const float xv = x[i];
    float x0 = rc[tb] * xv - rm1[tb] * eps[j];
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    const float mean = c1[tb] * x0 + c2[tb] * xv;
    const float nz = (tb == 0) ? 0.f : 1.f;
    out[i] = mean + nz * expf(0.5f * logvar[tb]) * z[i];
  }
}

// ---------------------------------------------------------------------------------------
// column sums (bias gradients): db[c] += sum_p dy[p][c], dy bf16 [P][C]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ db, long P, int C) {
  extern __shared__ float red[];  // [256][8]
  pdl_trigger();
  pdl_wait();
  const int c8n = C / 8, pl_n = blockDim.x / c8n;
  const int ci = threadIdx.x % c8n, pl = threadIdx.x / c8n;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (pl < pl_n) {
#pragma unroll 4
    for (long p = (long)blockIdx.x * pl_n + pl; p < P; p += (long)gridDim.x * pl_n) {
      const uint4 u = *reinterpret_cast<const uint4*>(dy + p * C + ci * 8);
      float2 f2;
      f2 = unpack_bf16x2(u.x); a[0] += f2.x; a[1] += f2.y;
      f2 = unpack_bf16x2(u.y); a[2] += f2.x; a[3] += f2.y;
      f2 = unpack_bf16x2(u.z); a[4] += f2.x; a[5] += f2.y;
      f2 = unpack_bf16x2(u.w); a[6] += f2.x; a[7] += f2.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = a[j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < pl_n; ++q) s += red[(q * c8n + c / 8) * 8 + (c & 7)];
    atomicAdd(&db[c], s);
  }
}

__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, long n8) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const uint4 ua = reinterpret_cast<const uint4*>(a)[i], ub = reinterpret_cast<const uint4*>(b)[i];
    float2 x, y;
    uint4 o;
    x = unpack_bf16x2(ua.x); y = unpack_bf16x2(ub.x); o.x = pack_bf16x2(x.x + y.x, x.y + y.y);
    x = unpack_bf16x2(ua.y); y = unpack_bf16x2(ub.y); o.y = pack_bf16x2(x.x + y.x, x.y + y.y);
    x = unpack_bf16x2(ua.z); y = unpack_bf16x2(ub.z); o.z = pack_bf16x2(x.x + y.x, x.y + y.y);
    x = unpack_bf16x2(ua.w); y = unpack_bf16x2(ub.w); o.w = pack_bf16x2(x.x + y.x, x.y + y.y);
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// ---------------------------------------------------------------------------------------
// Fused Adam (optax.adam defaults, bias-corrected) + EMA (trainer.py:367-382) over the flat state.
// hp (device): {lr, b1, b2, eps, 1-b1^t, 1-b2^t, ema_decay, do_ema, grad_scale}
// ---------------------------------------------------------------------------------------
// Optional global-norm clip (utils.py:127-152): with hp[9] = max_grad_norm > 0 and sq = sum g^2 over the raw flat
// gradient, l2 = sqrt(sq * gs^2 + hp[10]) and every gradient is scaled by min(max_grad_norm / (l2 + hp[10]), 1).
__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, float* __restrict__ ema, const float* __restrict__ hp,
                                const float* __restrict__ sq, long n) {
  const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], bc1 = hp[4], bc2 = hp[5], decay = hp[6];
  const bool do_ema = hp[7] != 0.f;
  float gs = hp[8];
  if (sq != nullptr && hp[9] > 0.f) {
    const float l2 = sqrtf(sq[0] * gs * gs + hp[10]);
    gs *= fminf(hp[9] / (l2 + hp[10]), 1.f);
  }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pi = p[i] - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
    if (do_ema) ema[i] = decay * ema[i] + (1.f - decay) * pi;
  }
}

// Standard normal draws (Philox4x32-10 counter RNG): element i of stream (seed, subsequence) is a pure
// function of (seed, subsequence, i), so results do not depend on the launch geometry or on how a
// sample batch is sharded over GPUs (gaussian_diffusion.py:254,309,445 draw these with jax.random).
__global__ void randn_kernel(float* __restrict__ out, long n, unsigned long long seed, unsigned long long subseq,
                             unsigned long long elem_offset) {
  const long n4 = (n + 3) / 4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    curandStatePhilox4_32_10_t st;
    // one Philox counter block (4 x 32-bit outputs) per group of 4 elements; `offset` counts outputs
    curand_init(seed, subseq, 4ull * (unsigned long long)(elem_offset / 4 + i), &st);
    const float4 z = curand_normal4(&st);
    const long b = i * 4;
    if (b + 0 < n) out[b + 0] = z.x;
    if (b + 1 < n) out[b + 1] = z.y;
    if (b + 2 < n) out[b + 2] = z.z;
    if (b + 3 < n) out[b + 3] = z.w;
  }
}

// Device-resident timestep variant for the captured sampling loop (gaussian_diffusion.py:311-316): the Philox
// subsequence is subseq_mul * t + subseq_add with t read from device memory, so one CUDA graph serves every
// timestep with no host-side argument changes.
__global__ void randn_t_kernel(float* __restrict__ out, long n, const unsigned long long* __restrict__ prm,
                               const int* __restrict__ t) {
  const unsigned long long seed = prm[0], elem_offset = prm[2];
  const unsigned long long subseq = prm[1] + (unsigned long long)t[0];
  const long n4 = (n + 3) / 4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, subseq, 4ull * (unsigned long long)(elem_offset / 4 + i), &st);
    const float4 z = curand_normal4(&st);
    const long b = i * 4;
    if (b + 0 < n) out[b + 0] = z.x;
    if (b + 1 < n) out[b + 1] = z.y;
    if (b + 2 < n) out[b + 2] = z.z;
    if (b + 3 < n) out[b + 3] = z.w;
  }
}
// t[b] -= 1 for the next replay of the sampling graph (the reverse loop counts T-1 .. 0)
__global__ void countdown_kernel(int* __restrict__ t, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) t[b] -= 1;
}

static int ew_grid(long n, int threads = 256) {
  return (int)std::max<long>(1, std::min<long>((n + threads - 1) / threads, (long)num_sms() * 8));
}

}  // namespace vdn

using namespace vdn;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int vdn_init_conv_fwd(const float* x, const float* w, const float* bias, void* out, int B, int Cin, int F,
                                 int H, int W, int Cout, int ks, void* stream) {
  VDN_REQUIRE(H % kIT == 0 && W % kIT == 0 && Cout % 32 == 0 && (ks & 1) && ks <= 7 && Cin >= 1 && Cin <= 4, VDN_E_SHAPE,
              "init_conv_fwd: unsupported shape H=%d W=%d Cout=%d ks=%d Cin=%d", H, W, Cout, ks, Cin);
  const bool generic_only = tune_on("VDN_INIT_CONV_GENERIC");
  if (ks == kI7 && H % kI7TH == 0 && W % kI7TW == 0 && !generic_only) {
    const size_t smem7 = (size_t)(kI7 * kI7 * Cin * Cout + Cin * (kI7TH + kI7 - 1) * 48) * sizeof(float);
    static bool cfg7 = false;
    if (!cfg7) {
      cudaFuncSetAttribute(init_conv7_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      cfg7 = true;
    }
    init_conv7_fwd_kernel<<<dim3((H / kI7TH) * (W / kI7TW), B * F), 256, smem7, ST(stream)>>>(
        x, w, bias, reinterpret_cast<bf16*>(out), B, Cin, F, H, W, Cout);
    return check_launch("init_conv7_fwd");
  }
  const int hs = kIT + ks - 1;
  const size_t smem = (size_t)(ks * ks * Cin * Cout + Cin * hs * hs) * sizeof(float);
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(init_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cfg = true;
  }
  init_conv_fwd_kernel<<<dim3((H / kIT) * (W / kIT), B * F), 256, smem, ST(stream)>>>(
      x, w, bias, reinterpret_cast<bf16*>(out), B, Cin, F, H, W, Cout, ks);
  return check_launch("init_conv_fwd");
}

extern "C" int vdn_init_conv_wgrad(const float* x, const void* dy, float* dw, float* dbias, int B, int Cin, int F,
                                   int H, int W, int Cout, int ks, void* stream) {
  VDN_REQUIRE(H % kIT == 0 && W % kIT == 0 && Cout % 32 == 0 && (ks & 1) && ks <= 7 && Cin >= 1 && Cin <= 3, VDN_E_SHAPE,
              "init_conv_wgrad: unsupported shape");
  const bool generic_only = tune_on("VDN_INIT_CONV_GENERIC");
  if (ks == kI7 && !generic_only && (Cout <= 128 ? (Cout == 32 || Cout == 64 || Cout == 128) : Cout % 128 == 0)) {
    const int cb = std::min(Cout, 128);
    const int hs7 = kIT + kI7 - 1;
    (void)hs7;
    const size_t smem7 = (size_t)552 * sizeof(float) + (size_t)256 * cb * 2 + (size_t)kI7 * kI7 * cb * sizeof(float);
    static bool cfg7 = false;
    if (!cfg7) {
      cudaFuncSetAttribute(init_conv7_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
      cfg7 = true;
    }
    const int n_total = (H / kIT) * (W / kIT) * B * F;
    // persistent blocks: an equal number of tiles each, about two blocks per SM
    const int per_block = std::max(1, ceil_div(n_total, 2 * num_sms()));
    const int grid = ceil_div(n_total, per_block);
    init_conv7_wgrad_kernel<<<grid, kW7Threads, smem7, ST(stream)>>>(x, reinterpret_cast<const bf16*>(dy), dw, dbias, B, Cin,
                                                                     F, H, W, Cout, n_total);
    return check_launch("init_conv7_wgrad");
  }
  const int hs = kIT + ks - 1;
  const size_t smem = (size_t)(Cin * hs * hs + 256 * 33) * sizeof(float);
  const int n_tiles = (H / kIT) * (W / kIT);
  const int tpb = 1;  // one 16x16 tile per block: 640 blocks at v2_2 (the kernel is latency bound, not atomic bound)
  init_conv_wgrad_kernel<<<dim3(ceil_div(n_tiles, tpb), B * F), 256, smem, ST(stream)>>>(
      x, reinterpret_cast<const bf16*>(dy), dw, dbias, B, Cin, F, H, W, Cout, ks, tpb);
  return check_launch("init_conv_wgrad");
}

extern "C" int vdn_final_conv_fwd(const void* h, const float* w, const float* bias, float* out, long P, int C, int Co,
                                  void* stream) {
  VDN_REQUIRE(C % 8 == 0 && Co >= 1 && Co <= 4, VDN_E_SHAPE, "final_conv_fwd: C=%d Co=%d unsupported", C, Co);
  const int grid = ew_grid(P);
  const size_t smem = (size_t)C * Co * sizeof(float);
  const bf16* hp = reinterpret_cast<const bf16*>(h);
  switch (Co) {
    case 1: final_conv_fwd_kernel<1><<<grid, 256, smem, ST(stream)>>>(hp, w, bias, out, P, C); break;
    case 2: final_conv_fwd_kernel<2><<<grid, 256, smem, ST(stream)>>>(hp, w, bias, out, P, C); break;
    case 3: final_conv_fwd_kernel<3><<<grid, 256, smem, ST(stream)>>>(hp, w, bias, out, P, C); break;
    default: final_conv_fwd_kernel<4><<<grid, 256, smem, ST(stream)>>>(hp, w, bias, out, P, C); break;
  }
  return check_launch("final_conv_fwd");
}

extern "C" int vdn_final_conv_bwd(const void* h, const float* dout, const float* w, void* dh, float* dw, float* db,
                                  long P, int C, int Co, void* stream) {
  VDN_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0 && Co >= 1 && Co <= 4, VDN_E_SHAPE, "final_conv_bwd: C=%d Co=%d unsupported", C, Co);
  const int pl_n = 256 / (C / 8);
  const int grid = (int)std::max<long>(1, std::min<long>((P + pl_n * 16 - 1) / (pl_n * 16), num_sms() * 4));
  const size_t smem = (size_t)(C * Co + 256 * 8 * Co) * sizeof(float);
  const bf16* hp = reinterpret_cast<const bf16*>(h);
  bf16* dhp = reinterpret_cast<bf16*>(dh);
  switch (Co) {
    case 1: final_conv_bwd_kernel<1><<<grid, 256, smem, ST(stream)>>>(hp, dout, w, dhp, dw, db, P, C); break;
    case 2: final_conv_bwd_kernel<2><<<grid, 256, smem, ST(stream)>>>(hp, dout, w, dhp, dw, db, P, C); break;
    case 3: final_conv_bwd_kernel<3><<<grid, 256, smem, ST(stream)>>>(hp, dout, w, dhp, dw, db, P, C); break;
    default: final_conv_bwd_kernel<4><<<grid, 256, smem, ST(stream)>>>(hp, dout, w, dhp, dw, db, P, C); break;
  }
  return check_launch("final_conv_bwd");
}

extern "C" int vdn_time_mlp_fwd(const int* time, const float* w1, const float* b1, const float* w2, const float* b2,
                                float* emb_out, float* h1_out, float* t_out, int B, int dim, void* stream) {
  VDN_REQUIRE(B > 0 && dim >= 4 && dim % 2 == 0, VDN_E_SHAPE, "time_mlp_fwd: bad shape");
  time_mlp_fwd_kernel<<<B, 256, (size_t)5 * dim * sizeof(float), ST(stream)>>>(time, w1, b1, w2, b2, emb_out, h1_out,
                                                                                t_out, dim);
  return check_launch("time_mlp_fwd");
}

extern "C" int vdn_time_mlp_bwd(const float* dt, const float* emb, const float* h1, const float* w2, float* dw1,
                                float* db1, float* dw2, float* db2, float* dh1_ws, int B, int dim, void* stream) {
  const int td = 4 * dim;
  const size_t smem = (size_t)3 * B * td * sizeof(float);
  VDN_REQUIRE(smem <= 48 * 1024, VDN_E_SHAPE, "time_mlp_bwd: B=%d dim=%d does not fit shared memory", B, dim);
  const int blocks = std::max(1, std::min(148, (td * td + 255) / 256));
  time_mlp_bwd_kernel<<<blocks, 256, smem, ST(stream)>>>(dt, emb, h1, w2, dw1, db1, dw2, db2, dh1_ws, B, dim);
  return check_launch("time_mlp_bwd");
}

extern "C" int vdn_time_heads_fwd(const float* t, const void* heads_dev, int n_heads, float* e_pre, float* ss,
                                  int ss_ld, int B, int td, void* stream) {
  VDN_REQUIRE(n_heads > 0 && B > 0, VDN_E_SHAPE, "time_heads_fwd: bad shape");
  time_heads_fwd_kernel<<<dim3(n_heads, B), 256, (size_t)(td + 32) * sizeof(float), ST(stream)>>>(
      t, reinterpret_cast<const vdn_time_head*>(heads_dev), e_pre, ss, ss_ld, td);
  return check_launch("time_heads_fwd");
}

extern "C" int vdn_time_heads_bwd(const float* t, const void* heads_dev, int n_heads, const float* e_pre,
                                  const float* dss, int ss_ld, float* de_ws, float* dt, int B, int td, void* stream) {
  cudaError_t e = cudaMemsetAsync(dt, 0, (size_t)B * td * sizeof(float), ST(stream));
  VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "time_heads_bwd memset: %s", cudaGetErrorString(e));
  const vdn_time_head* hd = reinterpret_cast<const vdn_time_head*>(heads_dev);
  time_heads_bwd_ln_kernel<<<dim3(n_heads, B), 256, 0, ST(stream)>>>(hd, e_pre, dss, ss_ld, de_ws);
  int rc = check_launch("time_heads_bwd_ln");
  if (rc) return rc;
  time_heads_bwd_w_kernel<<<dim3(n_heads, (td + 7) / 8), 256, 0, ST(stream)>>>(t, hd, de_ws, ss_ld, dt, B, td);
  return check_launch("time_heads_bwd_w");
}

extern "C" int vdn_q_sample(const float* x_start, const float* noise, const int* t, const float* sqrt_ac,
                            const float* sqrt_1mac, float* out, int B, long per_sample, int normalize, void* stream) {
  const long n = (long)B * per_sample;
  q_sample_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(x_start, noise, t, sqrt_ac, sqrt_1mac, out, n, per_sample, normalize);
  return check_launch("q_sample");
}

extern "C" int vdn_loss(const float* pred, const float* noise, float* loss, float* dpred, int B, int C, long FHW,
                        int l1, void* stream) {
  const long n = (long)B * C * FHW;
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), ST(stream));
  VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "loss memset: %s", cudaGetErrorString(e));
  loss_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(pred, noise, loss, dpred, n, C, FHW, l1);
  return check_launch("loss");
}

extern "C" int vdn_p_sample(const float* x, const float* eps, const float* z, const int* t, const float* recip,
                            const float* recipm1, const float* coef1, const float* coef2, const float* logvar,
                            float* out, int B, int C, long FHW, int clip, void* stream) {
  const long n = (long)B * C * FHW;
  p_sample_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(x, eps, z, t, recip, recipm1, coef1, coef2, logvar, out, n,
                                                      (long)C * FHW, C, FHW, clip);
  return check_launch("p_sample");
}

extern "C" int vdn_colsum(const void* dy, float* db, long P, int C, void* stream) {
  VDN_REQUIRE(C % 8 == 0 && C / 8 <= 256, VDN_E_SHAPE, "colsum: C=%d unsupported", C);
  const int pl_n = 256 / (C / 8);
  const int grid = (int)std::max<long>(1, std::min<long>((P + pl_n * 4 - 1) / (pl_n * 4), num_sms() * 8));
  cudaError_t le = launch_pdl(colsum_kernel, dim3(grid), dim3(256), 256 * 8 * sizeof(float), ST(stream), 1,
                              reinterpret_cast<const bf16*>(dy), db, P, C);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "colsum launch: %s", cudaGetErrorString(le));
  return check_launch("colsum");
}

extern "C" int vdn_add_bf16(const void* a, const void* b, void* out, long n, void* stream) {
  VDN_REQUIRE(n % 8 == 0, VDN_E_SHAPE, "add_bf16: n must be a multiple of 8");
  add_bf16_kernel<<<ew_grid(n / 8), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(a),
                                                          reinterpret_cast<const bf16*>(b),
                                                          reinterpret_cast<bf16*>(out), n / 8);
  return check_launch("add_bf16");
}

extern "C" int vdn_randn(float* out, long n, unsigned long long seed, unsigned long long subseq,
                         unsigned long long elem_offset, void* stream) {
  VDN_REQUIRE(out && n > 0 && elem_offset % 4 == 0, VDN_E_SHAPE, "randn: bad args (elem_offset must be a multiple of 4)");
  randn_kernel<<<ew_grid((n + 3) / 4), 256, 0, ST(stream)>>>(out, n, seed, subseq, elem_offset);
  return check_launch("randn");
}

extern "C" int vdn_randn_t(float* out, long n, const unsigned long long* params_dev, const int* t_dev, void* stream) {
  VDN_REQUIRE(out && t_dev && params_dev && n > 0, VDN_E_SHAPE, "randn_t: bad args");
  randn_t_kernel<<<ew_grid((n + 3) / 4), 256, 0, ST(stream)>>>(out, n, params_dev, t_dev);
  return check_launch("randn_t");
}

extern "C" int vdn_countdown(int* t_dev, int B, void* stream) {
  VDN_REQUIRE(t_dev && B > 0, VDN_E_SHAPE, "countdown: bad args");
  countdown_kernel<<<(B + 127) / 128, 128, 0, ST(stream)>>>(t_dev, B);
  return check_launch("countdown");
}

extern "C" int vdn_adam_ema(float* p, const float* g, float* m, float* v, float* ema, const float* hp_dev, long n,
                            void* stream) {
  adam_ema_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(p, g, m, v, ema, hp_dev, nullptr, n);
  return check_launch("adam_ema");
}

// sum of squares of a flat fp32 buffer, accumulated (+=) into out[0]: block reduction + one atomic per block
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long n, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  const long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 q = g4[i];
    acc = fmaf(q.x, q.x, fmaf(q.y, q.y, fmaf(q.z, q.z, fmaf(q.w, q.w, acc))));
  }
  if (blockIdx.x == 0)
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) acc = fmaf(g[i], g[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

extern "C" int vdn_grad_sqnorm(const float* g, long n, float* out, void* stream) {
  VDN_REQUIRE(g && out && n > 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, VDN_E_SHAPE, "grad_sqnorm: bad args");
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), ST(stream));
  VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "grad_sqnorm memset: %s", cudaGetErrorString(e));
  sqnorm_kernel<<<std::min<long>(148 * 4, (n / 4 + 255) / 256 + 1), 256, 0, ST(stream)>>>(g, n, out);
  return check_launch("grad_sqnorm");
}

extern "C" int vdn_adam_ema_clip(float* p, const float* g, float* m, float* v, float* ema, const float* hp_dev,
                                 const float* sqnorm_dev, long n, void* stream) {
  adam_ema_kernel<<<ew_grid(n), 256, 0, ST(stream)>>>(p, g, m, v, ema, hp_dev, sqnorm_dev, n);
  return check_launch("adam_ema_clip");
}
