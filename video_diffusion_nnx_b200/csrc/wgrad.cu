// Weight-gradient GEMM on tcgen05: dW[tap][ci][co] += sum_pixels A_tap[p][ci] * G[p][co]
// (the wgrad of every conv / projection on the path; reference: jax.value_and_grad at
// trainer.py:361 differentiating modules.py:71-91,162-165,219-222,261-276 and utils.py:113,125).
//
// GEMM view: M = (tap, source, channel chunk) "atoms" of CW <= 64 input channels, N = output
// channels, K = pixels. Both operands are MN-major in shared memory: a TMA box of
// (CW channels x 128 pixels) lands as 128 K-rows of CW*2 bytes, which is exactly the canonical
// MN-major swizzled UMMA layout, so activations and gradients are consumed straight from their
// channels-last HBM layout with no transpose. Up to 128/CW atoms (different taps = different
// shifted boxes, zero-filled at the borders) are stacked along M of one 128-row MMA.
// K (all pixels) is split across CTAs; partial tiles are reduced with fp32 atomics into dW.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 6;
constexpr int kKPix = 128;  // pixels per K step (one TMA box)

struct WgMaps {
  CUtensorMap m[4];  // M-side tensors (sources or parity views)
  CUtensorMap n;     // N-side tensor
};

struct WgArgs {
  int H, W, n_img;             // pixel grid
  int n_pix_tiles;             // ceil(P / 128)
  int tiles_per_split;
  int cw;                      // channels per M atom (16/32/64)
  int atoms_total;             // taps * n_src * chunks
  int atoms_per_tile;          // 128 / cw
  int n_src, chunks;           // chunks of cw channels per source
  int src_c;                   // channels per source
  int n_taps;
  signed char tap_dx[16], tap_dy[16], tap_map[16];   // M-side shift / map per tap
  signed char ntap_dx[16], ntap_dy[16];              // N-side shift per tap (0 unless `per_tap_n`)
  int per_tap_n;               // 1: the N-side box depends on the tap (then atoms_per_tile spans ONE tap)
  int BN, cwn, n_atoms_n;      // N tile, channels per N atom, atoms per N tile
  int stages, tmem_cols;
  float* out;
  long tap_stride, row_stride, col_stride;
  int tap_perm[16];
  // bias gradient (column sums of the N-side tensor) for free: a spare M atom of the last M tile is filled
  // with ones, so one accumulator row of that tile is sum_pixels g[p][n]
  float* dbias;
  int ones_tile;  // M tile holding the ones atom (its slot = first unused atom), or -1
};

template <int dummy>
__global__ void __launch_bounds__(kWgThreads) wgrad_kernel(const __grid_constant__ WgMaps maps, const WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int4 s_atom[8];  // per M atom: (tensor map, channel offset, dx, dy)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;
  // 1024-byte aligned view of the dynamic smem; offset arithmetic keeps the pointer in the shared space
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const int atom_bytes_m = kKPix * a.cw * 2;
  const int atom_bytes_n = kKPix * a.cwn * 2;
  const int a_bytes = a.atoms_per_tile * atom_bytes_m;  // 128 * 128 * 2 = 32 KB
  const int b_bytes = a.n_atoms_n * atom_bytes_n;
  const int stage_bytes = a_bytes + b_bytes;
  const int S = a.stages;

  const int m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
  const int atom0 = m_tile * a.atoms_per_tile;
  const int n_atoms = min(a.atoms_per_tile, a.atoms_total - atom0);
  const int t_begin = split * a.tiles_per_split;
  const int t_end = min(a.n_pix_tiles, t_begin + a.tiles_per_split);
  const int n_steps = t_end - t_begin;

  pdl_trigger();
  const bool ones_here = a.dbias != nullptr && m_tile == a.ones_tile;
  if (ones_here) {  // bf16 1.0 in every K row of the spare atom, in every stage (TMA never writes this slot)
    for (int st = 0; st < S; ++st) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(smem + st * stage_bytes + n_atoms * atom_bytes_m);
      for (int i = threadIdx.x; i < atom_bytes_m / 4; i += blockDim.x) dst[i] = 0x3F803F80u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // visible to the tensor-core (async) proxy
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // global memory (TMA loads, dW atomics) only below

  if (n_steps > 0) {
    if (warp == 0) {
      // per-atom TMA parameters (map, channel offset, shift) are computed once by the first lanes
      if (lane < n_atoms) {
        const int id = atom0 + lane;
        const int chunk = id % a.chunks;
        const int src = (id / a.chunks) % a.n_src;
        const int tap = id / (a.chunks * a.n_src);
        s_atom[lane] = make_int4(a.tap_map[tap] + src, chunk * a.cw, a.tap_dx[tap], a.tap_dy[tap]);
      }
      __syncwarp();
      if (elect_one()) {  // single elected lane: keeps the loop on the uniform datapath (no waterfall loops)
        const int hw = a.H * a.W;
        const int tap_n = a.per_tap_n ? (atom0 / (a.n_src * a.chunks)) : 0;
        const int ndx = a.ntap_dx[tap_n], ndy = a.ntap_dy[tap_n];
        const uint32_t tx_bytes = (uint32_t)(n_atoms * atom_bytes_m + b_bytes);
        const int p_begin = t_begin * kKPix;
        int n0 = p_begin / hw;
        int rem = p_begin - n0 * hw;
        int st = 0;
        uint32_t ph = 1u;
        for (int it = 0; it < n_steps; ++it) {
          const int y0 = rem / a.W, x0 = rem - y0 * a.W;
          mbar_wait(&empty_bar[st], ph);
          uint8_t* sa = smem + st * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          mbar_expect_tx(&full_bar[st], tx_bytes);
          for (int i = 0; i < n_atoms; ++i) {
            const int4 ai = s_atom[i];
            tma_load_4d(sa + i * atom_bytes_m, &maps.m[ai.x], &full_bar[st], ai.y, x0 + ai.z, y0 + ai.w, n0);
          }
          for (int j = 0; j < a.n_atoms_n; ++j)
            tma_load_4d(sb + j * atom_bytes_n, &maps.n, &full_bar[st], n_tile * a.BN + j * a.cwn, x0 + ndx, y0 + ndy, n0);
          rem += kKPix;
          while (rem >= hw) {
            rem -= hw;
            ++n0;
          }
          if (++st == S) {
            st = 0;
            ph ^= 1u;
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_bf16(128, a.BN, 1, 1);  // both operands MN-major
        const uint32_t lay_m = umma_layout_type(a.cw * 2), lay_n = umma_layout_type(a.cwn * 2);
        // descriptor halves: hi = SBO (stride between 8-row K groups) | version | swizzle; lo = addr | LBO (atom stride)
        const uint32_t hi_m = ((8u * a.cw * 2u) >> 4) | (1u << 14) | (lay_m << 29);
        const uint32_t hi_n = ((8u * a.cwn * 2u) >> 4) | (1u << 14) | (lay_n << 29);
        const uint32_t lbo_m = (((uint32_t)atom_bytes_m >> 4) & 0x3FFFu) << 16;
        const uint32_t lbo_n = (((uint32_t)atom_bytes_n >> 4) & 0x3FFFu) << 16;
        const uint32_t kstep_m = (uint32_t)(16 * a.cw * 2) >> 4, kstep_n = (uint32_t)(16 * a.cwn * 2) >> 4;
        const uint32_t s0 = smem_u32(smem) >> 4, stage16 = (uint32_t)stage_bytes >> 4, ab16 = (uint32_t)a_bytes >> 4;
        int st = 0;
        uint32_t ph = 0u, sa16 = s0;
        for (int it = 0; it < n_steps; ++it) {
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          uint32_t lo_m = sa16 | lbo_m, lo_n = (sa16 + ab16) | lbo_n;
#pragma unroll
          for (int k = 0; k < kKPix / 16; ++k) {
            // 16 pixels (two 8-row groups) per MMA: advance both operands by 16 rows
            umma_bf16(tmem_base, (static_cast<uint64_t>(hi_m) << 32) | lo_m, (static_cast<uint64_t>(hi_n) << 32) | lo_n,
                      idesc, (it | k) != 0 ? 1u : 0u);
            lo_m += kstep_m;
            lo_n += kstep_n;
          }
          tc_commit(&empty_bar[st]);
          sa16 += stage16;
          if (++st == S) {
            st = 0;
            ph ^= 1u;
            sa16 = s0;
          }
        }
        tc_commit(&tmem_full_bar);
      }
      __syncwarp();
    } else {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;  // accumulator row = (atom, channel in atom)
      mbar_wait(&tmem_full_bar, 0);
      tc_fence_after();
      const int ai = r / a.cw;
      const bool valid = ai < n_atoms;
      const int id = atom0 + (valid ? ai : 0);
      const int chunk = id % a.chunks;
      const int src = (id / a.chunks) % a.n_src;
      const int tap = min(id / (a.chunks * a.n_src), a.n_taps - 1);  // the ones-only tile has no tap of its own
      const long row = (long)src * a.src_c + chunk * a.cw + (r % a.cw);
      float* orow = a.out + a.tap_perm[tap] * a.tap_stride + row * a.row_stride + (long)(n_tile * a.BN) * a.col_stride;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const bool bias_row = ones_here && r == n_atoms * a.cw;  // first row of the ones atom = column sums of g
      for (int c0 = 0; c0 < a.BN; c0 += 16) {
        uint32_t raw[16];
        tmem_ld_32x16(taddr + (uint32_t)c0, raw);
        tmem_ld_wait();
        if (bias_row) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.dbias + n_tile * a.BN + c0 + j),
                         "f"(__uint_as_float(raw[j])), "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])),
                         "f"(__uint_as_float(raw[j + 3]))
                         : "memory");
        }
        if (valid) {
          if (a.col_stride == 1) {  // contiguous output row: 128-bit vector reductions (4x fewer L2 atomics)
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c0 + j), "f"(__uint_as_float(raw[j])),
                           "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])),
                           "f"(__uint_as_float(raw[j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(orow + (long)(c0 + j) * a.col_stride, __uint_as_float(raw[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// Test-only reference (CUDA cores): one thread per dW element, loops over all pixels.
struct WgRefArgs {
  int kind, n_img, H, W, n_src, C, n_taps, Cout, transposed_roles;
  int tap_dy[16], tap_dx[16];
  const bf16* src[2];
  const bf16* g;
  float* out;
};
__global__ void wgrad_ref_kernel(const WgRefArgs a) {
  const int cin_tot = a.n_src * a.C;
  const long total = (long)a.n_taps * cin_tot * a.Cout;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % a.Cout);
  const int ci = (int)((idx / a.Cout) % cin_tot);
  const int t = (int)(idx / ((long)a.Cout * cin_tot));
  const int s = ci / a.C, c = ci % a.C;
  float acc = 0.f;
  for (int n = 0; n < a.n_img; ++n)
    for (int y = 0; y < a.H; ++y)
      for (int x = 0; x < a.W; ++x) {
        if (a.kind == VDN_TAP_UP) {
          // transposed conv: dW[a,b][ci][co] = sum x[i,j][ci] * g[2i+2-a, 2j+2-b][co]; g on the 2H x 2W grid
          const int gy = 2 * y + 2 - a.tap_dy[t], gx = 2 * x + 2 - a.tap_dx[t];
          if (gy < 0 || gy >= 2 * a.H || gx < 0 || gx >= 2 * a.W) continue;
          acc += __bfloat162float(a.src[s][(((long)n * a.H + y) * a.W + x) * a.C + c]) *
                 __bfloat162float(a.g[(((long)n * 2 * a.H + gy) * 2 * a.W + gx) * a.Cout + co]);
        } else {
          int sy, sx, SH, SW;
          if (a.kind == VDN_TAP_DOWN) {
            SH = 2 * a.H; SW = 2 * a.W; sy = 2 * y + a.tap_dy[t] - 1; sx = 2 * x + a.tap_dx[t] - 1;
          } else {
            SH = a.H; SW = a.W; sy = y + a.tap_dy[t]; sx = x + a.tap_dx[t];
          }
          if (sy < 0 || sy >= SH || sx < 0 || sx >= SW) continue;
          acc += __bfloat162float(a.src[s][(((long)n * SH + sy) * SW + sx) * a.C + c]) *
                 __bfloat162float(a.g[(((long)n * a.H + y) * a.W + x) * a.Cout + co]);
        }
      }
  a.out[idx] += acc;
}

static int cw_for(int c) { return (c % 64 == 0) ? 64 : (c % 32 == 0) ? 32 : 16; }

}  // namespace vdn

using namespace vdn;

// kind VDN_TAP_UNIT: src (n_img,H,W,C) x n_src, g (n_img,H,W,Cout), taps = (dy,dx) shifts
// kind VDN_TAP_DOWN: src (n_img,2H,2W,C),       g (n_img,H,W,Cout), taps = kernel indices (ky,kx) in 0..3
// kind VDN_TAP_UP  : src (n_img,H,W,C),         g (n_img,2H,2W,Cout), taps = kernel indices (a,b) in 0..3
// dw: fp32 [n_taps][n_src*C][Cout] (reference kernel layout), accumulated into (+=).
extern "C" int vdn_colsum(const void* dy, float* db, long P, int C, void* stream);
extern "C" int vdn_wgrad_bias(int kind, const void* src0, const void* src1, const void* g, float* dw, float* dbias,
                              int n_img, int H, int W, int n_src, int C, int Cout, int n_taps, const int* tap_dy,
                              const int* tap_dx, void* stream);

extern "C" int vdn_wgrad(int kind, const void* src0, const void* src1, const void* g, float* dw, int n_img, int H,
                         int W, int n_src, int C, int Cout, int n_taps, const int* tap_dy, const int* tap_dx,
                         void* stream) {
  return vdn_wgrad_bias(kind, src0, src1, g, dw, nullptr, n_img, H, W, n_src, C, Cout, n_taps, tap_dy, tap_dx, stream);
}

extern "C" int vdn_wgrad_bias(int kind, const void* src0, const void* src1, const void* g, float* dw, float* dbias,
                              int n_img, int H, int W, int n_src, int C, int Cout, int n_taps, const int* tap_dy,
                              const int* tap_dx, void* stream) {
  VDN_REQUIRE(kind >= 0 && kind <= 2 && src0 && g && dw, VDN_E_SHAPE, "wgrad: bad args");
  VDN_REQUIRE(n_taps >= 1 && n_taps <= 16 && (n_src == 1 || (n_src == 2 && src1 && kind == VDN_TAP_UNIT)), VDN_E_SHAPE,
              "wgrad: bad taps/sources");
  VDN_REQUIRE(C % 16 == 0 && Cout % 16 == 0, VDN_E_SHAPE, "wgrad: C=%d Cout=%d must be multiples of 16", C, Cout);
  auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
  VDN_REQUIRE((pow2(W) || W % 128 == 0) && (pow2(H) || W >= 128), VDN_E_SHAPE, "wgrad: H=%d W=%d unsupported", H, W);

  // roles: normally M side = activations (ci), N side = gradient (co). For the transposed conv the
  // shifted/strided tensor is the gradient, so it takes the M side and the store is transposed.
  const bool swap = kind == VDN_TAP_UP;
  const int Cm = swap ? Cout : C;   // channels on the M side (per source)
  const int Cn = swap ? C : Cout;   // channels on the N side
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.H = H; a.W = W; a.n_img = n_img;
  const long P = (long)n_img * H * W;
  a.n_pix_tiles = (int)((P + kKPix - 1) / kKPix);
  a.cw = cw_for(Cm);
  a.n_src = n_src;
  a.chunks = Cm / a.cw;
  a.src_c = Cm;
  a.n_taps = n_taps;
  a.atoms_total = n_taps * n_src * a.chunks;
  a.atoms_per_tile = 128 / a.cw;
  a.per_tap_n = 0;
  a.cwn = cw_for(Cn);
  a.BN = std::min(Cn, 128);
  while (Cn % a.BN != 0) a.BN -= a.cwn;
  VDN_REQUIRE(a.BN >= 16 && a.BN % 16 == 0 && a.BN % a.cwn == 0, VDN_E_SHAPE, "wgrad: cannot tile N=%d", Cn);
  a.n_atoms_n = a.BN / a.cwn;
  for (int t = 0; t < 16; ++t) a.tap_perm[t] = t;
  a.tap_stride = (long)n_src * C * Cout;
  if (!swap) { a.row_stride = Cout; a.col_stride = 1; } else { a.row_stride = 1; a.col_stride = Cout; }
  a.out = dw;

  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int bw = std::min(W, 128), bh = std::min(H, 128 / bw), bn = 128 / (bw * bh);
  int rc;
  auto enc_plain = [&](CUtensorMap* m, const void* base, int Cc, int cwid) {
    const uint64_t dims[4] = {(uint64_t)Cc, (uint64_t)W, (uint64_t)H, (uint64_t)n_img};
    const uint64_t str[3] = {(uint64_t)Cc * 2, (uint64_t)W * Cc * 2, (uint64_t)H * W * Cc * 2};
    const uint32_t box[4] = {(uint32_t)cwid, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    return encode_tmap_bf16(m, base, 4, dims, str, box, cwid * 2);
  };
  auto enc_parity = [&](CUtensorMap* m, const void* base0, int Cc, int cwid, int ry, int rx) {
    const uint64_t SW2 = 2 * (uint64_t)W, SH2 = 2 * (uint64_t)H;
    const uint64_t dims[4] = {(uint64_t)Cc, (uint64_t)W, (uint64_t)H, (uint64_t)n_img};
    const uint64_t str[3] = {2 * (uint64_t)Cc * 2, 2 * SW2 * Cc * 2, SH2 * SW2 * Cc * 2};
    const uint32_t box[4] = {(uint32_t)cwid, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    const uint8_t* base = reinterpret_cast<const uint8_t*>(base0) + ((uint64_t)ry * SW2 + rx) * Cc * 2;
    return encode_tmap_bf16(m, base, 4, dims, str, box, cwid * 2);
  };
  if (kind == VDN_TAP_UNIT) {
    const void* srcs[2] = {src0, src1};
    for (int s = 0; s < n_src; ++s)
      if ((rc = enc_plain(&maps.m[s], srcs[s], C, a.cw))) return rc;
    if ((rc = enc_plain(&maps.n, g, Cout, a.cwn))) return rc;
    for (int t = 0; t < n_taps; ++t) {
      a.tap_map[t] = 0;
      a.tap_dy[t] = (signed char)tap_dy[t];
      a.tap_dx[t] = (signed char)tap_dx[t];
    }
  } else if (kind == VDN_TAP_DOWN) {
    for (int ry = 0; ry < 2; ++ry)
      for (int rx = 0; rx < 2; ++rx)
        if ((rc = enc_parity(&maps.m[ry * 2 + rx], src0, C, a.cw, ry, rx))) return rc;
    if ((rc = enc_plain(&maps.n, g, Cout, a.cwn))) return rc;
    for (int t = 0; t < n_taps; ++t) {
      const int ky = tap_dy[t], kx = tap_dx[t];
      a.tap_map[t] = (signed char)(((ky + 1) & 1) * 2 + ((kx + 1) & 1));
      a.tap_dy[t] = (signed char)((ky - 1) >> 1);
      a.tap_dx[t] = (signed char)((kx - 1) >> 1);
    }
  } else {
    // M side: parity views of g (2H x 2W): g[2i+2-a] -> a=0: (r0,q+1) a=1: (r1,q0) a=2: (r0,q0) a=3: (r1,q-1)
    for (int ry = 0; ry < 2; ++ry)
      for (int rx = 0; rx < 2; ++rx)
        if ((rc = enc_parity(&maps.m[ry * 2 + rx], g, Cout, a.cw, ry, rx))) return rc;
    if ((rc = enc_plain(&maps.n, src0, C, a.cwn))) return rc;
    for (int t = 0; t < n_taps; ++t) {
      const int ky = tap_dy[t], kx = tap_dx[t];
      a.tap_map[t] = (signed char)((ky & 1) * 2 + (kx & 1));
      a.tap_dy[t] = (signed char)((2 - ky) >> 1);
      a.tap_dx[t] = (signed char)((2 - kx) >> 1);
    }
  }

  int m_tiles = ceil_div(a.atoms_total, a.atoms_per_tile);
  const int n_tiles = Cn / a.BN;
  // bias gradient: through a ones atom when the gradient is the N side - in the spare slot of the last M tile, or, when
  // the atoms fill their tiles exactly (1x1 projections of 128 / 256 channels), in an extra M tile that holds nothing
  // else: one more CTA column of a launch-latency-bound GEMM instead of a separate column-sum launch (26 per step)
  a.dbias = nullptr;
  a.ones_tile = -1;
  bool bias_by_colsum = false;
  if (dbias) {
    if (!swap && a.atoms_total % a.atoms_per_tile != 0) {
      a.dbias = dbias;
      a.ones_tile = m_tiles - 1;
    } else if (!swap && !tune_on("VDN_WG_COLSUM")) {
      a.dbias = dbias;
      a.ones_tile = m_tiles;
      m_tiles += 1;
    } else {
      bias_by_colsum = true;
    }
  }
  const int base_ctas = m_tiles * n_tiles;
  // split K (pixels) over CTAs to fill the machine, but keep >= 16 K steps per split: every split adds a
  // full set of fp32 reductions on the output tile (128-bit red.v4.f32 where the output row is contiguous) and a CTA's
  // prologue + first TMA round trip; measured on the training step: 4 -> 5.80 ms, 8 -> 5.77, 16 -> 5.74, 24 -> 5.74.
  const int min_k = std::max(1, tune_int("VDN_WG_MINK", 16));
  // CTAs per SM to aim for: 1 for the small problems of config_v2_2 (the weight gradients run on a side stream next to
  // the dependency chain; fewer, longer CTAs leave the chain more of the machine: 6.81 -> 6.75 ms per step), 2 once the
  // GEMM is large enough to be throughput bound by itself (v2_3x: 81.7 vs 84.8 ms per step)
  const int fill_env = tune_is_set("VDN_WG_FILL") ? std::max(1, tune_int("VDN_WG_FILL", 0)) : 0;
  const double work = (double)P * a.atoms_total * a.cw * Cn;
  const int fill = fill_env ? fill_env : (work >= 2e10 ? 2 : 1);
  // never past `fill` CTAs per SM: shared memory admits one (or two) CTAs per SM, so 150 CTAs are two waves - the whole
  // launch then takes as long as 296 would (the 3 x 50 splits of the 64x64-level conv, the 18 x 2 x 5 of the 8x8 level)
  const int want = tune_on("VDN_WG_CEIL") ? (fill * num_sms() + base_ctas - 1) / base_ctas : (fill * num_sms()) / base_ctas;
  int splits = std::max(1, std::min(std::max(1, a.n_pix_tiles / min_k), want));
  a.tiles_per_split = ceil_div(a.n_pix_tiles, splits);
  splits = ceil_div(a.n_pix_tiles, a.tiles_per_split);
  const int a_bytes = a.atoms_per_tile * kKPix * a.cw * 2;
  const int b_bytes = a.n_atoms_n * kKPix * a.cwn * 2;
  const int stage_bytes = a_bytes + b_bytes;
  a.stages = std::max(2, std::min(kWgMaxStages, (196 * 1024) / stage_bytes));
  int cols = 32;
  while (cols < a.BN) cols *= 2;
  a.tmem_cols = cols;
  const int smem_bytes = a.stages * stage_bytes + 1024;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "wgrad cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cfg = true;
  }
  cudaError_t le = launch_pdl(wgrad_kernel<0>, dim3(m_tiles, n_tiles, splits), dim3(kWgThreads), (size_t)smem_bytes,
                              reinterpret_cast<cudaStream_t>(stream), 1, maps, a);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "wgrad launch: %s", cudaGetErrorString(le));
  rc = check_launch("wgrad_kernel");
  if (rc == 0 && bias_by_colsum) {
    const long Pg = kind == VDN_TAP_UP ? 4 * P : P;  // the gradient of the transposed conv lives on the 2H x 2W grid
    rc = vdn_colsum(g, dbias, Pg, Cout, stream);
  }
  return rc;
}

extern "C" int vdn_wgrad_ref(int kind, const void* src0, const void* src1, const void* g, float* dw, int n_img, int H,
                             int W, int n_src, int C, int Cout, int n_taps, const int* tap_dy, const int* tap_dx,
                             void* stream) {
  WgRefArgs a;
  memset(&a, 0, sizeof(a));
  a.kind = kind; a.n_img = n_img; a.H = H; a.W = W; a.n_src = n_src; a.C = C; a.n_taps = n_taps; a.Cout = Cout;
  for (int t = 0; t < n_taps; ++t) { a.tap_dy[t] = tap_dy[t]; a.tap_dx[t] = tap_dx[t]; }
  a.src[0] = reinterpret_cast<const bf16*>(src0);
  a.src[1] = reinterpret_cast<const bf16*>(src1);
  a.g = reinterpret_cast<const bf16*>(g);
  a.out = dw;
  const long total = (long)n_taps * n_src * C * Cout;
  wgrad_ref_kernel<<<(unsigned)((total + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("wgrad_ref_kernel");
}
