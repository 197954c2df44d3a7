// Tap-GEMM: TMA-fed implicit-GEMM on tcgen05 tensor cores with TMEM accumulators.
//
// One kernel family serves every convolution / projection of the Unet3D hot path
// (reference call sites: modules.py:71-91,162-165,219-222,261-276; utils.py:113,125):
//
//   out[m, n] = sum_{tap, src, c} A_src[pixel(m) + shift(tap), c] * Wp[n, (tap,src,c)]
//
// Rows m are pixels of an (n_img, H, W) grid, 128 per CTA. For each (tap, src, 64-channel
// chunk) the producer warp issues ONE 4-D TMA box load (chunk, bw, bh, bn) at the shifted
// coordinate; out-of-image elements are zero-filled by the TMA unit (SAME padding), and the
// box lands in shared memory already in the canonical K-major 128B-swizzled UMMA layout.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-5 = epilogue
// (tcgen05.ld -> bias / residual / GroupNorm partial sums -> bf16 or fp32 store).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
// Epilogue warps of the one-tile kernel: 4 (one per TMEM lane quarter). Eight (two per quarter, half the columns each)
// drain a tile faster - 9.8 instead of 11.4 us on the 8x8-level conv - but 320 threads x 130 registers leave room for
// ONE CTA per SM, and every launch with more CTAs than SMs lost its co-resident partner (16x16 level 11.8 -> 16.2 us,
// training step 6.69 -> 7.34 ms): not kept.
// Both exist as template instances: EW = 4 with four co-resident CTAs per SM for launches with more CTAs than SMs, EW = 8
// with one CTA per SM (no register cap, no spills) for launches that cannot fill the SMs anyway (the small-M levels).
constexpr int kEpiWarpsWide = 8;

struct TapMaps {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap o[2];  // persistent kernel only: outputs [M][ld] (second: columns >= split_col), box (min(BN,64), 128)
  CUtensorMap r[2];  // persistent kernel with a residual: the residual tensors, same geometry as o[]
};

struct TapArgs {
  int M, N, BN;
  int H, W;          // row grid
  int bw, bh, bn;    // TMA box (pixels); bw*bh*bn == 128
  int n_taps, n_src, chunks;  // chunks of BK channels per source
  int stages;
  int tmem_cols;
  int n_maps;
  signed char tap_dx[16], tap_dy[16], tap_map[16];
  // epilogue
  const float* bias;
  const void* res;
  const void* res2;
  void* out;
  void* out2;
  int split_col;
  int ld_out, ld_out2;
  int out_f32;
  int scatter, py, px;  // VDN_TAP_UP parity scatter into a 2H x 2W grid
  float* gn_sums;
  int gn_groups, cpg, rows_per_sample, n_samples;
  int n_ntiles, n_items, stg_bufs;  // persistent kernel: (M tile, N tile) items, staging buffers
  int res_tma;       // generic kernel: byte offset (in dynamic smem) of the TMA-fetched residual tile, or 0
  int splits;        // split-K kernel: K ranges per output tile = cluster size along z
  float* ws;         // split-K kernel: fp32 partial tiles [splits][m_tiles][n_tiles][128][BN]
  long long* trace;  // debug (vdn_debug_tapgemm_trace), 1024 entries: clock64 stamps of CTA (0,0): [4][64] = producer0 / producer1 / MMA /
                     // epilogue, then [256][3] global-timer stamps (prologue done, predecessor complete, exit) of the first 256 CTAs
};

// ---------------------------------------------------------------------------------------
// epilogue helpers
// ---------------------------------------------------------------------------------------
template <int CPG16>
__device__ __forceinline__ void gn_accumulate16(const float (&v)[16], bool valid, float* gsum_base, int group0,
                                                int lane) {
  // v: 16 consecutive output channels of this thread's row; groups of CPG16 channels.
#pragma unroll
  for (int j = 0; j < 16; j += CPG16) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CPG16; ++k) {
      float x = valid ? v[j + k] : 0.f;
      s1 += x;
      s2 += x * x;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      float* p = gsum_base + 2 * (group0 + j / CPG16);
      atomicAdd(p, s1);
      atomicAdd(p + 1, s2);
    }
  }
}

// Same reduction, but lane 0 accumulates into this warp's private smem slots (no atomics).
template <int CPG16>
__device__ __forceinline__ void gn_accumulate16_warp(const float (&v)[16], bool valid, float* slot, int lane) {
#pragma unroll
  for (int j = 0; j < 16; j += CPG16) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CPG16; ++k) {
      float x = valid ? v[j + k] : 0.f;
      s1 += x;
      s2 += x * x;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      slot[2 * (j / CPG16)] += s1;
      slot[2 * (j / CPG16) + 1] += s2;
    }
  }
}

template <int BK, int kEpiWarps>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, kEpiWarps == 8 ? 1 : 4)
tapgemm_kernel(const __grid_constant__ TapMaps maps, const TapArgs args) {
  constexpr int kEpiThreads = 32 * kEpiWarps;      // warps: 0 TMA producer, 1 MMA issuer, 2.. epilogue
  constexpr int kSwizzle = BK * 2;                 // bytes per smem row == swizzle span
  constexpr int kABytes = kTileM * BK * 2;         // 16 KB / 8 KB / 4 KB
  constexpr uint32_t kLayout = umma_layout_type(kSwizzle);
  constexpr uint32_t kSBO = 8 * kSwizzle;          // byte stride between 8-row groups

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ __align__(8) uint64_t res_bar;  // residual tile landed (args.res_tma)
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_gn[kEpiWarps][16];  // per epilogue warp: (sum, sumsq) of up to 8 groups of this N tile
  __shared__ float s_bias[256];  // bias row of this N tile

  // warp-uniform role index (shfl broadcast): keeps the producer / MMA loops on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int BN = args.BN;
  const int b_bytes = (BN * BK * 2 + 1023) & ~1023;
  const int stage_bytes = kABytes + b_bytes;
  // 1024-byte aligned view of the dynamic smem; offset arithmetic keeps the pointer in the shared space
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));

  const int n_tile = blockIdx.x;
  const int m0 = blockIdx.y * kTileM;
  const int S = args.stages;
  const int n_steps = args.n_taps * args.n_src * args.chunks;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < args.n_maps; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&res_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, (uint32_t)args.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // debug timeline of every CTA (global timer, ns): prologue done / predecessor complete / exit
  const int trace_cta = blockIdx.y * gridDim.x + blockIdx.x;
  const bool trace_on = args.trace != nullptr && threadIdx.x == 0 && trace_cta < 256;
  if (trace_on) args.trace[256 + 3 * trace_cta] = (long long)globaltimer_ns();
  pdl_wait();  // the prologue above overlapped the previous kernel's tail; global memory is touched below
  if (trace_on) args.trace[256 + 3 * trace_cta + 1] = (long long)globaltimer_ns();

  if (warp == 0) {
    // ================= TMA producer =================
    // One elected lane per producer warp issues; `elect_one` (not `lane == 0`) lets the compiler treat the region as a
    // single thread, so descriptors / coordinates stay in uniform registers instead of per-instruction waterfall loops.
    if (elect_one()) {
      // one producer thread: a second producer warp on alternate steps measured no gain - the K loop is bound by the
      // tensor pipe (220 cycles per step against a 208-cycle floor, tools/trace_tapgemm.py)
      const int hw = args.H * args.W;
      const int n0 = m0 / hw;
      const int rem = m0 - n0 * hw;
      const int y0 = rem / args.W;
      const int x0 = rem - y0 * args.W;
      const uint32_t tx_bytes = (uint32_t)(kABytes + BN * BK * 2);
      if (args.res_tma) {
        // the residual tile [128 rows][BN] (plain row-major) travels while the K loop runs; per-thread residual loads in
        // the write-out loop cost 2.6 - 4.3 us per launch at the 8x8 / 16x16 levels
        const int cb0 = n_tile * BN;
        const bool second = args.split_col > 0 && cb0 >= args.split_col;
        mbar_expect_tx(&res_bar, (uint32_t)(kTileM * BN * 2));
        tma_load_2d(smem + args.res_tma, second ? &maps.r[1] : &maps.r[0], &res_bar, second ? cb0 - args.split_col : cb0, m0);
      }
      int st = 0, kcol = 0, step = 0;
      uint32_t ph = 1u;  // parity to wait for on the empty barrier of the slot
      for (int t = 0; t < args.n_taps; ++t) {
        const int cx = x0 + args.tap_dx[t];
        const int cy = y0 + args.tap_dy[t];
        for (int s = 0; s < args.n_src; ++s) {
          const CUtensorMap* am = &maps.a[args.tap_map[t] + s];
          for (int c = 0; c < args.chunks; ++c, ++step) {
            {
              mbar_wait(&empty_bar[st], ph);
              if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && step < 64) args.trace[(step & 1) * 64 + step] = clock64();
              uint8_t* sa = smem + st * stage_bytes;
              uint8_t* sb = sa + kABytes;
              mbar_expect_tx(&full_bar[st], tx_bytes);
              tma_load_4d(sa, am, &full_bar[st], c * BK, cx, cy, n0);
              tma_load_2d(sb, &maps.b, &full_bar[st], kcol, n_tile * BN);
            }
            kcol += BK;
            if (++st == S) {
              st = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kTileM, BN, 0, 0);
      // descriptor halves: hi = SBO | version | swizzle mode; lo = (address >> 4) | LBO(16 B) << 16
      const uint32_t desc_hi = (kSBO >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (1u << 16);
      int st = 0;
      uint32_t ph = 0u;
      uint32_t a_lo = a_lo0;
      for (int it = 0; it < n_steps; ++it) {
        mbar_wait(&full_bar[st], ph);
        if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && it < 64) args.trace[2 * 64 + it] = clock64();
        tc_fence_after();
        const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 K-elements = 32 bytes inside the swizzle span (encoded >> 4)
          umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k),
                    (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k), idesc, (it | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[st]);  // frees the smem slot when these MMAs retire
        a_lo += stage16;
        if (++st == S) {
          st = 0;
          ph ^= 1u;
          a_lo = a_lo0;
        }
      }
      tc_commit(&tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..) =================
    // A one-tile CTA has nothing to overlap its epilogue with: on the 36-step K loops of the 8x8 level it used to be
    // 6900 of the CTA's 16200 cycles (tools/trace_tapgemm.py) - bias loads from global memory behind every TMEM load,
    // a TMEM wait per 32 columns, divisions in the write-out loop. Now the bias row of the tile is staged in shared
    // memory while the K loop runs, the TMEM load of the next 32 columns is in flight while the current 32 are
    // processed, and the write-out indexes with shifts. (With kEpiWarps = 8, two warps share a TMEM lane quarter and
    // drain half of the columns each.)
    const int ew = warp - 2;       // 0..kEpiWarps-1
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int chalf = ew >> 2;     // which half of the columns (kEpiWarps = 8, tiles of >= 64 columns)
    const int r = quarter * 32 + lane;
    const int m = m0 + r;
    const bool valid = m < args.M;
    const int et = threadIdx.x - 64;  // epilogue thread id 0..255
    const int col_base = n_tile * BN;
    if (et < 16 * kEpiWarps) s_gn[et >> 4][et & 15] = 0.f;
    for (int c = et; c < BN; c += kEpiThreads) s_bias[c] = args.bias ? __ldg(args.bias + col_base + c) : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && et == 0) args.trace[3 * 64 + 0] = clock64();
    mbar_wait(&tmem_full_bar, 0);
    if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && et == 0) args.trace[3 * 64 + 1] = clock64();
    tc_fence_after();

    uint8_t* outp;
    const uint8_t* resp;
    int ld, col_o;
    if (args.split_col > 0 && col_base >= args.split_col) {
      outp = reinterpret_cast<uint8_t*>(args.out2);
      resp = reinterpret_cast<const uint8_t*>(args.res2);
      ld = args.ld_out2;
      col_o = col_base - args.split_col;
    } else {
      outp = reinterpret_cast<uint8_t*>(args.out);
      resp = reinterpret_cast<const uint8_t*>(args.res);
      ld = args.ld_out;
      col_o = col_base;
    }
    const int esz = args.out_f32 ? 4 : 2;
    const bool staged = !args.scatter;  // stage the tile in smem and store coalesced rows
    // direct (scatter) mode: output row of this thread
    long orow = m;
    if (args.scatter) {
      const int hw = args.H * args.W;
      const int n = m / hw;
      const int rem = m - n * hw;
      const int y = rem / args.W;
      const int x = rem - y * args.W;
      orow = ((long)n * (2 * args.H) + (2 * y + args.py)) * (2 * args.W) + (2 * x + args.px);
    }
    // GroupNorm partial sums: CTA-level reduction in smem when the tile lies in one sample, else per warp
    const bool gn_on = args.gn_sums != nullptr;
    const bool gn_cta = gn_on && (args.rows_per_sample % kTileM == 0);
    float* gdst = nullptr;
    if (gn_on) {
      const int sample = min(m0 + (gn_cta ? 0 : quarter * 32), args.M - 1) / args.rows_per_sample;
      const int rep = blockIdx.y % kGnReplicas;
      gdst = args.gn_sums + ((long)(rep * args.n_samples + sample) * args.gn_groups) * 2;
    }
    const int g_tile0 = col_base / args.cpg;  // first group covered by this N tile
    // groups of a power-of-two >= 16 channels that tile BN evenly (every ResnetBlock conv with >= 128 channels)
    const bool gn_wide = gn_cta && args.cpg >= 16 && (args.cpg & (args.cpg - 1)) == 0 && BN % args.cpg == 0;
    const int cpg_log = 31 - __clz(args.cpg);
    float ga[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) ga[q] = 0.f;
    const int pitch = BN * esz + 16;          // smem row pitch of the staged tile (bytes)
    uint8_t* srow = smem + (size_t)r * pitch;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // columns of this warp: half of the tile (tiles below 64 columns are drained by the first warp of the quarter)
    const bool halves = kEpiWarps == 8 && BN >= 64;
    const int cb = halves ? chalf * (BN >> 1) : 0;
    const int ce = halves ? cb + (BN >> 1) : (chalf == 0 ? BN : 0);

    // processes 16 consecutive accumulator columns starting at tile column cl
    auto process16 = [&](const uint32_t* raw16, int cl) {
      float v[16];
      const int cg = col_base + cl;  // global output column of v[0]
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw16[j]) + s_bias[cl + j];
      if (gn_on) {
        const int cpg = args.cpg;
        if (gn_wide) {
          // per-thread partial sums over the whole tile (four independent chains per 16 columns), ONE 16-shuffle
          // warp reduction after the column loop: a shuffle reduction per chunk sits on the serial path of the epilogue
          float p1[4], p2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float x = valid ? v[q] : 0.f;
            p1[q] = x;
            p2[q] = x * x;
          }
#pragma unroll
          for (int j = 4; j < 16; ++j) {
            const float x = valid ? v[j] : 0.f;
            p1[j & 3] += x;
            p2[j & 3] = fmaf(x, x, p2[j & 3]);
          }
          const float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
          const int g = cl >> cpg_log;  // group within this N tile (warp-uniform)
#pragma unroll
          for (int gi = 0; gi < 8; ++gi)
            if (gi == g) {
              ga[2 * gi] += s1;
              ga[2 * gi + 1] += s2;
            }
        } else if (gn_cta) {
          float* slot = &s_gn[ew][2 * (cg / cpg - g_tile0)];
          if (cpg >= 16) gn_accumulate16_warp<16>(v, valid, slot, lane);
          else if (cpg == 8) gn_accumulate16_warp<8>(v, valid, slot, lane);
          else if (cpg == 4) gn_accumulate16_warp<4>(v, valid, slot, lane);
          else gn_accumulate16_warp<2>(v, valid, slot, lane);
        } else {
          float* base = gdst + 2 * (cg / cpg);
          if (cpg >= 16) gn_accumulate16<16>(v, valid, base, 0, lane);
          else if (cpg == 8) gn_accumulate16<8>(v, valid, base, 0, lane);
          else if (cpg == 4) gn_accumulate16<4>(v, valid, base, 0, lane);
          else gn_accumulate16<2>(v, valid, base, 0, lane);
        }
      }
      if (staged) {
        if (args.out_f32) {
          float4* sp = reinterpret_cast<float4*>(srow + cl * 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) sp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          uint4* sp = reinterpret_cast<uint4*>(srow + cl * 2);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 q;
            q.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            sp[j] = q;
          }
        }
      } else {
        const long eoff = orow * ld + (col_o + cl);
        if (resp && valid) {
          if (args.out_f32) {
            const float4* rp = reinterpret_cast<const float4*>(resp + eoff * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 q = rp[j];
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          } else {
            const uint4* rp = reinterpret_cast<const uint4*>(resp + eoff * 2);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint4 q = rp[j];
              float2 f;
              f = unpack_bf16x2(q.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
              f = unpack_bf16x2(q.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
              f = unpack_bf16x2(q.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
              f = unpack_bf16x2(q.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
            }
          }
        }
        if (valid) {
          if (args.out_f32) {
            float4* op = reinterpret_cast<float4*>(outp + eoff * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(outp + eoff * 2);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4 q;
              q.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              op[j] = q;
            }
          }
        }
      }
    };

    if (cb < ce) {
      // software pipeline over 16-column chunks: the TMEM load of the next chunk is in flight while the current one is
      // processed. (Two 32-column buffers would be one TMEM instruction fewer per pair, but cost 32 more registers: at
      // 130 registers only two CTAs fit on an SM instead of four, and every launch with more CTAs than SMs slowed down -
      // training step 6.69 -> 6.99 ms. This version stays within the register budget of four co-resident CTAs.)
      uint32_t raw_a[16], raw_b[16];
      tmem_ld_32x16(taddr + (uint32_t)cb, raw_a);
      for (int c0 = cb; c0 < ce; c0 += 32) {
        tmem_ld_wait();
        const bool two = c0 + 16 < ce;
        if (two) tmem_ld_32x16(taddr + (uint32_t)(c0 + 16), raw_b);
        process16(raw_a, c0);
        if (two) {
          tmem_ld_wait();
          if (c0 + 32 < ce) tmem_ld_32x16(taddr + (uint32_t)(c0 + 32), raw_a);
          process16(raw_b, c0 + 16);
        }
      }
    }
    if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && et == 0) args.trace[3 * 64 + 4] = clock64();
    if (gn_wide) {
      const float tot = warp_sum16(ga, lane);
      if ((lane & 1) == 0) s_gn[ew][lane >> 1] = tot;
    }
    if (staged) {
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (args.res_tma) mbar_wait(&res_bar, 0);
      if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && et == 0) args.trace[3 * 64 + 5] = clock64();
      // coalesced write-out: consecutive threads store consecutive 16-byte segments of a row
      const int spr = (BN * esz) >> 4;  // 16B segments per row
      const int total = kTileM * spr;
      const bool spr_pow2 = (spr & (spr - 1)) == 0;
      const int spr_log = 31 - __clz(spr);
      for (int idx = et; idx < total; idx += kEpiThreads) {
        const int rr = spr_pow2 ? (idx >> spr_log) : idx / spr;
        const int sg = idx - rr * spr;
        const int mm = m0 + rr;
        if (mm >= args.M) break;
        uint4 q = *reinterpret_cast<const uint4*>(smem + (size_t)rr * pitch + sg * 16);
        const long goff = ((long)mm * ld + col_o) * esz + sg * 16;
        if (resp) {
          const uint4 rq = args.res_tma ? *reinterpret_cast<const uint4*>(smem + args.res_tma + (size_t)rr * (BN * 2) + sg * 16)
                                        : *reinterpret_cast<const uint4*>(resp + goff);
          if (args.out_f32) {
            q.x = __float_as_uint(__uint_as_float(q.x) + __uint_as_float(rq.x));
            q.y = __float_as_uint(__uint_as_float(q.y) + __uint_as_float(rq.y));
            q.z = __float_as_uint(__uint_as_float(q.z) + __uint_as_float(rq.z));
            q.w = __float_as_uint(__uint_as_float(q.w) + __uint_as_float(rq.w));
          } else {
            float2 a, b;
            a = unpack_bf16x2(q.x); b = unpack_bf16x2(rq.x); q.x = pack_bf16x2(a.x + b.x, a.y + b.y);
            a = unpack_bf16x2(q.y); b = unpack_bf16x2(rq.y); q.y = pack_bf16x2(a.x + b.x, a.y + b.y);
            a = unpack_bf16x2(q.z); b = unpack_bf16x2(rq.z); q.z = pack_bf16x2(a.x + b.x, a.y + b.y);
            a = unpack_bf16x2(q.w); b = unpack_bf16x2(rq.w); q.w = pack_bf16x2(a.x + b.x, a.y + b.y);
          }
        }
        *reinterpret_cast<uint4*>(outp + goff) = q;
      }
      if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && et == 0) args.trace[3 * 64 + 6] = clock64();
      if (gn_cta && et < 2 * ((BN + args.cpg - 1) / args.cpg)) {
        float tot = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < kEpiWarps; ++w8) tot += s_gn[w8][et];
        atomicAdd(gdst + 2 * g_tile0 + et, tot);
      }
    }
  }

  if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) args.trace[3 * 64 + 2] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)args.tmem_cols);
  }
  if (args.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) args.trace[3 * 64 + 3] = clock64();
  if (trace_on) args.trace[256 + 3 * trace_cta + 2] = (long long)globaltimer_ns();
}


// ---------------------------------------------------------------------------------------
// Persistent variant for launches with many tiles per SM and a short K loop (the q|k|v / out projections and
// their dgrads, 1x1 residual convs): there the one-tile-per-CTA kernel above is a serial chain per CTA
// (prologue -> TMA latency -> a handful of MMAs -> epilogue -> exit) that only co-resident CTAs overlap, and it
// measured ~3x off the HBM-write bound (N = 768, K = 128 at 1M pixels: 743 us against 250 us of output traffic).
// Here a CTA walks (M tile, N tile) items; the TMA producer runs ahead across items through the smem ring, the
// accumulator is double buffered in TMEM (the MMAs of item j+1 overlap the epilogue of item j), and the epilogue
// stages the bf16 tile in the swizzled layout of the output tensor map and hands it to the TMA store engine, so the
// HBM write is asynchronous. With a residual operand the staged tile is added and stored by the threads instead.
// No GroupNorm sums, no fp32 output, no parity scatter (those launches keep the kernel above).
// ---------------------------------------------------------------------------------------
constexpr int kPersistEpiThreads = 256;  // 8 epilogue warps
constexpr int kPersistThreads = 64 + kPersistEpiThreads;

template <int BK, bool kRes>
__global__ void __launch_bounds__(kPersistThreads) tapgemm_persist_kernel(const __grid_constant__ TapMaps maps,
                                                                       const TapArgs args) {
  constexpr int kSwizzle = BK * 2;
  constexpr int kABytes = kTileM * BK * 2;
  constexpr uint32_t kLayout = umma_layout_type(kSwizzle);
  constexpr uint32_t kSBO = 8 * kSwizzle;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t rfull_bar[2];  // residual tile landed in staging buffer b (kRes)
  __shared__ uint32_t tmem_base_smem;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int BN = args.BN;
  const int b_bytes = (BN * BK * 2 + 1023) & ~1023;
  const int stage_bytes = kABytes + b_bytes;
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const int S = args.stages;
  uint8_t* stg = smem + S * stage_bytes;
  const int n_steps = args.n_taps * args.n_src * args.chunks;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < args.n_maps; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.o[0]);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kPersistEpiThreads);
      mbar_init(&rfull_bar[b], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, (uint32_t)args.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      const int hw = args.H * args.W;
      const uint32_t tx_bytes = (uint32_t)(kABytes + BN * BK * 2);
      int st = 0;
      uint32_t ph = 1u;
      for (int item = blockIdx.x; item < args.n_items; item += gridDim.x) {
        const int n_tile = item % args.n_ntiles;
        const int m0 = (item / args.n_ntiles) * kTileM;
        const int n0 = m0 / hw;
        const int rem = m0 - n0 * hw;
        const int y0 = rem / args.W;
        const int x0 = rem - y0 * args.W;
        int kcol = 0;
        for (int t = 0; t < args.n_taps; ++t) {
          const int cx = x0 + args.tap_dx[t];
          const int cy = y0 + args.tap_dy[t];
          for (int s = 0; s < args.n_src; ++s) {
            const CUtensorMap* am = &maps.a[args.tap_map[t] + s];
            for (int c = 0; c < args.chunks; ++c) {
              mbar_wait(&empty_bar[st], ph);
              uint8_t* sa = smem + st * stage_bytes;
              mbar_expect_tx(&full_bar[st], tx_bytes);
              tma_load_4d(sa, am, &full_bar[st], c * BK, cx, cy, n0);
              tma_load_2d(sa + kABytes, &maps.b, &full_bar[st], kcol, n_tile * BN);
              kcol += BK;
              if (++st == S) {
                st = 0;
                ph ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kTileM, BN, 0, 0);
      const uint32_t desc_hi = (kSBO >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (1u << 16);
      int st = 0;
      uint32_t ph = 0u;
      uint32_t a_lo = a_lo0;
      uint32_t tph0 = 1u, tph1 = 1u;  // tempty parity to wait for, per accumulator buffer
      int j = 0;
      for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++j) {
        const int buf = j & 1;
        mbar_wait(&tempty_bar[buf], buf ? tph1 : tph0);
        if (buf) tph1 ^= 1u; else tph0 ^= 1u;
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
        for (int it = 0; it < n_steps; ++it) {
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tacc, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k),
                      (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k), idesc, (it | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[st]);
          a_lo += stage16;
          if (++st == S) {
            st = 0;
            ph ^= 1u;
            a_lo = a_lo0;
          }
        }
        tc_commit(&tfull_bar[buf]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..9) =================
    // Two warps per TMEM lane quarter, each draining half of the tile's columns: four epilogue warps are one in-order
    // warp per scheduler, and every tcgen05.wait::ld / conversion chain is then paid in full. The TMEM load of the
    // next 32 columns is in flight while the current 32 are converted and staged.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;           // 0..255
    const int chalf = (warp - 2) >> 2;
    const int subw = BN < 64 ? BN : 64;        // columns per staged sub-tile = one swizzle span
    const int n_sub = BN / subw;
    const int row_b = subw * 2;                // 64 or 128 bytes
    const int sub_bytes = kTileM * row_b;
    const int buf_bytes = n_sub * sub_bytes;
    // chunk position of chunk c of row r: 128-byte rows XOR (r & 7), 64-byte rows XOR ((r >> 1) & 3)
    const uint32_t swz = row_b == 128 ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
    // columns of this warp: half of the tile (32-column tiles are drained by the first warp of the quarter alone)
    const int cb = BN >= 64 ? chalf * (BN >> 1) : 0;
    const int ce = BN >= 64 ? cb + (BN >> 1) : (chalf == 0 ? BN : 0);
    // Residual operand (kRes): the residual TILE is fetched by TMA into the staging buffer the item will use, one
    // item ahead, in the same swizzled layout the output tile is staged in; the epilogue adds the accumulator to it in
    // place and the tile leaves through the TMA store like every other tile. (Per-thread 16-byte residual loads put
    // 2.2 us of exposed latency on every tile of the K = 256 -> 32 projections of the 64x64 level: 27.8 us with a
    // residual against 17.2 us without; residual == out stays legal - a tile is read before it is written.)
    auto fetch_residual = [&](int it, int jj) {  // called by ONE thread, after the buffer's previous store has been read
      if constexpr (kRes) {
        const int cbase = (it % args.n_ntiles) * BN;
        const int mm0 = (it / args.n_ntiles) * kTileM;
        const bool second = args.split_col > 0 && cbase >= args.split_col;
        const CUtensorMap* rm = second ? &maps.r[1] : &maps.r[0];
        const int co = second ? cbase - args.split_col : cbase;
        const int b = jj & 1;
        uint8_t* dst = stg + (args.stg_bufs > 1 ? b * buf_bytes : 0);
        mbar_expect_tx(&rfull_bar[b], (uint32_t)buf_bytes);
        for (int sub = 0; sub < n_sub; ++sub) tma_load_2d(dst + sub * sub_bytes, rm, &rfull_bar[b], co + sub * subw, mm0);
      }
    };
    if (kRes && et == 0 && (int)blockIdx.x < args.n_items) fetch_residual(blockIdx.x, 0);
    int j = 0;
    for (int item = blockIdx.x; item < args.n_items; item += gridDim.x, ++j) {
      const int n_tile = item % args.n_ntiles;
      const int m0 = (item / args.n_ntiles) * kTileM;
      const int col_base = n_tile * BN;
      const CUtensorMap* omap;
      int col_o;
      if (args.split_col > 0 && col_base >= args.split_col) {
        col_o = col_base - args.split_col;
        omap = &maps.o[1];
      } else {
        col_o = col_base;
        omap = &maps.o[0];
      }
      const int buf = j & 1;
      uint8_t* sbuf = stg + (args.stg_bufs > 1 ? buf * buf_bytes : 0);
      // the TMA store that last read this staging buffer must have finished reading it
      if (et == 0) {
        if constexpr (kRes) {
          // the OTHER buffer receives the next item's residual now: its last store (item j-1) must be through with it
          bulk_wait_read_0();
          const int nitem = item + (int)gridDim.x;
          if (nitem < args.n_items && args.stg_bufs > 1) fetch_residual(nitem, j + 1);
        } else {
          if (args.stg_bufs > 1) bulk_wait_read_1(); else bulk_wait_read_0();
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[buf], (uint32_t)(j >> 1) & 1u);
      if constexpr (kRes) mbar_wait(&rfull_bar[buf], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN);
      auto process = [&](const uint32_t (&raw)[32], int c0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[16];
          const int cl = c0 + h * 16;
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(raw[h * 16 + q]);
          if (args.bias) {
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(args.bias + col_base + cl + q));
              v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
            }
          }
          const int sub = cl / subw;
          uint8_t* rowp = sbuf + sub * sub_bytes + r * row_b;
          const uint32_t c16 = (uint32_t)(cl - sub * subw) >> 3;  // first 16-byte chunk of these 16 columns
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint4* slot = reinterpret_cast<uint4*>(rowp + (((c16 + q) ^ swz) << 4));
            if constexpr (kRes) {
              const uint4 rv = *slot;
              float2 f;
              f = unpack_bf16x2(rv.x); v[8 * q + 0] += f.x; v[8 * q + 1] += f.y;
              f = unpack_bf16x2(rv.y); v[8 * q + 2] += f.x; v[8 * q + 3] += f.y;
              f = unpack_bf16x2(rv.z); v[8 * q + 4] += f.x; v[8 * q + 5] += f.y;
              f = unpack_bf16x2(rv.w); v[8 * q + 6] += f.x; v[8 * q + 7] += f.y;
            }
            uint4 u;
            u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
            u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
            u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
            u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
            *slot = u;
          }
        }
      };
      if (cb < ce) {
        uint32_t raw_a[32], raw_b[32];
        tmem_ld_32x32(taddr + (uint32_t)cb, raw_a);
        for (int c0 = cb; c0 < ce; c0 += 64) {
          tmem_ld_wait();
          const bool two = c0 + 32 < ce;
          if (two) tmem_ld_32x32(taddr + (uint32_t)(c0 + 32), raw_b);
          process(raw_a, c0);
          if (two) {
            tmem_ld_wait();
            if (c0 + 64 < ce) tmem_ld_32x32(taddr + (uint32_t)(c0 + 64), raw_a);
            process(raw_b, c0 + 32);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);  // accumulator buffer is free for item j+2
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) {
        for (int sub = 0; sub < n_sub; ++sub) tma_store_2d(omap, sbuf + sub * sub_bytes, col_o + sub * subw, m0);
        bulk_commit();
        if constexpr (kRes) {
          // single staging buffer: the next item's residual goes into this buffer once the store has read it
          const int nitem = item + (int)gridDim.x;
          if (args.stg_bufs == 1 && nitem < args.n_items) {
            bulk_wait_read_0();
            fetch_residual(nitem, j + 1);
          }
        }
      }
    }
    if (et == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)args.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------
// Split-K variant for the small-M levels (the 8x8 level of config_v2_2: M = 2560 = 20 row tiles against 148 SMs).
// With one CTA per output tile those launches either leave half of the SMs idle or cut N into 64-column tiles,
// which re-reads every activation tile N/64 times from L2 (the K loop then runs at the L2 -> SM rate of the chip,
// tools/probe_smallm.py) and still leaves each CTA a 36-step serial K loop plus a full epilogue.
// Here an output tile of 128 x BN (BN = 128 where N allows) belongs to a thread-block CLUSTER of `splits` CTAs, each
// of which runs 1/splits of the K steps. The partial accumulators meet in an L2-resident fp32 workspace (plain
// coalescable stores; DSMEM moves ~20 B/clk per SM and is far slower than L2 for 64 KB tiles), one cluster barrier
// (release / acquire covers the global stores; the hardware co-schedules a cluster, so waiting is deadlock-free
// whatever else shares the GPU), then every CTA finalises its own slice of the tile's ROWS: sum of the partials,
// bias, GroupNorm partial sums, residual, conversion, store - the epilogue is split `splits` ways as well.
// Warps: 0 TMA producer, 1 MMA issuer, 2-9 epilogue (two per TMEM lane quarter).
// ---------------------------------------------------------------------------------------
constexpr int kSplitEpiThreads = 256;
constexpr int kSplitThreads = 64 + kSplitEpiThreads;

template <int BK>
__global__ void __launch_bounds__(kSplitThreads, 1) tapgemm_splitk_kernel(const __grid_constant__ TapMaps maps,
                                                                         const TapArgs args) {
  constexpr int kSwizzle = BK * 2;
  constexpr int kABytes = kTileM * BK * 2;
  constexpr uint32_t kLayout = umma_layout_type(kSwizzle);
  constexpr uint32_t kSBO = 8 * kSwizzle;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_gn[16];      // (sum, sumsq) of up to 8 groups of this N tile
  __shared__ __align__(16) float s_bias[256];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int BN = args.BN;
  const int b_bytes = (BN * BK * 2 + 1023) & ~1023;
  const int stage_bytes = kABytes + b_bytes;
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));

  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int m0 = m_tile * kTileM;
  const int split = blockIdx.z;  // == rank in the (1, 1, splits) cluster
  const int S = args.stages;
  const int n_steps = args.n_taps * args.n_src * args.chunks;
  const int k_begin = (int)(((long)n_steps * split) / args.splits);
  const int k_end = (int)(((long)n_steps * (split + 1)) / args.splits);
  const int col_base = n_tile * BN;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < args.n_maps; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, (uint32_t)args.tmem_cols);
    tmem_relinquish();
  }
  if (threadIdx.x < 16) s_gn[threadIdx.x] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int trace_cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const bool trace_on = args.trace != nullptr && threadIdx.x == 0 && trace_cta < 256;
  if (trace_on) args.trace[256 + 3 * trace_cta] = (long long)globaltimer_ns();
  pdl_wait();
  if (trace_on) args.trace[256 + 3 * trace_cta + 1] = (long long)globaltimer_ns();
  const bool trace0 = args.trace != nullptr && trace_cta == 0;

  float* part = args.ws + ((((size_t)split * gridDim.y + m_tile) * gridDim.x + n_tile) * kTileM) * BN;

  if (warp == 0) {
    // ================= TMA producer: K steps [k_begin, k_end) of the (tap, source, chunk) sequence =================
    if (elect_one()) {
      const int hw = args.H * args.W;
      const int n0 = m0 / hw;
      const int rem = m0 - n0 * hw;
      const int y0 = rem / args.W;
      const int x0 = rem - y0 * args.W;
      const uint32_t tx_bytes = (uint32_t)(kABytes + BN * BK * 2);
      const int per_tap = args.n_src * args.chunks;
      int t = k_begin / per_tap;
      int r2 = k_begin - t * per_tap;
      int sidx = r2 / args.chunks;
      int c = r2 - sidx * args.chunks;
      int kcol = k_begin * BK;
      int st = 0;
      uint32_t ph = 1u;
      for (int step = k_begin; step < k_end; ++step) {
        mbar_wait(&empty_bar[st], ph);
        uint8_t* sa = smem + st * stage_bytes;
        mbar_expect_tx(&full_bar[st], tx_bytes);
        tma_load_4d(sa, &maps.a[args.tap_map[t] + sidx], &full_bar[st], c * BK, x0 + args.tap_dx[t], y0 + args.tap_dy[t], n0);
        tma_load_2d(sa + kABytes, &maps.b, &full_bar[st], kcol, col_base);
        kcol += BK;
        if (++c == args.chunks) {
          c = 0;
          if (++sidx == args.n_src) {
            sidx = 0;
            ++t;
          }
        }
        if (++st == S) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kTileM, BN, 0, 0);
      const uint32_t desc_hi = (kSBO >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (1u << 16);
      int st = 0;
      uint32_t ph = 0u;
      uint32_t a_lo = a_lo0;
      const int n_my = k_end - k_begin;
      for (int it = 0; it < n_my; ++it) {
        mbar_wait(&full_bar[st], ph);
        if (trace0 && it < 64) args.trace[2 * 64 + it] = clock64();
        tc_fence_after();
        const uint32_t b_lo = a_lo + (uint32_t)(kABytes >> 4);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k),
                    (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k), idesc, (it | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[st]);
        a_lo += stage16;
        if (++st == S) {
          st = 0;
          ph ^= 1u;
          a_lo = a_lo0;
        }
      }
      tc_commit(&tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ================= epilogue, part 1: accumulator -> fp32 partial tile in the workspace =================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int chalf = ew >> 2;
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;
    for (int cc = et; cc < BN; cc += kSplitEpiThreads) s_bias[cc] = args.bias ? __ldg(args.bias + col_base + cc) : 0.f;
    const int cb = BN >= 32 ? chalf * (BN >> 1) : 0;
    const int ce = BN >= 32 ? cb + (BN >> 1) : (chalf == 0 ? BN : 0);
    if (trace0 && et == 0) args.trace[3 * 64 + 0] = clock64();
    mbar_wait(&tmem_full_bar, 0);
    if (trace0 && et == 0) args.trace[3 * 64 + 1] = clock64();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float4* prow = reinterpret_cast<float4*>(part + (size_t)r * BN);
    auto dump16 = [&](const uint32_t (&raw)[16], int cl) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        __stcg(prow + (cl >> 2) + j, make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                 __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])));
    };
    if (cb < ce) {
      uint32_t raw_a[16], raw_b[16];
      tmem_ld_32x16(taddr + (uint32_t)cb, raw_a);
      for (int c0 = cb; c0 < ce; c0 += 32) {
        tmem_ld_wait();
        const bool two = c0 + 16 < ce;
        if (two) tmem_ld_32x16(taddr + (uint32_t)(c0 + 16), raw_b);
        dump16(raw_a, c0);
        if (two) {
          tmem_ld_wait();
          if (c0 + 32 < ce) tmem_ld_32x16(taddr + (uint32_t)(c0 + 32), raw_a);
          dump16(raw_b, c0 + 16);
        }
      }
    }
    if (trace0 && et == 0) args.trace[3 * 64 + 4] = clock64();
  }

  // every partial of this tile is in the workspace once all CTAs of the cluster have passed the barrier
  tc_fence_before();
  cluster_sync_all();

  if (warp >= 2) {
    // ================= epilogue, part 2: this CTA's row slice of the tile =================
    const int et = threadIdx.x - 64;
    if (trace0 && et == 0) args.trace[3 * 64 + 5] = clock64();
    const int splits = args.splits;
    const int row_lo = split == 0 ? 0 : ((kTileM * split) / splits) & ~7;
    const int row_hi = split == splits - 1 ? kTileM : ((kTileM * (split + 1)) / splits) & ~7;
    const int chunks_log = 31 - __clz(BN >> 2);  // 16-byte fp32 chunks per row (BN is a power of two)
    const int chunk = et & ((1 << chunks_log) - 1);
    const int rows_per_pass = kSplitEpiThreads >> chunks_log;
    const int row_in_pass = et >> chunks_log;
    uint8_t* outp;
    const uint8_t* resp;
    int ld, col_o;
    if (args.split_col > 0 && col_base >= args.split_col) {
      outp = reinterpret_cast<uint8_t*>(args.out2);
      resp = reinterpret_cast<const uint8_t*>(args.res2);
      ld = args.ld_out2;
      col_o = col_base - args.split_col;
    } else {
      outp = reinterpret_cast<uint8_t*>(args.out);
      resp = reinterpret_cast<const uint8_t*>(args.res);
      ld = args.ld_out;
      col_o = col_base;
    }
    const size_t tile_stride = (size_t)gridDim.y * gridDim.x * kTileM * BN;  // floats between partials of one tile
    const float* tile0 = args.ws + (((size_t)m_tile * gridDim.x + n_tile) * kTileM) * BN + chunk * 4;
    const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[chunk * 4]);
    float s1 = 0.f, s2 = 0.f;
    for (int row = row_lo + row_in_pass; row < row_hi; row += 2 * rows_per_pass) {
      // two rows per iteration: 2 * splits independent 16-byte loads in flight per thread
      const int rowb = row + rows_per_pass;
      const bool has_b = rowb < row_hi;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      float4 pa[4], pb[4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p < splits) {
          pa[p] = __ldcg(reinterpret_cast<const float4*>(tile0 + p * tile_stride + (size_t)row * BN));
          if (has_b) pb[p] = __ldcg(reinterpret_cast<const float4*>(tile0 + p * tile_stride + (size_t)rowb * BN));
        }
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p < splits) {
          va.x += pa[p].x; va.y += pa[p].y; va.z += pa[p].z; va.w += pa[p].w;
          if (has_b) { vb.x += pb[p].x; vb.y += pb[p].y; vb.z += pb[p].z; vb.w += pb[p].w; }
        }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rr = h ? rowb : row;
        if (h && !has_b) break;
        const int m = m0 + rr;
        if (m >= args.M) continue;
        float4 v = h ? vb : va;
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        s1 += (v.x + v.y) + (v.z + v.w);
        s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        const size_t eoff = (size_t)m * ld + col_o + chunk * 4;
        if (args.out_f32) {
          if (resp) {
            const float4 q = *reinterpret_cast<const float4*>(resp + eoff * 4);
            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
          }
          *reinterpret_cast<float4*>(outp + eoff * 4) = v;
        } else {
          if (resp) {
            const uint2 q = *reinterpret_cast<const uint2*>(resp + eoff * 2);
            const float2 q0 = unpack_bf16x2(q.x), q1 = unpack_bf16x2(q.y);
            v.x += q0.x; v.y += q0.y; v.z += q1.x; v.w += q1.y;
          }
          uint2 o;
          o.x = pack_bf16x2(v.x, v.y);
          o.y = pack_bf16x2(v.z, v.w);
          *reinterpret_cast<uint2*>(outp + eoff * 2) = o;
        }
      }
    }
    if (args.gn_sums != nullptr) {
      // GroupNorm partial sums: the thread's four columns lie in one group; lanes of the same group form aligned runs
      const int cpg = args.cpg;
      int run = cpg >> 2;               // lanes per group within a row segment
      if (run > 32) run = 32;
      const int lanes_per_row = 1 << chunks_log;
      if (run > lanes_per_row && cpg < BN) run = lanes_per_row;
      for (int o = 1; o < run; o <<= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if ((lane & (run - 1)) == 0) {
        const int g = (chunk * 4) / cpg;  // group within this N tile
        atomicAdd(&s_gn[2 * g], s1);
        atomicAdd(&s_gn[2 * g + 1], s2);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kSplitEpiThreads) : "memory");
      const int groups_here = (BN + cpg - 1) / cpg;
      if (et < 2 * groups_here) {
        const int sample = min(m0, args.M - 1) / args.rows_per_sample;
        const int rep = m_tile % kGnReplicas;
        float* gdst = args.gn_sums + ((long)(rep * args.n_samples + sample) * args.gn_groups) * 2 + 2 * (col_base / cpg);
        atomicAdd(gdst + et, s_gn[et]);
      }
    }
    if (trace0 && et == 0) args.trace[3 * 64 + 6] = clock64();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)args.tmem_cols);
  }
  if (trace0 && threadIdx.x == 0) args.trace[3 * 64 + 3] = clock64();
  if (trace_on) args.trace[256 + 3 * trace_cta + 2] = (long long)globaltimer_ns();
}

// ---------------------------------------------------------------------------------------
// Reference kernel (CUDA cores, one thread per output element). Test-only.
// ---------------------------------------------------------------------------------------
struct RefArgs {
  int kind, M, N, H, W, n_src, C, n_taps;
  int tap_dy[16], tap_dx[16];
  const bf16* src[2];
  const bf16* wp;
  const float* bias;
  const void* res;
  const void* res2;
  void* out;
  void* out2;
  int split_col, ld_out, ld_out2, out_f32, py, px;
  float* gn_sums;
  int gn_groups, cpg, rows_per_sample;
};

__global__ void tapgemm_ref_kernel(const RefArgs a) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)a.M * a.N) return;
  const int m = (int)(idx / a.N);
  const int n = (int)(idx % a.N);
  const int hw = a.H * a.W;
  const int img = m / hw;
  const int rem = m % hw;
  const int y = rem / a.W, x = rem % a.W;
  const int ktot = a.n_taps * a.n_src * a.C;
  float acc = 0.f;
  for (int t = 0; t < a.n_taps; ++t) {
    int sy, sx, SH, SW;
    if (a.kind == VDN_TAP_DOWN) {
      SH = 2 * a.H; SW = 2 * a.W;
      sy = 2 * y + a.tap_dy[t] - 1;
      sx = 2 * x + a.tap_dx[t] - 1;
    } else {
      SH = a.H; SW = a.W;
      sy = y + a.tap_dy[t];
      sx = x + a.tap_dx[t];
    }
    if (sy < 0 || sy >= SH || sx < 0 || sx >= SW) continue;
    for (int s = 0; s < a.n_src; ++s) {
      const bf16* ap = a.src[s] + (((long)img * SH + sy) * SW + sx) * a.C;
      const bf16* wp = a.wp + (long)n * ktot + (long)(t * a.n_src + s) * a.C;
      for (int c = 0; c < a.C; ++c) acc += __bfloat162float(ap[c]) * __bfloat162float(wp[c]);
    }
  }
  if (a.bias) acc += a.bias[n];
  long orow = m;
  if (a.kind == VDN_TAP_UP) orow = ((long)img * (2 * a.H) + (2 * y + a.py)) * (2 * a.W) + (2 * x + a.px);
  void* outp = a.out;
  const void* resp = a.res;
  int ld = a.ld_out, col = n;
  if (a.split_col > 0 && n >= a.split_col) {
    outp = a.out2; resp = a.res2; ld = a.ld_out2; col = n - a.split_col;
  }
  const long eoff = orow * ld + col;
  if (resp) {
    acc += a.out_f32 ? reinterpret_cast<const float*>(resp)[eoff]
                     : __bfloat162float(reinterpret_cast<const bf16*>(resp)[eoff]);
  }
  if (a.gn_sums) {
    float* p = a.gn_sums + ((long)(m / a.rows_per_sample) * a.gn_groups + n / a.cpg) * 2;
    atomicAdd(p, acc);
    atomicAdd(p + 1, acc * acc);
  }
  if (a.out_f32) reinterpret_cast<float*>(outp)[eoff] = acc;
  else reinterpret_cast<bf16*>(outp)[eoff] = __float2bfloat16(acc);
}

// ---------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------
struct PackArgs {
  int taps, cin, cout, mode, ld, n_off, k_off;
  int perm[16];
};
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, const PackArgs a) {
  const long total = (long)a.taps * a.cin * a.cout;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    // iterate in DESTINATION order so that writes coalesce: (n, t, k)
    int kin = a.mode == 0 ? a.cin : a.cout;
    const int k = (int)(i % kin);
    const int t = (int)((i / kin) % a.taps);
    const int n = (int)(i / ((long)kin * a.taps));
    const int ci = a.mode == 0 ? k : n;
    const int co = a.mode == 0 ? n : k;
    const float v = src[((long)a.perm[t] * a.cin + ci) * a.cout + co];
    dst[(long)(a.n_off + n) * a.ld + a.k_off + (long)t * kin + k] = __float2bfloat16(v);
  }
}

}  // namespace vdn

namespace vdn {

// All weight repacks of a model in ONE launch. Work unit = one 32 x 32 (cin x cout) tile of one tap of one
// job; a block locates its job by binary search over the tile prefix sums. The tile goes through shared
// memory so that BOTH the fp32 reads (cout fastest in the reference layout) and the bf16 writes (K fastest in
// the packed operand: cin for the forward operand, cout for the dgrad operand) are coalesced.
__global__ void __launch_bounds__(256) pack_batched_kernel(const vdn_pack_job* __restrict__ jobs, int n_jobs,
                                                           long long total_tiles) {
  __shared__ float tile[32][33];
  // the tile prefix sums of all jobs live in shared memory: the per-tile binary search used to be ~10 dependent
  // global loads (3 us of latency per 4 KB tile: 73 us for the 10 M parameters of config_v2_2)
  constexpr int kMaxCached = 768;
  __shared__ long long s_begin[kMaxCached];
  const bool cached = n_jobs <= kMaxCached;
  if (cached)
    for (int i = threadIdx.x; i < n_jobs; i += blockDim.x) s_begin[i] = jobs[i].begin;
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  // a block owns a contiguous run of tiles: consecutive tiles mostly belong to the same job
  const long long per = (total_tiles + gridDim.x - 1) / gridDim.x;
  const long long t_begin = (long long)blockIdx.x * per, t_end = min(total_tiles, t_begin + per);
  int cur = -1;
  long long cur_lo = 0, cur_hi = 0;
  for (long long tid = t_begin; tid < t_end; ++tid) {
    if (tid < cur_lo || tid >= cur_hi) {
      int lo = 0, hi = n_jobs - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const long long b = cached ? s_begin[mid] : jobs[mid].begin;
        if (b <= tid) lo = mid; else hi = mid - 1;
      }
      cur = lo;
      cur_lo = cached ? s_begin[lo] : jobs[lo].begin;
      cur_hi = lo + 1 < n_jobs ? (cached ? s_begin[lo + 1] : jobs[lo + 1].begin) : total_tiles;
    }
    const vdn_pack_job& a = jobs[cur];
    const int tiles_ci = (a.cin + 31) >> 5, tiles_co = (a.cout + 31) >> 5;
    int r = (int)(tid - cur_lo);
    const int co0 = (r % tiles_co) * 32;
    r /= tiles_co;
    const int ci0 = (r % tiles_ci) * 32;
    const int t = r / tiles_ci;
    const float* src = a.src + (long)a.perm[t] * a.cin * a.cout;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ci = ci0 + ty + 8 * q, co = co0 + tx;
      tile[ty + 8 * q][tx] = (ci < a.cin && co < a.cout) ? src[(long)ci * a.cout + co] : 0.f;
    }
    __syncthreads();
    bf16* dst = reinterpret_cast<bf16*>(a.dst);
    if (a.mode == 0) {  // dst[(n_off + co)*ld + k_off + t*cin + ci]: ci fastest
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int co = co0 + ty + 8 * q, ci = ci0 + tx;
        if (ci < a.cin && co < a.cout)
          dst[(long)(a.n_off + co) * a.ld + a.k_off + (long)t * a.cin + ci] = __float2bfloat16(tile[tx][ty + 8 * q]);
      }
    } else {            // dst[(n_off + ci)*ld + k_off + t*cout + co]: co fastest
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ci = ci0 + ty + 8 * q, co = co0 + tx;
        if (ci < a.cin && co < a.cout)
          dst[(long)(a.n_off + ci) * a.ld + a.k_off + (long)t * a.cout + co] = __float2bfloat16(tile[ty + 8 * q][tx]);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static int validate_desc(const vdn_tapgemm_desc* d) {
  VDN_REQUIRE(d != nullptr, VDN_E_SHAPE, "tapgemm: null desc");
  VDN_REQUIRE(d->kind >= 0 && d->kind <= 2, VDN_E_SHAPE, "tapgemm: bad kind %d", d->kind);
  VDN_REQUIRE(d->n_img > 0 && d->H > 0 && d->W > 0, VDN_E_SHAPE, "tapgemm: empty grid");
  VDN_REQUIRE(is_pow2(d->W) || d->W % 128 == 0, VDN_E_SHAPE, "tapgemm: W=%d must be a power of two or a multiple of 128", d->W);
  VDN_REQUIRE(is_pow2(d->H) || d->W >= 128, VDN_E_SHAPE, "tapgemm: H=%d must be a power of two when W < 128", d->H);
  VDN_REQUIRE(d->n_src == 1 || d->n_src == 2, VDN_E_SHAPE, "tapgemm: n_src must be 1 or 2");
  VDN_REQUIRE(d->src_c >= 16 && d->src_c % 16 == 0, VDN_E_SHAPE, "tapgemm: src_c=%d must be a multiple of 16", d->src_c);
  VDN_REQUIRE(d->n_taps >= 1 && d->n_taps <= 16, VDN_E_SHAPE, "tapgemm: n_taps=%d out of range", d->n_taps);
  VDN_REQUIRE(d->n_out >= 16 && d->n_out % 16 == 0, VDN_E_SHAPE, "tapgemm: n_out=%d must be a multiple of 16", d->n_out);
  VDN_REQUIRE(d->kind != VDN_TAP_DOWN || d->n_src == 1, VDN_E_SHAPE, "tapgemm: DOWN supports one source");
  if (d->gn_groups > 0) {
    VDN_REQUIRE(d->n_out % d->gn_groups == 0, VDN_E_SHAPE, "tapgemm: n_out %% gn_groups != 0");
    const int cpg = d->n_out / d->gn_groups;
    VDN_REQUIRE(cpg >= 2 && is_pow2(cpg < 16 ? cpg : 16) && (cpg < 16 || cpg % 16 == 0), VDN_E_SHAPE,
                "tapgemm: channels per group %d unsupported", cpg);
    VDN_REQUIRE(d->rows_per_sample > 0 && d->rows_per_sample % 32 == 0, VDN_E_SHAPE,
                "tapgemm: rows_per_sample=%d must be a multiple of 32", d->rows_per_sample);
    VDN_REQUIRE(d->kind != VDN_TAP_UP && d->split_col == 0, VDN_E_SHAPE, "tapgemm: gn stats need a plain output");
  }
  return VDN_OK;
}

// N tile: as wide as possible (fewer re-reads of the A tile) while still giving every SM a CTA - but not below
// 64 columns when N allows it: 32-column tiles re-read A eight times at the 8x8 level and measured slower.
// A launch that reaches at least half of the SMs with 128-column (or narrower) tiles stops there: it runs the
// 8-epilogue-warp instance with one CTA per SM, and halving the tile again to get past 148 CTAs doubles the L2 reads of
// the activations for a second, partial wave (16x16 level, 128 -> 128: 80 CTAs x BN 128 = 9.4 us, 160 x BN 64 = 11.1 us).
static int pick_bn(int N, int m_tiles, int n_steps) {
  int best = -1;
  for (int bn = 256; bn >= (N >= 64 && N % 64 == 0 ? 64 : 32); bn >>= 1) {
    if (bn > N || N % bn != 0) continue;
    if (best < 0) best = bn;
    const int ctas = m_tiles * (N / bn);
    if (ctas >= num_sms()) return bn;
    if (n_steps >= 8 && bn <= 128 && bn >= 64 && 2 * ctas >= num_sms()) return bn;
    best = bn;
  }
  if (best > 0) return best;
  if (N <= 256) return N;
  return 16;
}

template <int BK, int EW>
static int launch_tapgemm(const TapMaps& maps, const TapArgs& args, int smem_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_kernel<BK, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid(args.N / args.BN, ceil_div(args.M, kTileM));
  cudaError_t le = launch_pdl(tapgemm_kernel<BK, EW>, grid, dim3(64 + 32 * EW), (size_t)smem_bytes, st, 1, maps, args);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "tapgemm launch: %s", cudaGetErrorString(le));
  return check_launch("tapgemm_kernel");
}

constexpr int kPersistSmemMax = 226 * 1024;  // of the SM's 227 KB per block; the kernel's static shared memory is < 256 B

template <int BK, bool kRes>
static int launch_tapgemm_persist(const TapMaps& maps, const TapArgs& args, int smem_bytes, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_persist_kernel<BK, kRes>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmemMax);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  cudaError_t le = launch_pdl(tapgemm_persist_kernel<BK, kRes>, dim3(grid), dim3(kPersistThreads), (size_t)smem_bytes, st, 1, maps, args);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "tapgemm_persist launch: %s", cudaGetErrorString(le));
  return check_launch("tapgemm_persist_kernel");
}

template <int BK>
static int launch_tapgemm_splitk(const TapMaps& maps, const TapArgs& args, int smem_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_splitk_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(args.N / args.BN, ceil_div(args.M, kTileM), args.splits);
  cfg.blockDim = dim3(kSplitThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  at[n].id = cudaLaunchAttributeClusterDimension;
  at[n].val.clusterDim.x = 1;
  at[n].val.clusterDim.y = 1;
  at[n].val.clusterDim.z = (unsigned)args.splits;
  ++n;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  cudaError_t le = cudaLaunchKernelEx(&cfg, tapgemm_splitk_kernel<BK>, maps, args);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "tapgemm_splitk launch: %s", cudaGetErrorString(le));
  return check_launch("tapgemm_splitk_kernel");
}

// Split-K plan of a launch (tapgemm_splitk_kernel): N tile width and number of K ranges, or false when the launch has
// enough output tiles to fill the SMs on its own / too few K steps to split / an epilogue the kernel does not have.
struct SplitPlan {
  int bn, splits;
};
static bool splitk_plan(const vdn_tapgemm_desc* d, SplitPlan* plan) {
  // Opt-in (tests / tools set VDN_SPLITK): measured slower than the one-tile-per-CTA kernel at config_v2_2's 8x8 level -
  // the K loop shrinks from 7500 to 3100 cycles, but the partial tiles' trip through L2 (64 KB out, cluster barrier,
  // 60 KB back in) costs 12000 cycles against the 3500-cycle epilogue it replaces (tools/probe_smallm.py).
  if (!tune_on("VDN_SPLITK")) return false;
  if (d->kind == VDN_TAP_UP) return false;
  const int C = d->src_c;
  const int BK = (C % 64 == 0) ? 64 : (C % 32 == 0) ? 32 : 16;
  const int n_steps = d->n_taps * d->n_src * (C / BK);
  const int M = d->n_img * d->H * d->W;
  const int m_tiles = ceil_div(M, kTileM);
  const int N = d->n_out;
  if (d->gn_groups > 0) {
    const int cpg = N / d->gn_groups;
    if (cpg < 4 || (cpg & (cpg - 1)) != 0 || d->rows_per_sample % kTileM != 0) return false;
  }
  const int max_splits = std::min(4, tune_int("VDN_SPLITK_S", 4));
  double best = 1e30;
  for (int bn = 128; bn >= 32; bn >>= 1) {
    if (tune_is_set("VDN_SPLITK_BN") && bn != tune_int("VDN_SPLITK_BN", bn)) continue;
    if (N % bn != 0) continue;
    if (d->split_col > 0 && d->split_col % bn != 0) continue;
    if (d->gn_groups > 0 && bn / (N / d->gn_groups) > 8) continue;
    const int tiles = m_tiles * (N / bn);
    const int splits = std::min(max_splits, num_sms() / std::max(1, tiles));
    if (splits < 2 || n_steps / splits < 3) continue;
    // cycles: K loop = max(tensor pipe, L2 -> SM operand traffic of the whole launch at ~10 KB/clk), + reduction
    const double mma = (bn == 128 ? 67.0 : bn == 64 ? 52.0 : 42.5) * (BK / 16) * ceil_div(n_steps, splits);
    const double traffic = (double)tiles * n_steps * (kTileM + bn) * BK * 2 / 10000.0;
    const double cost = std::max(mma, traffic) + 8.0 * bn + 600.0;
    if (cost < best) {
      best = cost;
      plan->bn = bn;
      plan->splits = splits;
    }
  }
  return best < 1e30;
}

static long long* g_tg_trace = nullptr;

static int tg_env_int(const char* name, int dflt) {
  return tune_int(name, dflt);
}

}  // namespace vdn

using namespace vdn;

extern "C" size_t vdn_tapgemm_workspace(const vdn_tapgemm_desc* d) {
  SplitPlan plan;
  if (!d || validate_desc(d) != VDN_OK || !splitk_plan(d, &plan)) return 0;
  const int m_tiles = ceil_div(d->n_img * d->H * d->W, kTileM);
  return (size_t)plan.splits * m_tiles * kTileM * d->n_out * sizeof(float);
}

extern "C" int vdn_tapgemm(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp,
                           const float* bias, const void* residual, const void* residual2, void* out, void* out2,
                           float* gn_sums, void* stream) {
  return vdn_tapgemm_ws(d, src0, src1, wp, bias, residual, residual2, out, out2, gn_sums, nullptr, 0, stream);
}

extern "C" int vdn_tapgemm_ws(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp,
                              const float* bias, const void* residual, const void* residual2, void* out, void* out2,
                              float* gn_sums, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  VDN_REQUIRE(src0 && wp && out, VDN_E_SHAPE, "tapgemm: null operand");
  VDN_REQUIRE(d->n_src == 1 || src1, VDN_E_SHAPE, "tapgemm: src1 missing");
  VDN_REQUIRE(d->split_col == 0 || out2, VDN_E_SHAPE, "tapgemm: out2 missing");
  VDN_REQUIRE(!(gn_sums && residual), VDN_E_SHAPE, "tapgemm: gn_sums and residual are mutually exclusive");
  if (rowconv_applicable(d, residual, gn_sums))
    return rowconv_launch(d, src0, src1, wp, bias, residual, residual2, out, out2, gn_sums,
                          reinterpret_cast<cudaStream_t>(stream));
  if (slabconv_applicable(d, residual, gn_sums))
    return slabconv_launch(d, src0, src1, wp, bias, residual, residual2, out, out2, gn_sums,
                           reinterpret_cast<cudaStream_t>(stream));

  const int C = d->src_c;
  const int BK = (C % 64 == 0) ? 64 : (C % 32 == 0) ? 32 : 16;
  TapArgs a;
  memset(&a, 0, sizeof(a));
  a.M = d->n_img * d->H * d->W;
  a.N = d->n_out;
  a.BN = pick_bn(d->n_out, ceil_div(a.M, kTileM), d->n_taps * d->n_src * (C / BK));
  {
    // Short-K GEMMs (e.g. the 32 -> 256 / 768 projections) are all epilogue: a 256-column tile holds 256 of the
    // SM's 512 TMEM columns, so only two CTAs are resident and nothing hides their load -> MMA -> store chain.
    // Narrower tiles cost a few extra reads of the (tiny) A tile and buy 4-8 resident CTAs.
    const int smallk_bn = tune_int("VDN_BN_SMALLK", 128);
    const int ktot = d->n_taps * d->n_src * d->src_c;
    if (ktot <= 64 && smallk_bn >= 32 && a.BN > smallk_bn && d->n_out % smallk_bn == 0) a.BN = smallk_bn;
  }
  if (tune_is_set("VDN_BN")) {  // tuning override (experiments only)
    const int v = tune_int("VDN_BN", 0);
    if (v >= 16 && v <= 256 && d->n_out % v == 0) a.BN = v;
  }
  if (d->split_col > 0) {
    while (d->split_col % a.BN != 0) a.BN /= 2;
    VDN_REQUIRE(a.BN >= 16, VDN_E_SHAPE, "tapgemm: split_col %d not tileable", d->split_col);
  }
  a.H = d->H;
  a.W = d->W;
  a.bw = std::min(d->W, 128);
  a.bh = std::min(d->H, 128 / a.bw);
  a.bn = 128 / (a.bw * a.bh);
  a.n_taps = d->n_taps;
  a.n_src = d->n_src;
  a.chunks = C / BK;
  a.bias = bias;
  a.res = residual;
  a.res2 = residual2;
  a.out = out;
  a.out2 = out2;
  a.split_col = d->split_col;
  a.ld_out = d->split_col > 0 ? d->split_col : d->n_out;
  a.ld_out2 = d->n_out - d->split_col;
  a.out_f32 = d->out_dtype == VDN_F32;
  a.scatter = d->kind == VDN_TAP_UP;
  a.py = d->py;
  a.px = d->px;
  a.gn_sums = gn_sums;
  a.gn_groups = gn_sums ? d->gn_groups : 0;
  a.cpg = d->gn_groups > 0 ? d->n_out / d->gn_groups : 1;
  a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  a.n_samples = std::max(1, a.M / a.rows_per_sample);
  if (!gn_sums) a.gn_sums = nullptr;
  a.trace = g_tg_trace;
  VDN_REQUIRE(!(gn_sums && residual), VDN_E_SHAPE, "tapgemm: gn_sums and residual are mutually exclusive");

  TapMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int swz = BK * 2;
  const uint32_t box[4] = {(uint32_t)BK, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bn};
  if (d->kind == VDN_TAP_DOWN) {
    // four parity views of the (n_img, 2H, 2W, C) source; view (ry,rx)[n][y][x] = src[n][2y+ry][2x+rx]
    const uint64_t SW = 2 * (uint64_t)d->W, SH = 2 * (uint64_t)d->H;
    for (int ry = 0; ry < 2; ++ry)
      for (int rx = 0; rx < 2; ++rx) {
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->n_img};
        const uint64_t str[3] = {2 * (uint64_t)C * 2, 2 * SW * C * 2, SH * SW * C * 2};
        const uint8_t* base = reinterpret_cast<const uint8_t*>(src0) + ((uint64_t)ry * SW + rx) * C * 2;
        rc = encode_tmap_bf16(&maps.a[ry * 2 + rx], base, 4, dims, str, box, swz);
        if (rc) return rc;
      }
    for (int t = 0; t < d->n_taps; ++t) {
      const int ky = d->tap_dy[t], kx = d->tap_dx[t];
      VDN_REQUIRE(ky >= 0 && ky < 4 && kx >= 0 && kx < 4, VDN_E_SHAPE, "tapgemm: DOWN tap out of range");
      const int ry = (ky + 1) & 1, rx = (kx + 1) & 1;
      a.tap_map[t] = (signed char)(ry * 2 + rx);
      a.tap_dy[t] = (signed char)((ky - 1) >> 1);
      a.tap_dx[t] = (signed char)((kx - 1) >> 1);
    }
  } else {
    const void* srcs[2] = {src0, src1};
    for (int s = 0; s < d->n_src; ++s) {
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->n_img};
      const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)d->W * C * 2, (uint64_t)d->H * d->W * C * 2};
      rc = encode_tmap_bf16(&maps.a[s], srcs[s], 4, dims, str, box, swz);
      if (rc) return rc;
    }
    for (int t = 0; t < d->n_taps; ++t) {
      VDN_REQUIRE(d->tap_dy[t] >= -64 && d->tap_dy[t] <= 64 && d->tap_dx[t] >= -64 && d->tap_dx[t] <= 64,
                  VDN_E_SHAPE, "tapgemm: tap shift out of range");
      a.tap_map[t] = 0;
      a.tap_dy[t] = (signed char)d->tap_dy[t];
      a.tap_dx[t] = (signed char)d->tap_dx[t];
    }
  }
  a.n_maps = d->kind == VDN_TAP_DOWN ? 4 : d->n_src;
  {
    const uint64_t ktot = (uint64_t)d->n_taps * d->n_src * C;
    const uint64_t dims[2] = {ktot, (uint64_t)d->n_out};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t bbox[2] = {(uint32_t)BK, (uint32_t)a.BN};
    rc = encode_tmap_bf16(&maps.b, wp, 2, dims, str, bbox, swz);
    if (rc) return rc;
  }

  const int a_bytes = kTileM * BK * 2;
  const int b_bytes = (a.BN * BK * 2 + 1023) & ~1023;
  const int stage_bytes = a_bytes + b_bytes;
  // Few CTAs per SM (small-M layers): the K loop is a latency chain, so pipeline as deep as shared
  // memory allows. Many CTAs per SM (large-M layers): keep shared memory low so that they co-reside.
  const int n_steps = d->n_taps * d->n_src * a.chunks;
  const int n_ctas = ceil_div(a.M, kTileM) * (d->n_out / a.BN);
  const int ctas_per_sm = std::min(6, ceil_div(n_ctas, num_sms()));
  const int budget = (200 * 1024) / ctas_per_sm;
  a.stages = std::max(2, std::min(std::min(kMaxStages, budget / stage_bytes), std::max(n_steps, 2)));
  if (tune_is_set("VDN_STAGES")) a.stages = std::max(1, std::min(kMaxStages, tune_int("VDN_STAGES", a.stages)));
  int cols = 32;
  while (cols < a.BN) cols *= 2;
  a.tmem_cols = cols;
  const int tile_bytes = d->kind == VDN_TAP_UP ? 0 : kTileM * (a.BN * (a.out_f32 ? 4 : 2) + 16);
  const int smem_bytes = std::max(a.stages * stage_bytes, tile_bytes) + 1024;

  // alignment of epilogue vector accesses
  VDN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!residual || (reinterpret_cast<uintptr_t>(residual) & 15) == 0) &&
                  (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
              VDN_E_ALIGN, "tapgemm: out/residual/bias must be 16B aligned");

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  // ---- split-K variant: small-M launches, when the caller provides the scratch (see tapgemm_splitk_kernel) ----
  {
    SplitPlan plan;
    if (workspace && splitk_plan(d, &plan)) {
      const size_t need = (size_t)plan.splits * ceil_div(a.M, kTileM) * kTileM * d->n_out * sizeof(float);
      VDN_REQUIRE(workspace_bytes >= need, VDN_E_SHAPE, "tapgemm: workspace of %zu bytes, %zu needed", workspace_bytes, need);
      VDN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, VDN_E_ALIGN, "tapgemm: workspace must be 16B aligned");
      TapArgs p = a;
      p.BN = plan.bn;
      p.splits = plan.splits;
      p.ws = reinterpret_cast<float*>(workspace);
      p.tmem_cols = std::max(32, plan.bn);
      const uint64_t ktot = (uint64_t)d->n_taps * d->n_src * C;
      const uint64_t dims[2] = {ktot, (uint64_t)d->n_out};
      const uint64_t str[1] = {ktot * 2};
      const uint32_t bbox[2] = {(uint32_t)BK, (uint32_t)plan.bn};
      rc = encode_tmap_bf16(&maps.b, wp, 2, dims, str, bbox, swz);
      if (rc) return rc;
      const int sb = a_bytes + ((plan.bn * BK * 2 + 1023) & ~1023);
      p.stages = std::max(2, std::min(std::min(kMaxStages, (200 * 1024) / sb), ceil_div(n_steps, plan.splits)));
      const int ssmem = p.stages * sb + 1024;
      if (BK == 64) return launch_tapgemm_splitk<64>(maps, p, ssmem, st);
      if (BK == 32) return launch_tapgemm_splitk<32>(maps, p, ssmem, st);
      return launch_tapgemm_splitk<16>(maps, p, ssmem, st);
    }
  }

  // ---- persistent variant: many tiles per SM and a short K loop (see tapgemm_persist_kernel) ----
  {
    const int m_tiles = ceil_div(a.M, kTileM);
    const long items = (long)m_tiles * (d->n_out / a.BN);
    // From one item per SM on (measured on the training step: 444 -> 148 items = 6.32 -> 6.25 ms). Narrow tiles with a
    // residual used to stay on the one-tile kernel (per-thread residual loads: 27.6 vs 23.6 us at K = 256 -> 32, 163840
    // pixels); with the residual tile arriving by TMA the persistent kernel does it in 20.3 us. VDN_PERSIST_NARROW_RES=0
    // restores the old routing.
    const int persist_min_items = tg_env_int("VDN_PERSIST_MIN_ITEMS", num_sms());
    const bool narrow_res = a.BN < 64 && (residual || residual2) && persist_min_items > 1 &&
                            tg_env_int("VDN_PERSIST_NARROW_RES", 1) == 0;
    const bool shape_ok = !gn_sums && !a.out_f32 && !a.scatter && a.BN >= 32 && (a.BN & (a.BN - 1)) == 0 && !narrow_res &&
                          n_steps <= tg_env_int("VDN_PERSIST_MAX_STEPS", 12) &&
                          items >= (long)persist_min_items;
    const bool persist_off = tune_on("VDN_NO_PERSIST");
    if (shape_ok && !persist_off) {
      TapArgs p = a;
      p.n_ntiles = d->n_out / a.BN;
      p.n_items = (int)items;
      p.tmem_cols = 32;
      while (p.tmem_cols < 2 * a.BN) p.tmem_cols *= 2;
      const int subw = std::min(a.BN, 64);
      const int buf_bytes = kTileM * a.BN * 2;
      // staging: two buffers unless that leaves fewer than two K stages per accumulator in flight
      p.stg_bufs = (1024 + 2 * buf_bytes + std::min(2 * n_steps, 4) * stage_bytes <= kPersistSmemMax) ? 2 : 1;
      if (p.stg_bufs == 1 && 1024 + 2 * buf_bytes + 2 * stage_bytes <= kPersistSmemMax) p.stg_bufs = 2;
      int S = std::min(kMaxStages, (kPersistSmemMax - 1024 - p.stg_bufs * buf_bytes) / stage_bytes);
      S = std::min(S, std::max(2, 3 * n_steps));
      if (S >= 2) {
        p.stages = S;
        const uint64_t Mrows = (uint64_t)a.M;
        const uint32_t obox[2] = {(uint32_t)subw, (uint32_t)kTileM};
        const uint64_t dims0[2] = {(uint64_t)a.ld_out, Mrows};
        const uint64_t str0[1] = {(uint64_t)a.ld_out * 2};
        rc = encode_tmap_bf16(&maps.o[0], out, 2, dims0, str0, obox, subw * 2);
        if (rc) return rc;
        if (residual) {
          rc = encode_tmap_bf16(&maps.r[0], residual, 2, dims0, str0, obox, subw * 2);
          if (rc) return rc;
        }
        if (d->split_col > 0) {
          const uint64_t dims1[2] = {(uint64_t)a.ld_out2, Mrows};
          const uint64_t str1[1] = {(uint64_t)a.ld_out2 * 2};
          rc = encode_tmap_bf16(&maps.o[1], out2, 2, dims1, str1, obox, subw * 2);
          if (rc) return rc;
          if (residual2) {
            rc = encode_tmap_bf16(&maps.r[1], residual2, 2, dims1, str1, obox, subw * 2);
            if (rc) return rc;
          }
        }
        const int psmem = 1024 + S * stage_bytes + p.stg_bufs * buf_bytes;
        const int cps = std::max(1, std::min(512 / p.tmem_cols, (227 * 1024) / (psmem + 1024)));
        int grid = (int)std::min<long>(items, (long)num_sms() * cps);
        if (tune_is_set("VDN_PERSIST_GRID")) grid = std::max(1, std::min((int)items, tune_int("VDN_PERSIST_GRID", grid)));
        const bool has_res = residual != nullptr || residual2 != nullptr;
        VDN_REQUIRE(!has_res || (residual != nullptr && (d->split_col == 0 || residual2 != nullptr)), VDN_E_SHAPE,
                    "tapgemm: a split output needs both residuals or none");
        VDN_REQUIRE(!has_res || ((reinterpret_cast<uintptr_t>(residual) & 15) == 0 &&
                                 (!residual2 || (reinterpret_cast<uintptr_t>(residual2) & 15) == 0)),
                    VDN_E_ALIGN, "tapgemm: residual must be 16B aligned");
        if (BK == 64) return has_res ? launch_tapgemm_persist<64, true>(maps, p, psmem, grid, st) : launch_tapgemm_persist<64, false>(maps, p, psmem, grid, st);
        if (BK == 32) return has_res ? launch_tapgemm_persist<32, true>(maps, p, psmem, grid, st) : launch_tapgemm_persist<32, false>(maps, p, psmem, grid, st);
        return has_res ? launch_tapgemm_persist<16, true>(maps, p, psmem, grid, st) : launch_tapgemm_persist<16, false>(maps, p, psmem, grid, st);
      }
    }
  }

  // launches that cannot fill the SMs anyway run the 8-epilogue-warp instance (one CTA per SM, no register cap)
  bool wide_epi = n_ctas <= num_sms() && a.BN >= 64;
  if (tune_is_set("VDN_EW")) wide_epi = tune_int("VDN_EW", 4) == 8;
  int smem_launch = smem_bytes;
  if (wide_epi && residual && !a.out_f32 && !a.scatter && (d->split_col == 0 || residual2) && !tune_on("VDN_NO_RES_TMA")) {
    // residual tile by TMA into its own shared-memory region behind the pipeline stages / the staged output tile
    const int res_bytes = kTileM * a.BN * 2;
    // the pipeline gives up stages to make room (a one-CTA-per-SM launch had taken all 200 KB for them)
    const int max_stages = (200 * 1024 - 1024 - 128 - res_bytes) / stage_bytes;
    if (max_stages >= 3 && !tune_is_set("VDN_STAGES")) a.stages = std::min(a.stages, max_stages);
    const int off = (std::max(a.stages * stage_bytes, tile_bytes) + 127) & ~127;
    if (off + res_bytes + 1024 <= 200 * 1024) {
      const uint64_t Mrows = (uint64_t)a.M;
      const uint32_t rbox[2] = {(uint32_t)a.BN, (uint32_t)kTileM};
      const uint64_t dims0[2] = {(uint64_t)a.ld_out, Mrows};
      const uint64_t str0[1] = {(uint64_t)a.ld_out * 2};
      rc = encode_tmap_bf16(&maps.r[0], residual, 2, dims0, str0, rbox, 0);
      if (rc) return rc;
      if (d->split_col > 0) {
        const uint64_t dims1[2] = {(uint64_t)a.ld_out2, Mrows};
        const uint64_t str1[1] = {(uint64_t)a.ld_out2 * 2};
        rc = encode_tmap_bf16(&maps.r[1], residual2, 2, dims1, str1, rbox, 0);
        if (rc) return rc;
      }
      a.res_tma = off;
      smem_launch = off + res_bytes + 1024;
    }
  }
  if (wide_epi) {
    if (BK == 64) return launch_tapgemm<64, kEpiWarpsWide>(maps, a, smem_launch, st);
    if (BK == 32) return launch_tapgemm<32, kEpiWarpsWide>(maps, a, smem_launch, st);
    return launch_tapgemm<16, kEpiWarpsWide>(maps, a, smem_launch, st);
  }
  if (BK == 64) return launch_tapgemm<64, 4>(maps, a, smem_bytes, st);
  if (BK == 32) return launch_tapgemm<32, 4>(maps, a, smem_bytes, st);
  return launch_tapgemm<16, 4>(maps, a, smem_bytes, st);
}

extern "C" void vdn_debug_tapgemm_trace(void* dev_buf) { vdn::g_tg_trace = reinterpret_cast<long long*>(dev_buf); }

extern "C" int vdn_tapgemm_ref(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp,
                               const float* bias, const void* residual, const void* residual2, void* out,
                               void* out2, float* gn_sums, void* stream) {
  int rc = validate_desc(d);
  if (rc) return rc;
  RefArgs a;
  memset(&a, 0, sizeof(a));
  a.kind = d->kind;
  a.M = d->n_img * d->H * d->W;
  a.N = d->n_out;
  a.H = d->H;
  a.W = d->W;
  a.n_src = d->n_src;
  a.C = d->src_c;
  a.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t) {
    a.tap_dy[t] = d->tap_dy[t];
    a.tap_dx[t] = d->tap_dx[t];
  }
  a.src[0] = reinterpret_cast<const bf16*>(src0);
  a.src[1] = reinterpret_cast<const bf16*>(src1);
  a.wp = reinterpret_cast<const bf16*>(wp);
  a.bias = bias;
  a.res = residual;
  a.res2 = residual2;
  a.out = out;
  a.out2 = out2;
  a.split_col = d->split_col;
  a.ld_out = d->split_col > 0 ? d->split_col : d->n_out;
  a.ld_out2 = d->n_out - d->split_col;
  a.out_f32 = d->out_dtype == VDN_F32;
  a.py = d->py;
  a.px = d->px;
  a.gn_sums = gn_sums;
  a.gn_groups = d->gn_groups;
  a.cpg = d->gn_groups > 0 ? d->n_out / d->gn_groups : 1;
  a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  const long total = (long)a.M * a.N;
  const int threads = 256;
  tapgemm_ref_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("tapgemm_ref_kernel");
}

extern "C" int vdn_pack_batched(const void* jobs_dev, int n_jobs, long long total, void* stream) {
  VDN_REQUIRE(jobs_dev && n_jobs > 0 && total > 0, VDN_E_SHAPE, "pack_batched: bad args");
  const int blocks = (int)std::min<long long>(total, 148 * 16);
  pack_batched_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const vdn_pack_job*>(jobs_dev), n_jobs, total);
  return check_launch("pack_batched_kernel");
}

extern "C" int vdn_pack_weight(const float* src, void* dst, int taps, int cin, int cout, int mode,
                               const int* perm_host, int ld, int n_off, int k_off, void* stream) {
  VDN_REQUIRE(src && dst, VDN_E_SHAPE, "pack_weight: null pointer");
  VDN_REQUIRE(taps >= 1 && taps <= 16 && cin > 0 && cout > 0 && (mode == 0 || mode == 1), VDN_E_SHAPE,
              "pack_weight: bad arguments");
  PackArgs a;
  a.taps = taps; a.cin = cin; a.cout = cout; a.mode = mode; a.ld = ld; a.n_off = n_off; a.k_off = k_off;
  for (int t = 0; t < 16; ++t) a.perm[t] = (perm_host && t < taps) ? perm_host[t] : t;
  const long total = (long)taps * cin * cout;
  const int threads = 256;
  const int blocks = (int)std::min<long>((total + threads - 1) / threads, 148 * 8);
  pack_weight_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, reinterpret_cast<bf16*>(dst), a);
  return check_launch("pack_weight_kernel");
}
