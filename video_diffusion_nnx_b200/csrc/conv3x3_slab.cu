// Slab (1,3,3) convolution for the wide-channel levels: tcgen05 implicit GEMM with operand reuse in shared memory.
//
// Serves the (1,3,3) convs of the Unet3D path (Block.proj, modules.py:162-165, forward and dgrad) with
// >= 64 input channels per source and enough pixels to fill the GPU. There the generic tap-GEMM is bound by
// L2 -> SM bandwidth, not by the tensor pipe: a 128 x 128 tile pulls 16 KB of pixels and 16 KB of weights
// per K = 64 step (268 MMA cycles) = 122 B/clk per SM against ~62 B/clk measured (836 of 1650 TFLOP/s).
//
// Here a CTA computes a GROUP of R vertically adjacent 128-pixel tiles (R * TR image rows, TR = 128 / W)
// of one image against one BN-column weight tile, with R accumulators in TMEM (R * BN <= 512 columns).
// One pipeline stage = (source, 32-channel chunk, dx):
//   * ONE TMA box of R*TR + 2 image rows (x shifted by dx, out-of-image pixels zero-filled = SAME padding)
//     - the "slab" - from which the A operand of (tile i, dy) is the row-offset view starting at slab row
//     i*TR + dy + 1, so a pixel row is fetched once per dx instead of three times, and
//   * the three weight tiles (dy = -1, 0, 1) of that (dx, source, chunk), each used by all R tiles.
// L2 -> SM traffic per MMA cycle drops 2.7x (R = 4, W = 128: 72 KB per 1608 cycles = 45 B/clk).
// The CTA is persistent over (group, N tile) items; the TMA producer runs ahead into the next item while
// the epilogue (bias / residual / GroupNorm partial sums / bf16 store through a staged tile) drains TMEM.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kSlEpiWarps = 8;  // two epilogue warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int kSlEpiThreads = 32 * kSlEpiWarps;
constexpr int kSlThreads = 64 + kSlEpiThreads;
constexpr int kSlMaxStages = 6;
constexpr int kSlTileM = 128;

struct SlabMaps {
  CUtensorMap a[2];  // per source: box (32, W, R*TR + 2, 1)
  CUtensorMap b;     // packed weights [N][9*n_src*C], box (32, BN)
  CUtensorMap o[2];  // outputs [M][ld] (second: columns >= split_col), box (64, 128), 128-byte swizzle
};

struct SlabArgs {
  int N, BN, R;
  int W, TR, GPI;  // GPI = groups per image
  int n_src, chunks, C;
  int n_ntiles, n_items;
  int S;
  int slab_bytes, b_bytes, stage_bytes;
  int tmem_cols;
  signed char tap_of[9];  // (dy+1)*3 + (dx+1) -> tap index in the packed operand
  const float* bias;
  const bf16* res;
  const bf16* res2;
  bf16* out;
  bf16* out2;
  int split_col, ld_out, ld_out2;
  float* gn_sums;
  int gn_groups, cpg, rows_per_sample, n_samples;
  int dbg;  // experiments only (VDN_SLAB_DBG): 1 skip epilogue work, 2 skip MMAs, 4 skip slab loads, 8 skip weight loads
};

// (sum, sumsq) of 16 consecutive channels of this thread's row, reduced over the warp's 32 rows into the warp's
// private shared-memory slots (groups of CPG16 channels; CPG16 = 16 also serves wider groups).
template <int CPG16>
__device__ __forceinline__ void sl_gn_accumulate16(const float (&v)[16], float* slot, int lane) {
#pragma unroll
  for (int j = 0; j < 16; j += CPG16) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CPG16; ++k) {
      s1 += v[j + k];
      s2 += v[j + k] * v[j + k];
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      slot[2 * (j / CPG16)] += s1;
      slot[2 * (j / CPG16) + 1] += s2;
    }
  }
}

template <int kSlBK>  // channels per stage (2*kSlBK bytes = swizzle span): 32 or 16
__global__ void __launch_bounds__(kSlThreads) conv3x3_slab_kernel(const __grid_constant__ SlabMaps maps,
                                                                  const SlabArgs a) {
  constexpr uint32_t kLayout = umma_layout_type(2 * kSlBK);  // 64- or 32-byte swizzle
  constexpr uint32_t kSBO = 8 * 2 * kSlBK;                   // 8 rows of one swizzle span

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kSlMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kSlMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar;
  __shared__ __align__(8) uint64_t tempty_bar[4];  // per accumulator: drained by the epilogue
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_gn[2][kSlEpiWarps][16];  // [staging parity][epilogue warp][(sum, sumsq) x groups of the N tile]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const int S = a.S, BN = a.BN, R = a.R;
  uint8_t* stg = smem + S * a.stage_bytes;  // staged output tile (epilogue)
  const int n_steps = a.n_src * a.chunks * 3;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.n_src; ++s) tma_prefetch_desc(&maps.a[s]);
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.o[0]);
    if (a.split_col > 0) tma_prefetch_desc(&maps.o[1]);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&tempty_bar[i], kSlEpiThreads);
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
  pdl_wait();  // everything above overlapped the tail of the previous kernel; global memory is touched below

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      const uint32_t tx = (uint32_t)(((a.dbg & 4) ? 0 : a.slab_bytes) + ((a.dbg & 8) ? 0 : 3 * a.b_bytes));
      int st = 0;
      uint32_t ph = 1u;  // parity to wait for on the empty barrier of the slot
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int nt = item % a.n_ntiles, g = item / a.n_ntiles;
        const int n = g / a.GPI;
        const int y0 = (g - n * a.GPI) * R * a.TR;
        for (int s = 0; s < a.n_src; ++s)
          for (int c = 0; c < a.chunks; ++c)
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
              mbar_wait(&empty_bar[st], ph);
              uint8_t* slab = smem + st * a.stage_bytes;
              mbar_expect_tx(&full_bar[st], tx);
              if (!(a.dbg & 4)) tma_load_4d(slab, &maps.a[s], &full_bar[st], c * kSlBK, dxi - 1, y0 - 1, n);
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi) {
                if (a.dbg & 8) break;
                const int t = a.tap_of[dyi * 3 + dxi];
                tma_load_2d(slab + a.slab_bytes + dyi * a.b_bytes, &maps.b, &full_bar[st],
                            (t * a.n_src + s) * a.C + c * kSlBK, nt * BN);
              }
              if (++st == S) {
                st = 0;
                ph ^= 1u;
              }
            }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kSlTileM, BN, 0, 0);
      // descriptor halves: hi = SBO | version | swizzle mode; lo = (address >> 4) | LBO(16 B) << 16
      const uint32_t desc_hi = (kSBO >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t stage16 = (uint32_t)a.stage_bytes >> 4, slab16 = (uint32_t)a.slab_bytes >> 4;
      const uint32_t b16 = (uint32_t)a.b_bytes >> 4;
      const uint32_t row16 = (uint32_t)(a.W * 2 * kSlBK) >> 4;    // one image row of the slab
      const uint32_t tile16 = row16 * (uint32_t)a.TR;              // TR rows = one 128-pixel tile
      const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (1u << 16);
      int st = 0;
      uint32_t ph = 0u, tph = 1u;
      uint32_t a_lo = a_lo0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        for (int it = 0; it < n_steps; ++it) {
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint32_t b_lo = a_lo + slab16;
          for (int i = 0; i < ((a.dbg & 2) ? 0 : R); ++i) {
            const uint32_t tacc = tmem_base + (uint32_t)(i * BN);
            if (it == 0) {  // accumulator i is reused as soon as the epilogue has drained it (not the whole group)
              mbar_wait(&tempty_bar[i], tph);
              tc_fence_after();
            }
            const uint32_t at = a_lo + (uint32_t)i * tile16;
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
              for (int k = 0; k < kSlBK / 16; ++k)
                umma_bf16(tacc, (static_cast<uint64_t>(desc_hi) << 32) | (at + (uint32_t)dyi * row16 + 2u * k),
                          (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + (uint32_t)dyi * b16 + 2u * k), idesc,
                          (it | dyi | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[st]);  // frees the smem slot when these MMAs retire
          a_lo += stage16;
          if (++st == S) {
            st = 0;
            ph ^= 1u;
            a_lo = a_lo0;
          }
        }
        tc_commit(&tfull_bar);
        tph ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..5) =================
    // Per tile: TMEM -> registers -> (+bias, GroupNorm partial sums) -> bf16 -> staging tile in shared memory in
    // the 128-byte-swizzled layout of the output tensor map, then ONE thread hands the tile to the TMA store
    // engine; the warps go straight on to the next accumulator (two staging buffers), so the HBM write is off
    // the epilogue's critical path. With a residual operand the tile is added and stored by the threads instead.
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;  // epilogue thread id 0..255
    const int chalf = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    const bool gn_on = a.gn_sums != nullptr;
    const int cpg = a.cpg;
    const uint32_t swz = (uint32_t)(r & 7);
    const int n_sub = BN >> 6;  // 64-column (128-byte) sub-tiles
    uint32_t tph = 0u;
    int tcount = 0;  // tiles finished by this CTA: its parity selects the staging / s_gn buffer
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int nt = item % a.n_ntiles, g = item / a.n_ntiles;
      const int col_base = nt * BN;
      bf16* outp;
      const bf16* resp;
      const CUtensorMap* omap;
      int ld, col_o;
      if (a.split_col > 0 && col_base >= a.split_col) {
        outp = a.out2; resp = a.res2; ld = a.ld_out2; col_o = col_base - a.split_col; omap = &maps.o[1];
      } else {
        outp = a.out; resp = a.res; ld = a.ld_out; col_o = col_base; omap = &maps.o[0];
      }
      const int g_tile0 = col_base / cpg;  // first GroupNorm group covered by this N tile
      mbar_wait(&tfull_bar, tph);
      tph ^= 1u;
      tc_fence_after();
      if (a.dbg & 1) {
        tc_fence_before();
        for (int i = 0; i < R; ++i) mbar_arrive(&tempty_bar[i]);
        continue;
      }
      for (int i = 0; i < R; ++i, ++tcount) {
        const int tile = g * R + i;
        const long m0 = (long)tile * kSlTileM;
        const int buf = tcount & 1;
        uint8_t* sbuf = stg + buf * (n_sub * 16384);
        if (et == 0) bulk_wait_read_1();  // the TMA store that read this staging buffer two tiles ago is done
        if (et < 16 * kSlEpiWarps) s_gn[buf][et >> 4][et & 15] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * BN);
        // GroupNorm partial sums: groups of >= 16 channels are accumulated per thread over the whole tile
        // (ga[2g] = sum, ga[2g+1] = sum of squares of group g of this N tile) and reduced over the warp ONCE per tile
        // with 16 shuffles; a shuffle reduction per 16-column chunk (80 per tile) sat on the epilogue's serial
        // path and cost 48 % of the kernel (378 vs 255 us at 128 -> 128 @128x128). Narrower groups keep the old path.
        float ga[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) ga[q] = 0.f;
        const bool gn_wide = gn_on && cpg >= 16 && (cpg & (cpg - 1)) == 0 && !(a.dbg & 32);
        const int cpg_log = 31 - __clz(cpg);
        auto process = [&](const uint32_t (&raw)[32], int c0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[16];
            const int cl = c0 + h * 16;    // column inside the tile
            const int cg = col_base + cl;  // global output column of v[0]
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[h * 16 + j]);
            if (a.bias) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + cg + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
            if (gn_wide) {
              // four independent partial chains (a single in-order warp per scheduler: a 16-deep dependent
              // chain costs 16 x the FADD latency)
              float p1[4], p2[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                p1[q] = v[q];
                p2[q] = v[q] * v[q];
              }
#pragma unroll
              for (int j = 4; j < 16; ++j) {
                p1[j & 3] += v[j];
                p2[j & 3] = fmaf(v[j], v[j], p2[j & 3]);
              }
              const float s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]), s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
              const int g = cl >> cpg_log;  // group within this N tile (cpg is a power of two >= 16; warp-uniform)
#pragma unroll
              for (int gi = 0; gi < 8; ++gi)
                if (gi == g) {
                  ga[2 * gi] += s1;
                  ga[2 * gi + 1] += s2;
                }
            } else if (gn_on && !(a.dbg & 32)) {
              float* slot = &s_gn[buf][warp - 2][2 * (cg / cpg - g_tile0)];
              if (cpg == 8) sl_gn_accumulate16<8>(v, slot, lane);
              else if (cpg == 4) sl_gn_accumulate16<4>(v, slot, lane);
              else sl_gn_accumulate16<2>(v, slot, lane);
            }
            // sub-tile of 64 columns; 16-byte chunk j of row r lives at chunk position j ^ (r & 7)
            uint8_t* rowp = sbuf + (cl >> 6) * 16384 + r * 128;
            const uint32_t j0 = (uint32_t)(cl & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4 q;
              q.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              q.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              q.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              q.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(rowp + (((j0 + j) ^ swz) << 4)) = q;
            }
          }
        };
        // the TMEM load of the next 32 columns is in flight while the current 32 are processed
        uint32_t raw_a[32], raw_b[32];
        const int cb = chalf * (BN >> 1), ce = cb + (BN >> 1);  // BN / 2 = 32 or 64 columns per warp
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
        if (gn_wide) {
          const float tot = warp_sum16(ga, lane);
          if ((lane & 1) == 0) s_gn[buf][warp - 2][lane >> 1] = tot;
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[i]);  // accumulator i has been read: the MMA warp may reuse it for the next item
        if (!resp) fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!resp) {
          if (et == 0) {
            for (int sub = 0; sub < n_sub; ++sub) tma_store_2d(omap, sbuf + sub * 16384, col_o + sub * 64, (int)m0);
            bulk_commit();
          }
        } else {
          // residual add: consecutive threads handle consecutive 16-byte segments of a row (coalesced)
          const int spr = BN >> 3;  // 16B segments per row
          const int total = kSlTileM * spr;
          for (int idx = et; idx < total; idx += kSlEpiThreads) {
            const int rr = idx / spr, sg = idx - rr * spr;
            uint4 q = *reinterpret_cast<const uint4*>(sbuf + (sg >> 3) * 16384 + rr * 128 + (((sg & 7) ^ (rr & 7)) << 4));
            const long goff = ((m0 + rr) * ld + col_o) * 2 + sg * 16;
            const uint4 rq = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(resp) + goff);
            float2 x, y;
            x = unpack_bf16x2(q.x); y = unpack_bf16x2(rq.x); q.x = pack_bf16x2(x.x + y.x, x.y + y.y);
            x = unpack_bf16x2(q.y); y = unpack_bf16x2(rq.y); q.y = pack_bf16x2(x.x + y.x, x.y + y.y);
            x = unpack_bf16x2(q.z); y = unpack_bf16x2(rq.z); q.z = pack_bf16x2(x.x + y.x, x.y + y.y);
            x = unpack_bf16x2(q.w); y = unpack_bf16x2(rq.w); q.w = pack_bf16x2(x.x + y.x, x.y + y.y);
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(outp) + goff) = q;
          }
        }
        if (gn_on && et < 2 * (BN / cpg) && !(a.dbg & 16)) {
          const int sample = (int)(m0 / a.rows_per_sample);
          float* gdst = a.gn_sums + ((long)((tile % kGnReplicas) * a.n_samples + sample) * a.gn_groups) * 2;
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < kSlEpiWarps; ++w) tot += s_gn[buf][w][et];
          atomicAdd(gdst + 2 * g_tile0 + et, tot);
        }
      }
    }
    if (et == 0) bulk_wait_all();  // staged tiles must stay in shared memory until the TMA stores have read them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

static int sl_env_int(const char* name, int dflt) {
  return tune_int(name, dflt);
}

// Geometry shared by applicable() and launch(): BN column tile, R tiles per group.
static bool slab_geometry(const vdn_tapgemm_desc* d, int* BN, int* R) {
  if (d->W != 16 && d->W != 32 && d->W != 64 && d->W != 128) return false;
  const int TR = kSlTileM / d->W;
  if (d->H % TR != 0) return false;
  const int TPI = d->H / TR;
  int bn;
  if (d->n_out % 128 == 0) bn = 128;
  else if (d->n_out == 64) bn = 64;
  else return false;
  if (d->split_col != 0 && d->split_col % bn != 0) return false;
  int r = 0;
  for (int cand = 4; cand >= 2; cand >>= 1)
    if (TPI % cand == 0 && cand * bn <= 512) {
      r = cand;
      break;
    }
  if (r == 0) return false;
  *BN = bn;
  *R = r;
  return true;
}

bool slabconv_applicable(const vdn_tapgemm_desc* d, const void* residual, const float* gn_sums) {
  if (tune_on("VDN_NO_SLABCONV")) return false;
  if (d->kind != VDN_TAP_UNIT || d->n_taps != 9 || d->out_dtype != VDN_BF16) return false;
  if (d->src_c < 64 || d->src_c % 32 != 0) return false;
  int BN, R;
  if (!slab_geometry(d, &BN, &R)) return false;
  unsigned seen = 0;  // the taps must be a permutation of the 3x3 neighbourhood
  for (int t = 0; t < 9; ++t) {
    if (d->tap_dy[t] < -1 || d->tap_dy[t] > 1 || d->tap_dx[t] < -1 || d->tap_dx[t] > 1) return false;
    seen |= 1u << ((d->tap_dy[t] + 1) * 3 + d->tap_dx[t] + 1);
  }
  if (seen != 0x1ffu) return false;
  if (gn_sums) {
    if (d->gn_groups <= 0 || d->n_out % d->gn_groups != 0 || residual) return false;
    const int cpg = d->n_out / d->gn_groups;
    if (cpg < 2 || BN % cpg != 0 || (cpg < 16 && (cpg & (cpg - 1)) != 0) || (cpg >= 16 && cpg % 16 != 0)) return false;
    if (d->rows_per_sample <= 0 || d->rows_per_sample % kSlTileM != 0) return false;
  }
  // worth it only when the items fill the GPU (smaller layers are latency bound: the generic kernel's
  // one-tile CTAs spread them over more SMs)
  const int TPI = d->H / (kSlTileM / d->W);
  const long items = (long)d->n_img * (TPI / R) * (d->n_out / BN);
  const int min_items = sl_env_int("VDN_SLAB_MIN_ITEMS", num_sms());
  // 16-pixel-wide levels (R = 2, half the weight-tile reuse) measured slower than the generic kernel (259 vs 223 us at
  // 1024 channels): production dispatch takes W >= 32 only; tests lower VDN_SLAB_MIN_ITEMS and reach every width
  if (d->W < 32 && min_items >= num_sms()) return false;
  return items >= min_items;
}

int slabconv_launch(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp, const float* bias,
                    const void* residual, const void* residual2, void* out, void* out2, float* gn_sums,
                    cudaStream_t st) {
  SlabArgs a;
  memset(&a, 0, sizeof(a));
  VDN_REQUIRE(slab_geometry(d, &a.BN, &a.R), VDN_E_SHAPE, "conv3x3_slab: unsupported geometry");
  const int C = d->src_c, W = d->W, H = d->H;
  const int kSlBK = sl_env_int("VDN_SLAB_BK", 32) == 16 ? 16 : 32;
  a.N = d->n_out;
  a.W = W;
  a.TR = kSlTileM / W;
  a.GPI = (H / a.TR) / a.R;
  a.n_src = d->n_src;
  a.C = C;
  a.chunks = C / kSlBK;
  a.n_ntiles = d->n_out / a.BN;
  a.n_items = d->n_img * a.GPI * a.n_ntiles;
  const int slab_rows = a.R * a.TR + 2;
  a.slab_bytes = slab_rows * W * kSlBK * 2;
  a.b_bytes = a.BN * kSlBK * 2;
  a.stage_bytes = a.slab_bytes + 3 * a.b_bytes;
  a.tmem_cols = 32;
  while (a.tmem_cols < a.R * a.BN) a.tmem_cols *= 2;
  for (int t = 0; t < 9; ++t) a.tap_of[(d->tap_dy[t] + 1) * 3 + (d->tap_dx[t] + 1)] = (signed char)t;
  a.bias = bias;
  a.res = reinterpret_cast<const bf16*>(residual);
  a.res2 = reinterpret_cast<const bf16*>(residual2);
  a.out = reinterpret_cast<bf16*>(out);
  a.out2 = reinterpret_cast<bf16*>(out2);
  a.split_col = d->split_col;
  a.ld_out = d->split_col > 0 ? d->split_col : d->n_out;
  a.ld_out2 = d->n_out - d->split_col;
  a.gn_sums = gn_sums;
  a.gn_groups = gn_sums ? d->gn_groups : 0;
  a.cpg = (gn_sums && d->gn_groups > 0) ? d->n_out / d->gn_groups : 1;
  a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  a.n_samples = std::max(1, (d->n_img * H * W) / a.rows_per_sample);

  a.dbg = sl_env_int("VDN_SLAB_DBG", 0);
  const int stg_bytes = 2 * (a.BN / 64) * 16384;  // two staging buffers of BN/64 swizzled 128 x 64 sub-tiles
  int S = kSlMaxStages;
  while (S > 2 && 1024 + S * a.stage_bytes + stg_bytes > 224 * 1024) --S;
  S = std::min(S, std::max(2, a.n_src * a.chunks * 3));
  if (tune_is_set("VDN_SLAB_S")) S = std::max(2, std::min(kSlMaxStages, tune_int("VDN_SLAB_S", S)));
  const int smem = 1024 + S * a.stage_bytes + stg_bytes;
  VDN_REQUIRE(smem <= 224 * 1024, VDN_E_SHAPE, "conv3x3_slab: shared memory %d B exceeds the SM", smem);
  a.S = S;

  SlabMaps maps;
  memset(&maps, 0, sizeof(maps));
  const void* srcs[2] = {src0, src1};
  int rc;
  for (int s = 0; s < d->n_src; ++s) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)d->n_img};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {(uint32_t)kSlBK, (uint32_t)W, (uint32_t)slab_rows, 1u};
    rc = encode_tmap_bf16(&maps.a[s], srcs[s], 4, dims, str, box, 2 * kSlBK);
    if (rc) return rc;
  }
  {
    const uint64_t ktot = (uint64_t)9 * d->n_src * C;
    const uint64_t dims[2] = {ktot, (uint64_t)d->n_out};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t bbox[2] = {(uint32_t)kSlBK, (uint32_t)a.BN};
    rc = encode_tmap_bf16(&maps.b, wp, 2, dims, str, bbox, 2 * kSlBK);
    if (rc) return rc;
  }
  {
    const uint64_t M = (uint64_t)d->n_img * H * W;
    const uint32_t obox[2] = {64u, (uint32_t)kSlTileM};
    const uint64_t dims0[2] = {(uint64_t)a.ld_out, M};
    const uint64_t str0[1] = {(uint64_t)a.ld_out * 2};
    rc = encode_tmap_bf16(&maps.o[0], out, 2, dims0, str0, obox, 128);
    if (rc) return rc;
    if (d->split_col > 0) {
      const uint64_t dims1[2] = {(uint64_t)a.ld_out2, M};
      const uint64_t str1[1] = {(uint64_t)a.ld_out2 * 2};
      rc = encode_tmap_bf16(&maps.o[1], out2, 2, dims1, str1, obox, 128);
      if (rc) return rc;
    }
  }
  VDN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!residual || (reinterpret_cast<uintptr_t>(residual) & 15) == 0) &&
                  (!out2 || (reinterpret_cast<uintptr_t>(out2) & 15) == 0) && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
              VDN_E_ALIGN, "conv3x3_slab: out/residual/bias must be 16B aligned");

  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_slab_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(conv3x3_slab_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  // CTAs per SM: limited by TMEM columns and shared memory
  const int cps = std::max(1, std::min(512 / a.tmem_cols, (227 * 1024) / (smem + 1024)));
  int grid = std::min(a.n_items, num_sms() * cps);
  if (tune_is_set("VDN_SLAB_GRID")) grid = std::max(1, std::min(a.n_items, tune_int("VDN_SLAB_GRID", grid)));  // tests: long runs per CTA
  cudaError_t le = kSlBK == 32
                       ? launch_pdl(conv3x3_slab_kernel<32>, dim3(grid), dim3(kSlThreads), (size_t)smem, st, 1, maps, a)
                       : launch_pdl(conv3x3_slab_kernel<16>, dim3(grid), dim3(kSlThreads), (size_t)smem, st, 1, maps, a);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "conv3x3_slab launch: %s", cudaGetErrorString(le));
  return check_launch("conv3x3_slab_kernel");
}

}  // namespace vdn
