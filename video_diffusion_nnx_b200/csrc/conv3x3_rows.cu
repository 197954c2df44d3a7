// Row-ring (1,3,3) convolution: persistent tcgen05 implicit GEMM with resident weights.
//
// Serves the high-resolution (1,3,3) convs of the Unet3D path (Block.proj, modules.py:162-165, forward
// and dgrad) where the generic tap-GEMM re-reads every input pixel nine times from L2. Here a CTA owns a
// contiguous run of 128-pixel tiles (TR = 128/W full image rows each) and keeps
//   * the whole packed weight matrix resident in shared memory (loaded once per CTA), and
//   * per (source, dx) a ring of image rows, each row loaded ONCE per dx shift by TMA (x shifted by dx,
//     out-of-image pixels zero-filled = SAME padding), so the three dy taps of a tile are just three
//     row-offset views of the same ring: L2->SM traffic drops from 9x to ~3.3x of the input.
// The accumulator is double-buffered in TMEM, so the epilogue (bias / residual / GroupNorm partial sums /
// bf16 store) of tile j overlaps the MMAs of tile j+1.
//
// Ring geometry: stage = TR rows. Tile i of an image needs padded rows [TR*i, TR*i + TR + 1]
// (padded row p = image row p-1) = the last two rows of stage i-1 and all of stage i, where stage i holds
// padded rows [TR*i + 2, TR*i + TR + 1]. The A operand of tap dy starts (dy - 1) rows before stage i. A
// copy of the first TR-1 rows of slot 0 is kept behind slot S-1 ("tail") so that windows that straddle
// the ring wrap stay contiguous for the UMMA descriptor.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

constexpr int kRcThreads = 192;
constexpr int kRcMaxStages = 6;
constexpr int kRcTileM = 128;

struct RowConvMaps {
  CUtensorMap a[2];   // per source: box (KC, W, TR, 1)
  CUtensorMap at[2];  // per source: box (KC, W, TR-1, 1)  (ring tail copy)
  CUtensorMap b;      // packed weights [N][9*n_src*KC], box (KC, N)
};

struct RowConvArgs {
  int n_tiles, TPI, TR, W, H;
  int S;
  int row_bytes, region_bytes;
  uint32_t stage_tx, tail_tx;
  unsigned long long wblk_code;  // 4 bits per (dy+1)*3 + (dx+1): index of the weight block holding that tap
  const float* bias;
  const bf16* res;
  const bf16* res2;
  bf16* out;
  bf16* out2;
  int split;  // 1: columns [N/2, N) go to out2 (dgrad of a concat input)
  float* gn_sums;
  int rows_per_sample, n_samples;
  long long* trace;  // optional debug timeline: [cta < 8][role 3][64] clock64 stamps
};

// Adds this warp's per-thread GroupNorm partial sums (8 groups) to gn_sums[sample] and clears them.
__device__ __forceinline__ void rc_flush_gn(float (&gs1)[8], float (&gs2)[8], float* dst, int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float s1 = warp_sum(gs1[g]), s2 = warp_sum(gs2[g]);
    if (lane == 0) {
      atomicAdd(dst + 2 * g, s1);
      atomicAdd(dst + 2 * g + 1, s2);
    }
    gs1[g] = gs2[g] = 0.f;
  }
}

#define RC_TRACE(role, idx)                                                                     \
  do {                                                                                          \
    if (a.trace && blockIdx.x < 8 && (idx) < 64) a.trace[(blockIdx.x * 3 + (role)) * 64 + (idx)] = clock64(); \
  } while (0)

// 64-bit UMMA shared-memory descriptor from its two halves (only the low word changes between operands).
__device__ __forceinline__ uint64_t rc_desc(uint32_t hi, uint32_t lo) {
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

template <int N, int KC, int NSRC>
__global__ void __launch_bounds__(kRcThreads) conv3x3_rows_kernel(const __grid_constant__ RowConvMaps maps,
                                                                  const RowConvArgs a) {
  constexpr int kSw = KC * 2;  // bytes per pixel of one source == swizzle span
  constexpr uint32_t kLayout = umma_layout_type(kSw);
  constexpr uint32_t kSBO = 8 * kSw;
  constexpr int kWBlock = N * KC * 2;  // one (tap, source) weight block
  constexpr uint32_t kTmemCols = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 256;
  constexpr int kCpg = N / 8;  // GroupNorm: 8 groups (modules.py:167)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kRcMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kRcMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t wfull_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  uint8_t* wsm = smem;
  uint8_t* regions = smem + 9 * NSRC * kWBlock;
  const int S = a.S, TR = a.TR, TPI = a.TPI;
  const int t0 = (int)((long)blockIdx.x * a.n_tiles / gridDim.x);
  const int t1 = (int)((long)(blockIdx.x + 1) * a.n_tiles / gridDim.x);

  pdl_trigger();
  if (a.trace && threadIdx.x == 0) {
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.trace[8 * 3 * 64 + blockIdx.x * 4 + 0] = (long long)gt;
    a.trace[8 * 3 * 64 + blockIdx.x * 4 + 2] = smid;
    a.trace[8 * 3 * 64 + blockIdx.x * 4 + 3] = clock64();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSRC; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.at[s]);
    }
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
    }
    mbar_init(&wfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // everything above overlapped the tail of the previous kernel; global memory is touched below

  if (warp == 0) {
    // ================= TMA producer =================
    if (t0 < t1 && elect_one()) {
      constexpr int n_blocks = 9 * NSRC;
      mbar_expect_tx(&wfull_bar, (uint32_t)(n_blocks * kWBlock));
#pragma unroll
      for (int blk = 0; blk < n_blocks; ++blk) tma_load_2d(wsm + blk * kWBlock, &maps.b, &wfull_bar, blk * KC, 0);
      RC_TRACE(0, 0);
      int slot = 0, n_loaded = 0;
      uint32_t ph = 0;  // phase of the ring pass the slot counter is in
      const int stage_bytes = TR * a.row_bytes;
      auto load_stage = [&](int n, int y) {
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        const bool tail = slot == 0;
        mbar_expect_tx(&full_bar[slot], a.stage_tx + (tail ? a.tail_tx : 0u));
        uint8_t* dst = regions + slot * stage_bytes;
#pragma unroll
        for (int s = 0; s < NSRC; ++s)
#pragma unroll
          for (int dxi = 0; dxi < 3; ++dxi) {
            uint8_t* reg = dst + (s * 3 + dxi) * a.region_bytes;
            tma_load_4d(reg, &maps.a[s], &full_bar[slot], 0, dxi - 1, y, n);
            if (tail) tma_load_4d(reg + S * stage_bytes, &maps.at[s], &full_bar[slot], 0, dxi - 1, y, n);
          }
        if (++slot == S) {
          slot = 0;
          ph ^= 1u;
        }
        ++n_loaded;
        RC_TRACE(0, n_loaded);
      };
      int n = t0 / TPI, i = t0 - n * TPI;
      for (int t = t0; t < t1; ++t) {
        if (t == t0 || i == 0) load_stage(n, TR * i + 1 - TR);  // priming stage: the two rows above the tile
        load_stage(n, TR * i + 1);
        if (++i == TPI) {
          i = 0;
          ++n;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (t0 < t1 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kRcTileM, N, 0, 0);
      // descriptor halves: hi = SBO | version | swizzle mode; lo = (address >> 4) | LBO(16 B) << 16
      const uint32_t desc_hi = (kSBO >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t a_lo0 = (smem_u32(regions) >> 4) | (1u << 16);
      const uint32_t b_lo0 = (smem_u32(wsm) >> 4) | (1u << 16);
      const uint32_t row16 = (uint32_t)a.row_bytes >> 4, region16 = (uint32_t)a.region_bytes >> 4;
      const uint32_t ring_rows = (uint32_t)(S * TR);
      RC_TRACE(1, 0);
      mbar_wait(&wfull_bar, 0);
      RC_TRACE(1, 1);
      int slot = 0;        // ring slot of the next stage to consume
      uint32_t ph = 0;     // its phase
      int prev_slot = 0;
      int i = t0 % TPI;
      uint32_t tph0 = 1u, tph1 = 1u;  // tempty parity to wait for, per accumulator buffer
      for (int t = t0, j = 0; t < t1; ++t, ++j) {
        if (t == t0 || i == 0) {  // priming stage of a run
          mbar_wait(&full_bar[slot], ph);
          prev_slot = slot;
          if (++slot == S) { slot = 0; ph ^= 1u; }
        }
        const int cslot = slot;
        mbar_wait(&full_bar[cslot], ph);
        if (++slot == S) { slot = 0; ph ^= 1u; }
        RC_TRACE(1, 2 + 3 * j);
        const int buf = j & 1;
        mbar_wait(&tempty_bar[buf], buf ? tph1 : tph0);
        if (buf) tph1 ^= 1u; else tph0 ^= 1u;
        tc_fence_after();
        RC_TRACE(1, 3 + 3 * j);
        const uint32_t tacc = tmem_base + (uint32_t)(buf * N);
        // A window of tap dy starts (dy - 1) rows before the current stage (wrapping into the ring tail copy)
        const uint32_t r_cur = (uint32_t)(cslot * TR);
        const uint32_t r_m1 = r_cur >= 2 ? r_cur - 2 : r_cur + ring_rows - 2;
        const uint32_t r_0 = r_cur >= 1 ? r_cur - 1 : r_cur + ring_rows - 1;
        const uint32_t arow[3] = {a_lo0 + r_m1 * row16, a_lo0 + r_0 * row16, a_lo0 + r_cur * row16};
        unsigned long long code = a.wblk_code;
#pragma unroll
        for (int dyi = 0; dyi < 3; ++dyi)
#pragma unroll
          for (int dxi = 0; dxi < 3; ++dxi) {
            const uint32_t blk = (uint32_t)code & 15u;
            code >>= 4;
            const uint32_t b_lo = b_lo0 + blk * (uint32_t)(NSRC * (kWBlock >> 4));
#pragma unroll
            for (int s = 0; s < NSRC; ++s) {
              const uint32_t a_lo = arow[dyi] + (uint32_t)(s * 3 + dxi) * region16;
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma_bf16(tacc, rc_desc(desc_hi, a_lo + 2u * k), rc_desc(desc_hi, b_lo + (uint32_t)(s * (kWBlock >> 4)) + 2u * k),
                          idesc, (dyi | dxi | s | k) != 0 ? 1u : 0u);
            }
          }
        RC_TRACE(1, 40 + j);
        if (++i == TPI) i = 0;
        const bool run_ends = (t + 1 == t1) || i == 0;
        tc_commit(&empty_bar[prev_slot]);
        if (run_ends) tc_commit(&empty_bar[cslot]);
        tc_commit(&tfull_bar[buf]);
        prev_slot = cslot;
        RC_TRACE(1, 4 + 3 * j);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (warps 2..5) =================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    float gs1[8], gs2[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) gs1[g] = gs2[g] = 0.f;
    int cur_sample = -1;
    const int rep = blockIdx.x % kGnReplicas;
    float bias_r[N];
#pragma unroll
    for (int q = 0; q < N; ++q) bias_r[q] = a.bias ? __ldg(a.bias + q) : 0.f;
    for (int t = t0, j = 0; t < t1; ++t, ++j) {
      const int buf = j & 1;
      if (threadIdx.x == 64) RC_TRACE(2, 3 * j);
      mbar_wait(&tfull_bar[buf], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      if (threadIdx.x == 64) RC_TRACE(2, 1 + 3 * j);
      uint32_t raw[N];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * N);
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 32) tmem_ld_32x32(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&raw[c0]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);  // accumulator buffer is free for tile j+2

      const long m = (long)t * kRcTileM + r;
      if (a.gn_sums) {
        const int sample = (int)(((long)t * kRcTileM) / a.rows_per_sample);
        if (sample != cur_sample) {
          if (cur_sample >= 0) rc_flush_gn(gs1, gs2, a.gn_sums + ((long)(rep * a.n_samples + cur_sample) * 8) * 2, lane);
          cur_sample = sample;
        }
      }
#pragma unroll
      for (int h = 0; h < N / 32; ++h) {
        float v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(raw[h * 32 + q]) + bias_r[h * 32 + q];
        if (a.gn_sums) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int g = (h * 32 + q) / kCpg;
            gs1[g] += v[q];
            gs2[g] += v[q] * v[q];
          }
        }
        bf16* op;
        const bf16* rp;
        long eoff;
        if (a.split) {  // N == 64 split into two 32-channel tensors
          op = h == 0 ? a.out : a.out2;
          rp = h == 0 ? a.res : a.res2;
          eoff = m * 32;
        } else {
          op = a.out;
          rp = a.res;
          eoff = m * N + h * 32;
        }
        if (rp) {
          const uint4* r4 = reinterpret_cast<const uint4*>(rp + eoff);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = r4[q];
            float2 f;
            f = unpack_bf16x2(u.x); v[8 * q + 0] += f.x; v[8 * q + 1] += f.y;
            f = unpack_bf16x2(u.y); v[8 * q + 2] += f.x; v[8 * q + 3] += f.y;
            f = unpack_bf16x2(u.z); v[8 * q + 4] += f.x; v[8 * q + 5] += f.y;
            f = unpack_bf16x2(u.w); v[8 * q + 6] += f.x; v[8 * q + 7] += f.y;
          }
        }
        uint4* o4 = reinterpret_cast<uint4*>(op + eoff);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
          u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
          u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
          u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
          o4[q] = u;
        }
      }
      if (threadIdx.x == 64) RC_TRACE(2, 2 + 3 * j);
    }
    if (a.gn_sums && cur_sample >= 0)
      rc_flush_gn(gs1, gs2, a.gn_sums + ((long)(rep * a.n_samples + cur_sample) * 8) * 2, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (a.trace && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[8 * 3 * 64 + blockIdx.x * 4 + 1] = (long long)gt;
  }
}

template <int N, int KC, int NSRC>
static int launch_rowconv(const RowConvMaps& maps, const RowConvArgs& a, int grid, int smem_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(conv3x3_rows_kernel<N, KC, NSRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    VDN_REQUIRE(e == cudaSuccess, VDN_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  cudaError_t le = launch_pdl(conv3x3_rows_kernel<N, KC, NSRC>, dim3(grid), dim3(kRcThreads), (size_t)smem_bytes, st, 1, maps, a);
  VDN_REQUIRE(le == cudaSuccess, VDN_E_CUDA, "conv3x3_rows launch: %s", cudaGetErrorString(le));
  return check_launch("conv3x3_rows_kernel");
}

static long long* g_rc_trace = nullptr;

bool rowconv_applicable(const vdn_tapgemm_desc* d, const void* residual, const float* gn_sums) {
  if (tune_on("VDN_NO_ROWCONV")) return false;
  if (d->kind != VDN_TAP_UNIT || d->n_taps != 9 || d->out_dtype != VDN_BF16) return false;
  if (d->W != 64 && d->W != 32) return false;
  const int TR = kRcTileM / d->W;
  if (d->H % TR != 0) return false;
  if (d->src_c != 32 && d->src_c != 64) return false;
  if (d->n_out != 32 && d->n_out != 64) return false;
  if (d->src_c == 64 && (d->n_src != 1 || d->n_out != 64 || d->W != 32)) return false;  // 64 -> 64 at the 32-wide level
  if (d->split_col != 0 && (d->n_out != 64 || d->split_col != 32)) return false;
  unsigned seen = 0;  // the taps must be a permutation of the 3x3 neighbourhood
  for (int t = 0; t < 9; ++t) {
    if (d->tap_dy[t] < -1 || d->tap_dy[t] > 1 || d->tap_dx[t] < -1 || d->tap_dx[t] > 1) return false;
    seen |= 1u << ((d->tap_dy[t] + 1) * 3 + d->tap_dx[t] + 1);
  }
  if (seen != 0x1ffu) return false;
  if (gn_sums) {
    if (d->gn_groups != 8 || d->rows_per_sample % kRcTileM != 0 || residual) return false;
  }
  return true;
}

int rowconv_launch(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp, const float* bias,
                   const void* residual, const void* residual2, void* out, void* out2, float* gn_sums,
                   cudaStream_t st) {
  const int KC = d->src_c, N = d->n_out, W = d->W, H = d->H;
  RowConvArgs a;
  memset(&a, 0, sizeof(a));
  a.TR = kRcTileM / W;
  a.TPI = H / a.TR;
  a.n_tiles = d->n_img * a.TPI;
  a.W = W;
  a.H = H;
  a.row_bytes = W * KC * 2;
  for (int t = 0; t < 9; ++t)  // canonical (dy, dx) position -> weight block (= tap index in the packed operand)
    a.wblk_code |= (unsigned long long)t << (4 * ((d->tap_dy[t] + 1) * 3 + (d->tap_dx[t] + 1)));
  a.bias = bias;
  a.res = reinterpret_cast<const bf16*>(residual);
  a.res2 = reinterpret_cast<const bf16*>(residual2);
  a.out = reinterpret_cast<bf16*>(out);
  a.out2 = reinterpret_cast<bf16*>(out2);
  a.split = d->split_col != 0;
  a.gn_sums = gn_sums;
  a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  a.n_samples = std::max(1, (d->n_img * H * W) / a.rows_per_sample);
  a.trace = g_rc_trace;

  // stages / CTAs per SM: two co-resident CTAs hide each other's pipeline bubbles when three stages fit
  const int w_bytes = 9 * d->n_src * N * KC * 2;
  auto smem_for = [&](int S) { return 1024 + w_bytes + d->n_src * 3 * (S * a.TR + a.TR - 1) * a.row_bytes; };
  int cps = 2, S = 3;
  if (smem_for(3) > 112 * 1024) {
    cps = 1;
    S = 4;
    while (S > 2 && smem_for(S) > 220 * 1024) --S;
  }
  if (tune_is_set("VDN_RC_S")) S = std::max(2, std::min(kRcMaxStages, tune_int("VDN_RC_S", S)));
  if (tune_is_set("VDN_RC_CPS")) cps = std::max(1, std::min(4, tune_int("VDN_RC_CPS", cps)));
  VDN_REQUIRE(smem_for(S) <= 224 * 1024, VDN_E_SHAPE, "conv3x3_rows: shared memory %d B exceeds the SM", smem_for(S));
  a.S = S;
  a.region_bytes = (S * a.TR + a.TR - 1) * a.row_bytes;
  a.stage_tx = (uint32_t)(d->n_src * 3 * a.TR * a.row_bytes);
  a.tail_tx = (uint32_t)(d->n_src * 3 * (a.TR - 1) * a.row_bytes);

  RowConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  const void* srcs[2] = {src0, src1};
  int rc;
  for (int s = 0; s < d->n_src; ++s) {
    const uint64_t dims[4] = {(uint64_t)KC, (uint64_t)W, (uint64_t)H, (uint64_t)d->n_img};
    const uint64_t str[3] = {(uint64_t)KC * 2, (uint64_t)W * KC * 2, (uint64_t)H * W * KC * 2};
    const uint32_t box[4] = {(uint32_t)KC, (uint32_t)W, (uint32_t)a.TR, 1u};
    const uint32_t tbox[4] = {(uint32_t)KC, (uint32_t)W, (uint32_t)(a.TR - 1), 1u};
    rc = encode_tmap_bf16(&maps.a[s], srcs[s], 4, dims, str, box, KC * 2);
    if (rc) return rc;
    rc = encode_tmap_bf16(&maps.at[s], srcs[s], 4, dims, str, tbox, KC * 2);
    if (rc) return rc;
  }
  {
    const uint64_t ktot = (uint64_t)9 * d->n_src * KC;
    const uint64_t dims[2] = {ktot, (uint64_t)N};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t bbox[2] = {(uint32_t)KC, (uint32_t)N};
    rc = encode_tmap_bf16(&maps.b, wp, 2, dims, str, bbox, KC * 2);
    if (rc) return rc;
  }
  VDN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!residual || (reinterpret_cast<uintptr_t>(residual) & 15) == 0) &&
                  (!out2 || (reinterpret_cast<uintptr_t>(out2) & 15) == 0) && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
              VDN_E_ALIGN, "conv3x3_rows: out/residual/bias must be 16B aligned");
  int grid = std::min(a.n_tiles, num_sms() * cps);
  if (tune_is_set("VDN_RC_GRID")) grid = std::max(1, std::min(a.n_tiles, tune_int("VDN_RC_GRID", grid)));  // tests: long runs per CTA
  const int smem = smem_for(S);
  if (KC == 64) return launch_rowconv<64, 64, 1>(maps, a, grid, smem, st);
  if (N == 32 && d->n_src == 1) return launch_rowconv<32, 32, 1>(maps, a, grid, smem, st);
  if (N == 32) return launch_rowconv<32, 32, 2>(maps, a, grid, smem, st);
  if (d->n_src == 1) return launch_rowconv<64, 32, 1>(maps, a, grid, smem, st);
  return launch_rowconv<64, 32, 2>(maps, a, grid, smem, st);
}

}  // namespace vdn

// Debug only: device buffer of 8*3*64 int64 that the row-ring conv kernel stamps with clock64() (NULL = off).
extern "C" void vdn_debug_rowconv_trace(void* dev_buf) { vdn::g_rc_trace = reinterpret_cast<long long*>(dev_buf); }
