/*
 * vdn.h - C ABI of the B200-native (sm_100a) Unet3D / GaussianDiffusion hot path.
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * maxsonate/video-diffusion-nnx. The reference has no FFI of
 * its own (it is pure Python on flax.nnx / XLA); every entry point below replaces
 * the XLA lowering of the reference call site cited next to it. A maintainer binds
 * them with jax.ffi (INTEGRATION.md shows the shim) or, in this image, ctypes.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in _host.
 *  - Activations are channels-last: (B, F, H, W, C) flattened to (P, C), P = B*F*H*W,
 *    exactly the layout Unet3D works in after unet3d.py:280.
 *  - `act` tensors are bf16 (2 bytes/element) on the bf16 path; statistics, losses,
 *    master weights and weight gradients are fp32.
 *  - No entry point allocates, frees or synchronises; all work is enqueued on `stream`
 *    (a cudaStream_t passed as void*). All are CUDA-graph capturable.
 *  - Return value: 0 = OK, negative = vdn_status. vdn_last_error() gives a thread-local
 *    message.
 */
#ifndef VDN_H_
#define VDN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  VDN_OK = 0,
  VDN_E_SHAPE = -1, /* unsupported shape / argument */
  VDN_E_ALIGN = -2, /* pointer or leading dimension not aligned as required */
  VDN_E_ARCH = -3,  /* not an sm_100 device / driver entry point missing */
  VDN_E_CUDA = -4   /* a CUDA call failed; see vdn_last_error() */
} vdn_status;

int vdn_version(void);
const char* vdn_last_error(void);
/* kernels launched (or recorded into a CUDA graph under capture) by this library so far */
unsigned long long vdn_launch_count(void);
/* Test / tool hook: set (enable != 0) or clear one experiment switch by its VDN_* name.
 * Product code never reads the process environment; without this call every switch has its compiled-in default. */
int vdn_debug_set(const char* name, int value, int enable);

/* ---------------------------------------------------------------------------------
 * Weight repacking (fp32 reference layout -> bf16 K-major GEMM operand).
 * Reference layouts: nnx.Conv kernel (kd,kh,kw,in,out) [modules.py:162, utils.py:113,125],
 * nnx.Linear (in,out), nnx.LinearGeneral (in,heads,dim)/(heads,dim,out) [modules.py:261-276].
 * All of them are [taps][cin][cout] with taps = kh*kw (kd == 1).
 *
 * mode 0 (forward operand):  dst[(n_off+co)*ld + k_off + t*cin + ci]  = src[perm[t]][ci][co]
 * mode 1 (dgrad  operand):   dst[(n_off+ci)*ld + k_off + t*cout + co] = src[perm[t]][ci][co]
 * perm == NULL means identity. dst is bf16.
 * --------------------------------------------------------------------------------- */
int vdn_pack_weight(const float* src, void* dst, int taps, int cin, int cout, int mode, const int* perm_host,
                    int ld, int n_off, int k_off, void* stream);

/* ---------------------------------------------------------------------------------
 * Tap-GEMM: the one tensor-core kernel family behind every convolution and projection
 * on the path (tcgen05.mma, TMEM accumulators, TMA-fed implicit GEMM).
 *
 *   out[m, n] = sum_{tap, src, c} A_src[pixel(m) + shift(tap), c] * Wp[n, (tap, src, c)] + bias[n]
 *               (+ residual[m, n])
 *
 * kind selects the gather/scatter geometry:
 *   VDN_TAP_UNIT   taps are (dy,dx) shifts on the same H x W grid, zero padded:
 *                  conv (1,3,3) fwd + dgrad [modules.py:162-165], 1x1 convs and
 *                  Linear/LinearGeneral projections [modules.py:71-91,219-222,261-276]
 *   VDN_TAP_DOWN   (1,4,4) stride (1,2,2) SAME conv: input grid 2H x 2W, output H x W
 *                  [utils.py:125]; also the dgrad of the transposed conv
 *   VDN_TAP_UP     one output-parity class (py,px) of the (1,4,4) stride-2 transposed
 *                  conv: input grid H x W, output scattered into 2H x 2W [utils.py:113];
 *                  also the dgrad of the strided conv. 4 taps (2x2) per class.
 * --------------------------------------------------------------------------------- */
enum { VDN_TAP_UNIT = 0, VDN_TAP_DOWN = 1, VDN_TAP_UP = 2 };
enum { VDN_BF16 = 0, VDN_F32 = 1 };

typedef struct {
  int kind;        /* VDN_TAP_* */
  int n_img;       /* B*F images */
  int H, W;        /* grid the GEMM rows (M = n_img*H*W) are laid over */
  int n_src;       /* 1 or 2 (channel concat of two sources, h first: unet3d.py:346,377) */
  int src_c;       /* channels of EACH source (both equal) */
  int n_taps;      /* <= 16 */
  int tap_dy[16];  /* VDN_TAP_UNIT/UP: shifts in grid pixels. VDN_TAP_DOWN: kernel row index 0..3 */
  int tap_dx[16];
  int n_out;       /* GEMM N (output channels) */
  int py, px;      /* VDN_TAP_UP: output parity class */
  int out_dtype;   /* VDN_BF16 / VDN_F32 */
  int split_col;   /* 0, or: columns >= split_col go to out2 (dgrad of a concat input) */
  int gn_groups;   /* 0, or: accumulate per-(sample,group) sum / sum-of-squares of the output */
  int rows_per_sample; /* F*H*W, used with gn_groups */
} vdn_tapgemm_desc;

/* src0/src1: bf16 (n_img, Hs, Ws, src_c). wp: packed bf16 [n_out][n_taps*n_src*src_c].
 * bias: fp32 [n_out] or NULL. residual (residual2 for the out2 part): same dtype/shape as out, or NULL.
 * residual may alias out (each element is read then written by the same thread).
 * gn_sums: fp32 [B][gn_groups][2], must be zeroed by the caller; or NULL. */
int vdn_tapgemm(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp, const float* bias,
                const void* residual, const void* residual2, void* out, void* out2, float* gn_sums, void* stream);

/* Same, with caller-owned scratch of >= vdn_tapgemm_workspace(d) bytes (16-byte aligned; contents undefined before and
 * after; not to be shared by launches that may run concurrently on different streams). With it, launches whose output
 * has too few 128-row tiles to fill the SMs (the 8x8 level of config_v2_2) split the K loop over thread-block clusters
 * and reduce through the scratch; launches that need no scratch ignore it. workspace == NULL: as vdn_tapgemm. */
int vdn_tapgemm_ws(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp, const float* bias,
                   const void* residual, const void* residual2, void* out, void* out2, float* gn_sums, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Reference (non tensor-core) implementation of the same contract, used only by the GPU
 * tests to localise faults. Not on any product path. */
/* Debug only: device buffer of 8*3*64 int64 that the row-ring (1,3,3) conv kernel stamps with clock64()
 * for its first 8 CTAs (producer / MMA / epilogue timelines); NULL switches it off. */
void vdn_debug_rowconv_trace(void* dev_buf);
/* Same for the generic / split-K tap-GEMM kernels: 1024 int64 = 4 x 64 clock64 stamps of CTA 0 (producer 0 / producer 1 /
 * MMA issuer per K step; epilogue), then [256][3] global-timer stamps (prologue done, predecessor complete, exit) per CTA. */
void vdn_debug_tapgemm_trace(void* dev_buf);

/* Generic timeline hook (tools only): device buffer of >= 1024 int64 that kernels with a timeline stamp with clock64()
 * (mha_folded_tc: per tile of CTA 0, 8 stamps = tile start, x landed, Y ready, y staged, core done, O ready, stored). */
void vdn_debug_trace_buffer(void* dev_buf);

int vdn_tapgemm_ref(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp,
                    const float* bias, const void* residual, const void* residual2, void* out, void* out2,
                    float* gn_sums, void* stream);

/* ---------------------------------------------------------------------------------
 * Batched repack: every packed operand of a model in one launch. jobs_dev points to a device
 * array of vdn_pack_job; the work unit is a 32 x 32 (cin x cout) tile of one tap: begin = exclusive prefix sum of
 * taps*ceil(cin/32)*ceil(cout/32), total = the sum over all jobs.
 * --------------------------------------------------------------------------------- */
typedef struct {
  const float* src; /* fp32 [taps][cin][cout] */
  void* dst;        /* bf16 packed operand */
  int taps, cin, cout, mode, ld, n_off, k_off;
  int perm[16];
  long long begin;
} vdn_pack_job;
int vdn_pack_batched(const void* jobs_dev, int n_jobs, long long total, void* stream);

/* ---------------------------------------------------------------------------------
 * Weight gradient of every conv / projection (tcgen05, MN-major operands, split-K + fp32 atomics):
 *   dw[tap][n_src*C][Cout] += sum_pixels src[p + tap]^T g[p]     (reference kernel layout)
 * Replaces the XLA transpose of the call sites listed for vdn_tapgemm under
 * jax.value_and_grad (trainer.py:361). kind as in vdn_tapgemm:
 *   UNIT: src (n_img,H,W,C) x n_src, g (n_img,H,W,Cout), taps = (dy,dx) shifts
 *   DOWN: src (n_img,2H,2W,C), g (n_img,H,W,Cout), taps = kernel indices (ky,kx) in 0..3
 *   UP  : src (n_img,H,W,C), g (n_img,2H,2W,Cout), taps = kernel indices (a,b) in 0..3
 * --------------------------------------------------------------------------------- */
int vdn_wgrad(int kind, const void* src0, const void* src1, const void* g, float* dw, int n_img, int H, int W,
              int n_src, int C, int Cout, int n_taps, const int* tap_dy, const int* tap_dx, void* stream);
/* Same, plus the bias gradient dbias[Cout] += sum_pixels g[p] (fp32, may be NULL): computed inside the GEMM through
 * a spare "ones" M atom when there is one, otherwise by a vdn_colsum pass. */
int vdn_wgrad_bias(int kind, const void* src0, const void* src1, const void* g, float* dw, float* dbias, int n_img,
                   int H, int W, int n_src, int C, int Cout, int n_taps, const int* tap_dy, const int* tap_dx,
                   void* stream);
/* test-only CUDA-core reference of the same contract */
int vdn_wgrad_ref(int kind, const void* src0, const void* src1, const void* g, float* dw, int n_img, int H, int W,
                  int n_src, int C, int Cout, int n_taps, const int* tap_dy, const int* tap_dx, void* stream);

/* ---------------------------------------------------------------------------------
 * Block / ResnetBlock normalisation (modules.py:150-243). gn_sums: fp32 [16][B][G][2] replica
 * slots of (sum x, sum x^2) per (sample, group) as accumulated by vdn_tapgemm. flax semantics:
 * statistics over (F,H,W,C/G), eps 1e-6, var = max(0, E[x^2]-E[x]^2).
 *   gn_silu_fwd:        out = silu( GN(x_raw)*gamma+beta [ *(scale+1)+shift ] )     modules.py:171-179
 *   resblock_tail_fwd:  out = silu(GN(b_raw)*gamma+beta) + LayerNorm_C(s)           modules.py:241-242
 *   gn_silu_bwd:        dx_raw, dgamma+=, dbeta+=, dss = (dscale | dshift) [B][2C]; T_ws fp32 [B][C][2];
 *                       dconv_bias (optional, +=) = column sums of dx_raw = gradient of the conv bias
 *   ln_bwd:             ds, dgamma+=, dbeta+= of the per-pixel LayerNorm (norm_2)
 * scale_shift: fp32 rows [B][ss_ld], scale = [0,C), shift = [C,2C); NULL when the block has no time
 * embedding. All activation tensors bf16 [B][rows_per_sample][C].
 * --------------------------------------------------------------------------------- */
int vdn_gn_silu_fwd(const void* x_raw, const float* gn_sums, const float* gamma, const float* beta,
                    const float* scale_shift, int ss_ld, void* out, int B, int rows_per_sample, int C, int G,
                    void* stream);
int vdn_resblock_tail_fwd(const void* b_raw, const float* gn_sums, const float* gamma, const float* beta,
                          const void* s, const float* ln_gamma, const float* ln_beta, void* out, int B,
                          int rows_per_sample, int C, int G, void* stream);
int vdn_gn_silu_bwd(const void* dy, const void* x_raw, const float* gn_sums, const float* gamma, const float* beta,
                    const float* scale_shift, int ss_ld, float* T_ws, void* dx_raw, float* dgamma, float* dbeta,
                    float* dss, int dss_ld, float* dconv_bias, int B, int rows_per_sample, int C, int G, void* stream);
/* vdn_gn_silu_bwd with T_ws zeroed by the CALLER: a step's layers take slices of one buffer zeroed once, instead of a
 * memset node per call on the dependency chain. */
int vdn_gn_silu_bwd_acc(const void* dy, const void* x_raw, const float* gn_sums, const float* gamma, const float* beta,
                        const float* scale_shift, int ss_ld, float* T_ws, void* dx_raw, float* dgamma, float* dbeta,
                        float* dss, int dss_ld, float* dconv_bias, int B, int rows_per_sample, int C, int G, void* stream);
int vdn_ln_bwd(const void* s, const void* dy, const float* ln_gamma, void* ds, float* dgamma, float* dbeta, long P,
               int C, void* stream);

/* ---------------------------------------------------------------------------------
 * Attention cores on the fused projection qkv bf16 [P][768] = (q | k | v), 8 heads x 32.
 *   mha_core: MultiheadAttention core, modules.py:285-324 (q / sqrt(32), softmax over keys; the mask
 *     and bias branches never run inside Unet3D because PreNorm drops kwargs, modules.py:146-148).
 *     mode 0: sequences over frames per pixel ('b f h w c -> b (h w) f c', unet3d.py:86-96);
 *     mode 1: sequences over the pixels of a frame ('b f (h w) c', unet3d.py:196-205).
 *     o bf16 [P][256], lse fp32 [P][8]. bwd: D_ws fp32 [P][8] scratch, dqkv bf16 [P][768].
 *   sla_core: SpatialLinearAttention core, modules.py:105-123: q softmax over the 32 features
 *     (unscaled), k softmax over the N = H*W tokens, ctx = k~^T v, out = q~ ctx.
 *     tok_out bf16 [P][256], ctx fp32 [n_img][8][32][32], kstat fp32 [n_img][8][2][32] (col max, col sum),
 *     ws: fp32 scratch of vdn_sla_workspace_floats(n_img, N). bwd: dctx fp32 scratch like ctx.
 * --------------------------------------------------------------------------------- */
int vdn_mha_core_fwd(const void* qkv, void* o, float* lse, int mode, int B, int F, int HW, void* stream);
int vdn_mha_core_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* D_ws, void* dqkv,
                     int mode, int B, int F, int HW, void* stream);
size_t vdn_sla_workspace_floats(int n_img, int N);
int vdn_sla_core_fwd(const void* qkv, void* tok_out, float* ctx, float* kstat, float* ws, int n_img, int N,
                     void* stream);
int vdn_sla_core_bwd(const void* qkv, const void* d_tok, const float* ctx, const float* kstat, float* dctx,
                     void* dqkv, int n_img, int N, void* stream);
/* vdn_sla_core_bwd with dctx zeroed by the CALLER (see vdn_gn_silu_bwd_acc). */
int vdn_sla_core_bwd_acc(const void* qkv, const void* d_tok, const float* ctx, const float* kstat, float* dctx,
                         void* dqkv, int n_img, int N, void* stream);
/* Fused SpatialLinearAttention block forward for C == 32 (modules.py:99-129 incl. the to_q/k/v and to_out 1x1
 * convs and the residual of unet3d.py:170-178): out = x + to_out(SLA(x)) with no q/k/v/tok tensor in HBM. x, out
 * bf16 [P][32]; w_qkv packed bf16 [768][32] (vdn_pack_weight mode 0 of the fused q|k|v kernel), w_out packed bf16
 * [32][256]; ctx / kstat / ws as in vdn_sla_core_fwd. Used by inference engines (the backward needs q/k/v). */
int vdn_sla_fused_fwd(const void* x, const void* w_qkv, const void* w_out, void* out, float* ctx, float* kstat,
                      float* ws, int n_img, int N, int C, void* stream);

/* Fused temporal attention forward (unet3d.py:86-96,118-120 + modules.py:285-323): the QKV projection on
 * tensor cores and the F x F attention core in one kernel; qkv is not materialised unless asked for.
 * x bf16 (B,F,H,W,C); w_hm bf16 [768][C], bias_hm fp32 [768]: head-major repack of the fused qkv
 * projection (row h*96 + part*32 + d) produced by vdn_qkv_headmajor_pack from w fp32 [C][768] / bias [768];
 * o bf16 [P][256]; optional qkv bf16 [P][768] (q|k|v) and lse fp32 [P][8] for the backward. */
int vdn_qkv_headmajor_pack(const float* w, const float* bias, void* dst, float* bias_dst, int C, void* stream);
int vdn_mha_temporal_fused_fwd(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv, float* lse,
                               int B, int F, int H, int W, int C, void* stream);
/* Tensor-core version of vdn_mha_temporal_fused_fwd (same contract): S = Q K^T and O = P V also run on
 * tcgen05 over the whole pixel tile (block-diagonal use of a 128x128 score tile). Instantiated for F in
 * {10, 16} (config_v2_2 / v2_3x), C % 32 == 0; for C == 32 and any F <= 16 a register-resident warp-MMA kernel
 * (csrc/mha_mma.cu) takes over. vdn_mha_temporal_tc_supported tells. */
int vdn_mha_temporal_tc_supported(int F, int C);
int vdn_mha_temporal_tc_fwd(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv, float* lse,
                            int B, int F, int H, int W, int C, void* stream);
/* Temporal attention core forward on a materialised q|k|v tensor (modules.py:296-323: q / sqrt(32), softmax over the
 * F <= 16 frames of a pixel, P V), for levels whose QKV projection runs as a tap-GEMM (training engines at C >= 64):
 * qkv bf16 [P][768] (q|k|v, head h at columns h*32) -> o bf16 [P][256], optional lse fp32 [P][8]. One warp per
 * (pixel, head) on register-resident bf16 MMAs (csrc/mha_mma.cu). */
int vdn_mha_temporal_core_fwd(const void* qkv, void* o, float* lse, int B, int F, int H, int W, void* stream);
/* Folded temporal attention BLOCK for inference engines, C == 32, F <= 16: out = x + out_proj(MHA(x)) in one kernel
 * (modules.py:285-326 + the residual of unet3d.py:86-96). vdn_mha_fold_pack builds, from the fp32 master weights
 * (fused q|k|v kernel [32][768] + bias [768], out kernel [256][32] + bias [32]), A_h = W_q,h W_k,h^T / sqrt(32),
 * u_h = b_q,h W_k,h^T / sqrt(32), M_h = W_v,h W_o,h and b' = sum_h b_v,h W_o,h + b_o; softmax(q k^T) only depends on
 * (x A_h + u_h) x^T, and out = sum_h (P_h x) M_h + b'. x, out bf16 [P][32]. */
int vdn_mha_fold_pack(const float* w_qkv, const float* b_qkv, const float* w_out, const float* b_out, void* fa,
                      float* fu, void* fm, float* fb, void* stream);
int vdn_mha_temporal_folded_fwd(const void* x, const void* fa, const float* fu, const void* fm, const float* fb,
                                void* out, int B, int F, int H, int W, int C, void* stream);
/* Tensor-core temporal attention core backward (S, dP, dQ, dK, dV), F <= 16: one warp per (pixel, head) on
 * register-resident bf16 MMAs (csrc/mha_mma.cu). dbias (optional, fp32 [768], +=) receives the column sums of
 * dqkv = the gradient of the q|k|v projection bias (modules.py:261-270), which saves a pass over dqkv. */
int vdn_mha_temporal_tc_bwd(const void* qkv, const void* d_o, const float* lse, void* dqkv, float* dbias, int B, int F,
                            int H, int W, void* stream);
/* Temporal attention core backward in one kernel (smem exchange of k, v, q, dO; P and dS computed once):
 * qkv / lse / o from the forward, d_o bf16 [P][256] -> dqkv bf16 [P][768]. F <= 16. */
int vdn_mha_temporal_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int F,
                         int H, int W, void* stream);

/* ---------------------------------------------------------------------------------
 * MultiheadAttention with the reference's OPTIONAL inputs, and RelativePositionBias (modules.py:285-326, :330-390).
 * Dead inside Unet3D (PreNorm drops kwargs, modules.py:146-148) - the fused kernels above do not carry them - but
 * live when the module is called directly (test_modules.py:242-271). Literal semantics:
 *   attn = softmax_j(q k^T / sqrt(dim)); focus_present_mask[b] != 0: off-diagonal entries are REPLACED, after the
 *   softmax, by finfo(float32).min (:307-316); pos_bias [heads][S][S] is ADDED after the softmax (:320-321);
 *   copy_v != 0 is the all-focus early return out(v) (:291-292): o = v.
 * qkv rows are [3][heads][dim] (q | k | v), o rows [heads][dim]; dtype VDN_BF16 or VDN_F32; dim in {8,16,32,64}.
 * Token t of sequence s is row (s / inner) * S * inner + (s % inner) + t * inner: inner = 1 for (..., S, C) inputs,
 * inner = H*W for the temporal view 'b f h w c -> b (h w) f c' of a (B,F,H,W,C) tensor (unet3d.py:86-96).
 * mask_b: device bytes [n_seq / seqs_per_batch] or NULL.
 *
 * vdn_rel_pos_bias: out[h][i][j] = embedding[bucket(i - j)][h] with the static 32 / 128 bucket parameters that
 * RelativePositionBias.__call__ uses whatever its constructor got (modules.py:386); buckets_out (int32 [n][n],
 * optional) returns the ids (integer work, bit-exact).
 * --------------------------------------------------------------------------------- */
int vdn_mha_core_ext_fwd(const void* qkv, void* o, int dtype, int heads, int dim, int n_seq, int S, int inner,
                         const unsigned char* mask_b, int seqs_per_batch, const float* pos_bias, int copy_v,
                         void* stream);
int vdn_rel_pos_bias(const float* embedding, int n, int heads, float* out, int* buckets_out, void* stream);

/* ---------------------------------------------------------------------------------
 * Small layers. init conv: nnx.Conv(channels, dim, (1,k,k)) on x fp32 (B,Cin,F,H,W) (unet3d.py:110-115,
 * :280-282) -> bf16 (B*F,H,W,Cout); w fp32 [k*k][Cin][Cout]. final conv: nnx.Conv(dim, out, 1)
 * (unet3d.py:251): h bf16 [P][C] -> fp32 [P][Co]. time MLP: SinusoidalPosEmb -> Linear -> gelu(tanh) ->
 * Linear (modules.py:30-45, unet3d.py:128-133). time heads: per ResnetBlock
 * LayerNorm(Linear(silu(t))) -> (scale | shift) (modules.py:202-208,233-238), all heads in one launch
 * through a device table of vdn_time_head.
 * --------------------------------------------------------------------------------- */
typedef struct {
  const float* w;    /* [4*dim][n_out] */
  const float* b;    /* [n_out] */
  const float* ln_g; /* [n_out] */
  const float* ln_b; /* [n_out] */
  float* dw;         /* gradients (+=); may be NULL for forward-only use */
  float* db;
  float* dln_g;
  float* dln_b;
  int n_out;         /* 2 * cout */
  int off;           /* column offset into the [B][ss_ld] scale/shift buffer */
} vdn_time_head;

int vdn_init_conv_fwd(const float* x, const float* w, const float* bias, void* out, int B, int Cin, int F, int H,
                      int W, int Cout, int ks, void* stream);
int vdn_init_conv_wgrad(const float* x, const void* dy, float* dw, float* dbias, int B, int Cin, int F, int H, int W,
                        int Cout, int ks, void* stream);
int vdn_final_conv_fwd(const void* h, const float* w, const float* bias, float* out, long P, int C, int Co,
                       void* stream);
int vdn_final_conv_bwd(const void* h, const float* dout, const float* w, void* dh, float* dw, float* db, long P,
                       int C, int Co, void* stream);
int vdn_time_mlp_fwd(const int* time, const float* w1, const float* b1, const float* w2, const float* b2,
                     float* emb_out, float* h1_out, float* t_out, int B, int dim, void* stream);
int vdn_time_mlp_bwd(const float* dt, const float* emb, const float* h1, const float* w2, float* dw1, float* db1,
                     float* dw2, float* db2, float* dh1_ws, int B, int dim, void* stream);
int vdn_time_heads_fwd(const float* t, const void* heads_dev, int n_heads, float* e_pre, float* ss, int ss_ld,
                       int B, int td, void* stream);
int vdn_time_heads_bwd(const float* t, const void* heads_dev, int n_heads, const float* e_pre, const float* dss,
                       int ss_ld, float* de_ws, float* dt, int B, int td, void* stream);

/* ---------------------------------------------------------------------------------
 * Diffusion math on fp32 images (B,C,F,H,W); Unet output eps is (B,F,H,W,C).
 *   q_sample: sqrt_ac[t]*x + sqrt_1mac[t]*noise, optionally x <- 2x-1 first
 *             (gaussian_diffusion.py:401-420, :499 / utils.py:271-280)
 *   loss:     mean |pred-noise| (l1) or mean (pred-noise)^2 (l2) and d loss / d pred
 *             (gaussian_diffusion.py:460-466); *loss is overwritten
 *   p_sample: predict_start_from_noise -> clip -> q_posterior mean -> + (t!=0) exp(logvar/2) z
 *             (gaussian_diffusion.py:120-261)
 *   randn:    N(0,1) from Philox4x32-10 (seed, subsequence, element offset): stands in for
 *             jax.random.normal (gaussian_diffusion.py:254,309,445)
 * --------------------------------------------------------------------------------- */
int vdn_q_sample(const float* x_start, const float* noise, const int* t, const float* sqrt_ac, const float* sqrt_1mac,
                 float* out, int B, long per_sample, int normalize, void* stream);
int vdn_loss(const float* pred, const float* noise, float* loss, float* dpred, int B, int C, long FHW, int l1,
             void* stream);
int vdn_p_sample(const float* x, const float* eps, const float* z, const int* t, const float* recip,
                 const float* recipm1, const float* coef1, const float* coef2, const float* logvar, float* out, int B,
                 int C, long FHW, int clip, void* stream);
int vdn_randn(float* out, long n, unsigned long long seed, unsigned long long subseq, unsigned long long elem_offset,
              void* stream);
/* Device-resident sampling loop (p_sample_loop, gaussian_diffusion.py:311-316): the same draws as vdn_randn with
 * (seed, subseq_add, elem_offset) = params_dev[0..2] and subseq = subseq_add + t_dev[0], all read on the device
 * (elem_offset must be a multiple of 4), and t_dev[b] -= 1; both are captured in the per-timestep CUDA graph so
 * that a timestep is one graph replay with no host-side argument update, for any key. */
int vdn_randn_t(float* out, long n, const unsigned long long* params_dev, const int* t_dev, void* stream);
int vdn_countdown(int* t_dev, int B, void* stream);

/* ---------------------------------------------------------------------------------
 * Training glue: bias gradients (column sums of a bf16 [P][C] gradient, +=), bf16 add, and the fused
 * optax.adam + EMA update over the flat fp32 state (trainer.py:367-382).
 * hp_dev (device, 9 floats): lr, b1, b2, eps, 1-b1^t, 1-b2^t, ema_decay, do_ema, grad_scale.
 * --------------------------------------------------------------------------------- */
int vdn_colsum(const void* dy, float* db, long P, int C, void* stream);
int vdn_add_bf16(const void* a, const void* b, void* out, long n, void* stream);
int vdn_adam_ema(float* p, const float* g, float* m, float* v, float* ema, const float* hp_dev, long n, void* stream);
/* Global-norm clip folded into the same update (utils.py:127-152, `max_grad_norm` of the trainer configs):
 * vdn_grad_sqnorm writes out[0] = sum g^2 of the flat gradient (after the all-reduce); vdn_adam_ema_clip reads
 * hp_dev[9] = max_grad_norm (0 = off) and hp_dev[10] = the reference's epsilon (1e-6):
 * l2 = sqrt(out[0] * grad_scale^2 + eps), g *= min(max_grad_norm / (l2 + eps), 1). */
int vdn_grad_sqnorm(const float* g, long n, float* out, void* stream);
int vdn_adam_ema_clip(float* p, const float* g, float* m, float* v, float* ema, const float* hp_dev,
                      const float* sqnorm_dev, long n, void* stream);

/* ---------------------------------------------------------------------------------
 * fp32-grade forward path (north_star: loss and predicted noise within 1e-3 of the fp32 reference; modules.py computes
 * in float32 throughout). Activations stay fp32; every GEMM is a split-bf16 product on the SAME tcgen05 tap-GEMM:
 *   x w ~= [x_hi | x_lo] [w_hi ; w_hi] + x_hi w_lo,  hi = bf16(v), lo = bf16(v - hi)
 * = vdn_tapgemm with the (hi, lo) pair as its two K-concatenated sources and fp32 output, then a second vdn_tapgemm
 * accumulating through the residual operand. These entry points supply the rest: the split (with optional channel
 * concat of two sources, unet3d.py:346,377), GroupNorm statistics / apply + SiLU (modules.py:171-179), the ResnetBlock
 * tail (modules.py:241-242), the SpatialLinearAttention core (modules.py:105-123), init / final convs
 * (unet3d.py:110-115, :251). The MultiheadAttention core is vdn_mha_core_ext_fwd with dtype VDN_F32.
 * sums: double [B][G][2] = (sum x, sum x^2). Forward only.
 * --------------------------------------------------------------------------------- */
/* weights: hi = float(bf16(w)), lo = w - hi as fp32 tensors, the sources vdn_pack_weight packs into w_hi / w_lo */
int vdn_f32_hilo(const float* src, float* hi, float* lo, long n, void* stream);
int vdn_f32_split(const float* src0, const float* src1, void* hi, void* lo, long P, int C0, int C1, void* stream);
int vdn_f32_gn_stats(const float* x, double* sums, int B, int rows_per_sample, int C, int G, void* stream);
int vdn_f32_gn_silu(const float* x, const double* sums, const float* gamma, const float* beta, const float* scale_shift,
                    int ss_ld, float* out, int B, int rows_per_sample, int C, int G, void* stream);
int vdn_f32_tail(const float* b_raw, const double* sums, const float* gamma, const float* beta, const float* s,
                 const float* ln_gamma, const float* ln_beta, float* out, int B, int rows_per_sample, int C, int G,
                 void* stream);
int vdn_f32_sla_core(const float* qkv, float* tok_out, float* ctx, int n_img, int N, void* stream);
int vdn_f32_init_conv(const float* x, const float* w, const float* bias, float* out, int B, int Cin, int F, int H, int W,
                      int Cout, int ks, void* stream);
int vdn_f32_final_conv(const float* h, const float* w, const float* bias, float* out, long P, int C, int Co, void* stream);

/* ---------------------------------------------------------------------------------
 * Scratch sizes. Every buffer (operands, results, scratch) is allocated and owned by the caller (XLA allocates
 * scratch as an extra result of the custom call); these return the BYTES of the scratch argument of the
 * entry point of the same name. Ops not listed need none.
 * --------------------------------------------------------------------------------- */
size_t vdn_sla_core_fwd_workspace(int n_img, int N);       /* ws of vdn_sla_core_fwd / vdn_sla_fused_fwd */
size_t vdn_sla_core_bwd_workspace(int n_img);              /* dctx */
size_t vdn_gn_silu_bwd_workspace(int B, int C);            /* T_ws */
size_t vdn_mha_core_bwd_workspace(long P);                 /* D_ws */
size_t vdn_time_heads_bwd_workspace(int B, int ss_ld);     /* de_ws */
size_t vdn_time_mlp_bwd_workspace(int B, int dim);         /* dh1_ws */
size_t vdn_tapgemm_workspace(const vdn_tapgemm_desc* d);   /* scratch of vdn_tapgemm_ws; 0 when the launch does not split K */

/* ---------------------------------------------------------------------------------
 * Data-parallel gradient exchange. The reference shards the batch over the `data`
 * mesh axis and GSPMD inserts the gradient all-reduce of the pjit'd train step (trainer.py:307-326,363-364);
 * here it is explicit: one communicator per process (one process per GPU), sum-reduce of one contiguous bucket
 * of the flat gradient, enqueued on `stream` (CUDA-graph capturable). The mean's 1/world is folded into
 * vdn_adam_ema (hp_dev[8]). NCCL is resolved at run time; without it these return VDN_E_ARCH.
 *   id exchange: rank 0 calls vdn_comm_unique_id(id_host) (vdn_comm_unique_id_bytes() bytes) and ships the bytes
 *   to every rank out of band (torch.distributed / MPI / a file); every rank then calls vdn_comm_init.
 *   max_ctas > 0 caps the SMs NCCL may occupy (the reduction overlaps the backward pass).
 * --------------------------------------------------------------------------------- */
int vdn_comm_unique_id_bytes(void);
int vdn_comm_unique_id(void* id_host);
int vdn_comm_init(void** comm_out, const void* id_host, int rank, int world, int max_ctas);
int vdn_comm_world(const void* comm, int* rank, int* world);
int vdn_allreduce_bucket(void* comm, void* buf, long count, int dtype, void* stream);
int vdn_comm_destroy(void* comm);

/* Library-owned state: cached TMA descriptors (keyed by pointer + shape; hits / misses since load) and
 * communicators. vdn_shutdown releases both; the library never owns device memory. */
int vdn_tmap_cache_stats(unsigned long long* hits, unsigned long long* misses);
int vdn_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* VDN_H_ */
