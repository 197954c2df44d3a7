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

/* Reference (non tensor-core) implementation of the same contract, used only by the GPU
 * tests to localise faults. Not on any product path. */
int vdn_tapgemm_ref(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp,
                    const float* bias, const void* residual, const void* residual2, void* out, void* out2,
                    float* gn_sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDN_H_ */
