// Internal host-side declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vdn.h"

namespace vdn {

// Experiment / test switches (the "VDN_*" names of DESIGN.md section 7). Product code never reads the process
// environment on a launch path: a switch keeps its compiled-in default unless a test or tool sets it through the
// C ABI (vdn_debug_set). Only a -DVDN_DEBUG build also seeds the table from environment variables of the same names.
bool tune_is_set(const char* name);
int tune_int(const char* name, int dflt);  // the value if set, else dflt
bool tune_on(const char* name);            // set and non-zero

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
void tmap_cache_clear();
long long* debug_trace_ptr();  // runtime.cu: vdn_debug_trace_buffer (NULL unless a tool set it)
void comm_shutdown();  // comm.cu

// conv3x3_rows.cu: persistent row-ring (1,3,3) conv; applicable() decides, launch() has vdn_tapgemm's contract
bool rowconv_applicable(const vdn_tapgemm_desc* d, const void* residual, const float* gn_sums);
int rowconv_launch(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp, const float* bias,
                   const void* residual, const void* residual2, void* out, void* out2, float* gn_sums,
                   cudaStream_t st);

// conv3x3_slab.cu: (1,3,3) conv of the wide-channel levels with operand reuse in shared memory (R row-adjacent
// tiles per CTA share every weight tile, the three dy taps share one pixel slab); same contract
bool slabconv_applicable(const vdn_tapgemm_desc* d, const void* residual, const float* gn_sums);
int slabconv_launch(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp, const float* bias,
                    const void* residual, const void* residual2, void* out, void* out2, float* gn_sums,
                    cudaStream_t st);

// sla_mma.cu: per-token SpatialLinearAttention products on warp-level tensor-core MMAs
int sla_apply_mma_launch(const void* qkv, const float* ctx, void* tok_out, int n_img, int N, cudaStream_t st);
int sla_bwd_tokens_mma_launch(const void* qkv, const void* d_tok, const float* ctx, const float* dctx,
                              const float* kstat, void* dqkv, int n_img, int N, cudaStream_t st);

int sla_ctx_partial_mma_launch(const void* qkv, int N, int tokens_per_split, int n_split, float* ctx_part,
                               float* ms_part, int n_img, cudaStream_t st);
int sla_dctx_mma_launch(const void* qkv, const void* d_tok, int N, int tokens_per_split, int n_split, float* dctx,
                        int n_img, cudaStream_t st);

// mha_mma.cu: temporal attention core backward, one warp per (pixel, head), register-resident
int mha_temporal_mma_bwd_launch(const void* qkv, const void* d_o, const float* lse, void* dqkv, int B, int F, int H,
                                int W, cudaStream_t st);

int mha_temporal_mma_fwd_launch(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv, float* lse,
                                int B, int F, int H, int W, cudaStream_t st);

// mha_spatial_mma.cu: mid-block spatial attention core (sequences of HW = 64k tokens) on warp-level MMAs
bool mha_spatial_mma_applicable(int HW);
int mha_spatial_mma_fwd_launch(const void* qkv, void* o, float* lse, int n_seq, int S, cudaStream_t st);
int mha_spatial_mma_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int n_seq,
                               int S, cudaStream_t st);

// mha_folded_tc.cu: folded temporal attention block (inference, C = 32) with its shared-weight GEMMs on tcgen05
bool mha_folded_tc_applicable(int F, int HW);
int mha_folded_tc_launch(const void* x, const void* fa, const float* fu, const void* fm, const float* fb, void* out, int B,
                         int F, int H, int W, cudaStream_t st);

// mha_train_tc.cu: temporal attention forward of the training engines (C = 32), projection on tcgen05
bool mha_train_tc_applicable(int F, int HW);
int mha_train_tc_launch(const void* x, const void* w_hm, const float* bias_hm, void* o, void* qkv, float* lse, int B, int F,
                        int H, int W, cudaStream_t st);

// sla_apply_tc.cu: apply pass of the fused SpatialLinearAttention forward (inference, C = 32) on tcgen05
bool sla_apply_tc_applicable(int N);
size_t sla_apply_tc_scratch_bytes(int n_img);
int sla_apply_tc_launch(const void* x, const void* w_qkv, const void* w_out, const float* ctx, void* gt_ws, void* out,
                        int n_img, int N, cudaStream_t st);

int sla_ctx_fused_launch(const void* x, const void* w_qkv, int N, int tokens_per_split, int n_split, float* ctx_part,
                         float* ms_part, int n_img, cudaStream_t st);
int sla_apply_fused_launch(const void* x, const void* w_qkv, const void* w_out, const float* ctx, void* out, int n_img,
                           int N, cudaStream_t st);

// Kernel launch with programmatic dependent launch (and optionally a thread-block cluster along grid.x). The
// kernel must call pdl_wait() before touching global memory. VDN_NO_PDL=1 falls back to plain stream order.
bool pdl_enabled();
template <typename Kern, typename... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int num_sms() { return 148; }

}  // namespace vdn
