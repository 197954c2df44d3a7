// XLA-FFI view of the C ABI in include/vdn.h: one `XLA_FFI_Error* (*)(XLA_FFI_CallFrame*)` symbol per entry point,
// to be registered from Python with jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(sym), platform="CUDA")
// (video_diffusion_nnx_b200/jax_ffi.py does that) and invoked with jax.ffi.ffi_call.
//
// Call sites these targets replace in the reference (maxsonate/video-diffusion-nnx): the body of `loss_fn` under
// jax.value_and_grad in the pjit'd train step (trainer.py:337-361) and of `pjit_step` in p_sample_loop
// (gaussian_diffusion.py:299-301), i.e. the XLA lowering of modules.py / unet3d.py / gaussian_diffusion.py.
//
// Build (where jaxlib is installed; NOT buildable in the image this repo was developed in - no jaxlib, no XLA
// headers - so this file is compiled there only in its stub form, see the #else branch):
//   g++ -O2 -std=c++17 -shared -fPIC ffi/vdn_ffi.cc -Iinclude -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include -Lvideo_diffusion_nnx_b200 -lvdn -Wl,-rpath,'$ORIGIN' -o video_diffusion_nnx_b200/libvdn_ffi.so
//
// Conventions: buffers arrive as ffi::AnyBuffer (the C ABI checks shapes / alignment itself and reports through
// vdn_last_error()); OPTIONAL operands of the C ABI are passed as zero-element buffers and mapped to NULL here
// (XLA FFI has no optional arguments); scalars are attributes; every handler only enqueues on the stream XLA passes
// (no allocation, no synchronisation: command-buffer / CUDA-graph compatible); scratch comes as an extra result
// sized by the vdn_*_workspace() queries.
#include "vdn.h"

#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define VDN_HAVE_XLA_FFI 1
#endif
#endif

#ifdef VDN_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using Buf = ffi::AnyBuffer;
using Out = ffi::Result<ffi::AnyBuffer>;

namespace {
inline ffi::Error Status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, vdn_last_error());
}
// zero-element buffer == "operand absent"
inline const void* P(const Buf& b) { return b.element_count() == 0 ? nullptr : b.untyped_data(); }
inline void* P(Out& b) { return b->element_count() == 0 ? nullptr : b->untyped_data(); }
inline const float* F(const Buf& b) { return static_cast<const float*>(P(b)); }
inline float* F(Out& b) { return static_cast<float*>(P(b)); }
}  // namespace

#define STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

// ---- tap-GEMM (conv (1,3,3) / 1x1 / Linear / LinearGeneral / Downsample / Upsample, forward and dgrad) ----------
static ffi::Error TapGemm(cudaStream_t st, Buf src0, Buf src1, Buf wp, Buf bias, Buf residual, Buf residual2, Out out,
                          Out out2, Out gn_sums, int32_t kind, int32_t n_img, int32_t H, int32_t W, int32_t n_src,
                          int32_t src_c, ffi::Span<const int32_t> tap_dy, ffi::Span<const int32_t> tap_dx,
                          int32_t n_out, int32_t py, int32_t px, int32_t out_dtype, int32_t split_col,
                          int32_t gn_groups, int32_t rows_per_sample) {
  vdn_tapgemm_desc d = {};
  d.kind = kind; d.n_img = n_img; d.H = H; d.W = W; d.n_src = n_src; d.src_c = src_c;
  d.n_taps = static_cast<int>(tap_dy.size());
  if (d.n_taps > 16 || tap_dx.size() != tap_dy.size()) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "taps");
  for (int i = 0; i < d.n_taps; ++i) { d.tap_dy[i] = tap_dy[i]; d.tap_dx[i] = tap_dx[i]; }
  d.n_out = n_out; d.py = py; d.px = px; d.out_dtype = out_dtype; d.split_col = split_col;
  d.gn_groups = gn_groups; d.rows_per_sample = rows_per_sample;
  return Status(vdn_tapgemm(&d, P(src0), P(src1), P(wp), F(bias), P(residual), P(residual2), P(out), P(out2),
                            F(gn_sums), st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_tapgemm_ffi, TapGemm,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>()
        .Attr<int32_t>("kind").Attr<int32_t>("n_img").Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("n_src")
        .Attr<int32_t>("src_c").Attr<ffi::Span<const int32_t>>("tap_dy").Attr<ffi::Span<const int32_t>>("tap_dx")
        .Attr<int32_t>("n_out").Attr<int32_t>("py").Attr<int32_t>("px").Attr<int32_t>("out_dtype")
        .Attr<int32_t>("split_col").Attr<int32_t>("gn_groups").Attr<int32_t>("rows_per_sample"));

// dw (and dbias) are accumulated in place: pass them as operands aliased to the results (input_output_aliases)
static ffi::Error WGrad(cudaStream_t st, Buf src0, Buf src1, Buf g, Buf dw_in, Buf dbias_in, Out dw, Out dbias,
                        int32_t kind, int32_t n_img, int32_t H, int32_t W, int32_t n_src, int32_t C, int32_t Cout,
                        ffi::Span<const int32_t> tap_dy, ffi::Span<const int32_t> tap_dx) {
  (void)dw_in; (void)dbias_in;
  return Status(vdn_wgrad_bias(kind, P(src0), P(src1), P(g), F(dw), F(dbias), n_img, H, W, n_src, C, Cout,
                               static_cast<int>(tap_dy.size()), tap_dy.begin(), tap_dx.begin(), st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_wgrad_ffi, WGrad,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>()
        .Attr<int32_t>("kind").Attr<int32_t>("n_img").Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("n_src")
        .Attr<int32_t>("C").Attr<int32_t>("Cout").Attr<ffi::Span<const int32_t>>("tap_dy")
        .Attr<ffi::Span<const int32_t>>("tap_dx"));

static ffi::Error PackWeight(cudaStream_t st, Buf src, Out dst, int32_t taps, int32_t cin, int32_t cout, int32_t mode,
                             ffi::Span<const int32_t> perm, int32_t ld, int32_t n_off, int32_t k_off) {
  return Status(vdn_pack_weight(F(src), P(dst), taps, cin, cout, mode, perm.size() ? perm.begin() : nullptr, ld, n_off,
                                k_off, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_pack_weight_ffi, PackWeight,
    STREAM.Arg<Buf>().Ret<Buf>().Attr<int32_t>("taps").Attr<int32_t>("cin").Attr<int32_t>("cout").Attr<int32_t>("mode")
        .Attr<ffi::Span<const int32_t>>("perm").Attr<int32_t>("ld").Attr<int32_t>("n_off").Attr<int32_t>("k_off"));

// ---- Block / ResnetBlock normalisation (modules.py:150-243) --------------------------------------------------------
static ffi::Error GnSiluFwd(cudaStream_t st, Buf x, Buf sums, Buf gamma, Buf beta, Buf ss, Out out, int32_t B,
                            int32_t rows, int32_t C, int32_t G, int32_t ss_ld) {
  return Status(vdn_gn_silu_fwd(P(x), F(sums), F(gamma), F(beta), F(ss), ss_ld, P(out), B, rows, C, G, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_gn_silu_fwd_ffi, GnSiluFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("rows")
        .Attr<int32_t>("C").Attr<int32_t>("G").Attr<int32_t>("ss_ld"));

static ffi::Error TailFwd(cudaStream_t st, Buf b_raw, Buf sums, Buf gamma, Buf beta, Buf s, Buf ln_g, Buf ln_b, Out out,
                          int32_t B, int32_t rows, int32_t C, int32_t G) {
  return Status(vdn_resblock_tail_fwd(P(b_raw), F(sums), F(gamma), F(beta), P(s), F(ln_g), F(ln_b), P(out), B, rows, C, G, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_resblock_tail_fwd_ffi, TailFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("B")
        .Attr<int32_t>("rows").Attr<int32_t>("C").Attr<int32_t>("G"));

// dgamma / dbeta / dss / dconv_bias accumulate in place (alias them to zero-initialised operands)
static ffi::Error GnSiluBwd(cudaStream_t st, Buf dy, Buf x, Buf sums, Buf gamma, Buf beta, Buf ss, Out T_ws, Out dx,
                            Out dgamma, Out dbeta, Out dss, Out dconv_bias, int32_t B, int32_t rows, int32_t C, int32_t G,
                            int32_t ss_ld, int32_t dss_ld) {
  return Status(vdn_gn_silu_bwd(P(dy), P(x), F(sums), F(gamma), F(beta), F(ss), ss_ld, F(T_ws), P(dx), F(dgamma), F(dbeta),
                                F(dss), dss_ld, F(dconv_bias), B, rows, C, G, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_gn_silu_bwd_ffi, GnSiluBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>()
        .Ret<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("rows").Attr<int32_t>("C").Attr<int32_t>("G")
        .Attr<int32_t>("ss_ld").Attr<int32_t>("dss_ld"));

static ffi::Error LnBwd(cudaStream_t st, Buf s, Buf dy, Buf ln_g, Out ds, Out dgamma, Out dbeta, int64_t P_, int32_t C) {
  return Status(vdn_ln_bwd(P(s), P(dy), F(ln_g), P(ds), F(dgamma), F(dbeta), static_cast<long>(P_), C, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_ln_bwd_ffi, LnBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int64_t>("P").Attr<int32_t>("C"));

// ---- attention (modules.py:64-129, :247-326) ---------------------------------------------------------------------------
static ffi::Error MhaCoreFwd(cudaStream_t st, Buf qkv, Out o, Out lse, int32_t mode, int32_t B, int32_t F_, int32_t HW) {
  return Status(vdn_mha_core_fwd(P(qkv), P(o), F(lse), mode, B, F_, HW, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_core_fwd_ffi, MhaCoreFwd,
    STREAM.Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("mode").Attr<int32_t>("B").Attr<int32_t>("F").Attr<int32_t>("HW"));

static ffi::Error MhaCoreBwd(cudaStream_t st, Buf qkv, Buf o, Buf d_o, Buf lse, Out D_ws, Out dqkv, int32_t mode,
                             int32_t B, int32_t F_, int32_t HW) {
  return Status(vdn_mha_core_bwd(P(qkv), P(o), P(d_o), F(lse), F(D_ws), P(dqkv), mode, B, F_, HW, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_core_bwd_ffi, MhaCoreBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("mode").Attr<int32_t>("B")
        .Attr<int32_t>("F").Attr<int32_t>("HW"));

static ffi::Error MhaTemporalFwd(cudaStream_t st, Buf x, Buf w_hm, Buf bias_hm, Out o, Out qkv, Out lse, int32_t B,
                                 int32_t F_, int32_t H, int32_t W, int32_t C) {
  const int rc = vdn_mha_temporal_tc_supported(F_, C)
                     ? vdn_mha_temporal_tc_fwd(P(x), P(w_hm), F(bias_hm), P(o), P(qkv), F(lse), B, F_, H, W, C, st)
                     : vdn_mha_temporal_fused_fwd(P(x), P(w_hm), F(bias_hm), P(o), P(qkv), F(lse), B, F_, H, W, C, st);
  return Status(rc);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_temporal_fwd_ffi, MhaTemporalFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("F")
        .Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("C"));

static ffi::Error MhaTemporalCoreFwd(cudaStream_t st, Buf qkv, Out o, Out lse, int32_t B, int32_t F_, int32_t H, int32_t W) {
  return Status(vdn_mha_temporal_core_fwd(P(qkv), P(o), F(lse), B, F_, H, W, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_temporal_core_fwd_ffi, MhaTemporalCoreFwd,
    STREAM.Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("F").Attr<int32_t>("H").Attr<int32_t>("W"));

static ffi::Error MhaTemporalBwd(cudaStream_t st, Buf qkv, Buf d_o, Buf lse, Out dqkv, Out dbias, int32_t B, int32_t F_,
                                 int32_t H, int32_t W) {
  return Status(vdn_mha_temporal_tc_bwd(P(qkv), P(d_o), F(lse), P(dqkv), F(dbias), B, F_, H, W, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_temporal_bwd_ffi, MhaTemporalBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("F").Attr<int32_t>("H")
        .Attr<int32_t>("W"));

static ffi::Error MhaFoldPack(cudaStream_t st, Buf w_qkv, Buf b_qkv, Buf w_out, Buf b_out, Out fa, Out fu, Out fm, Out fb) {
  return Status(vdn_mha_fold_pack(F(w_qkv), F(b_qkv), F(w_out), F(b_out), P(fa), F(fu), P(fm), F(fb), st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_fold_pack_ffi, MhaFoldPack,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>());

static ffi::Error MhaFoldedFwd(cudaStream_t st, Buf x, Buf fa, Buf fu, Buf fm, Buf fb, Out out, int32_t B, int32_t F_,
                               int32_t H, int32_t W, int32_t C) {
  return Status(vdn_mha_temporal_folded_fwd(P(x), P(fa), F(fu), P(fm), F(fb), P(out), B, F_, H, W, C, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_temporal_folded_fwd_ffi, MhaFoldedFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("F")
        .Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("C"));

static ffi::Error QkvHeadMajorPack(cudaStream_t st, Buf w, Buf bias, Out dst, Out bias_dst, int32_t C) {
  return Status(vdn_qkv_headmajor_pack(F(w), F(bias), P(dst), F(bias_dst), C, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_qkv_headmajor_pack_ffi, QkvHeadMajorPack,
    STREAM.Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("C"));

// MultiheadAttention with the optional post-softmax mask / bias (modules.py:291-321)
static ffi::Error MhaCoreExtFwd(cudaStream_t st, Buf qkv, Buf mask_b, Buf pos_bias, Out o, int32_t dtype, int32_t heads,
                                int32_t dim, int32_t n_seq, int32_t S, int32_t inner, int32_t seqs_per_batch,
                                int32_t copy_v) {
  return Status(vdn_mha_core_ext_fwd(P(qkv), P(o), dtype, heads, dim, n_seq, S, inner,
                                     static_cast<const unsigned char*>(P(mask_b)), seqs_per_batch, F(pos_bias), copy_v, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_mha_core_ext_fwd_ffi, MhaCoreExtFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("dtype").Attr<int32_t>("heads").Attr<int32_t>("dim")
        .Attr<int32_t>("n_seq").Attr<int32_t>("S").Attr<int32_t>("inner").Attr<int32_t>("seqs_per_batch")
        .Attr<int32_t>("copy_v"));

static ffi::Error RelPosBias(cudaStream_t st, Buf embedding, Out out, Out buckets, int32_t n, int32_t heads) {
  return Status(vdn_rel_pos_bias(F(embedding), n, heads, F(out), static_cast<int*>(P(buckets)), st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_rel_pos_bias_ffi, RelPosBias,
    STREAM.Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n").Attr<int32_t>("heads"));

static ffi::Error SlaCoreFwd(cudaStream_t st, Buf qkv, Out tok, Out ctx, Out kstat, Out ws, int32_t n_img, int32_t N) {
  return Status(vdn_sla_core_fwd(P(qkv), P(tok), F(ctx), F(kstat), F(ws), n_img, N, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_sla_core_fwd_ffi, SlaCoreFwd,
    STREAM.Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_img").Attr<int32_t>("N"));

static ffi::Error SlaCoreBwd(cudaStream_t st, Buf qkv, Buf d_tok, Buf ctx, Buf kstat, Out dctx, Out dqkv, int32_t n_img,
                             int32_t N) {
  return Status(vdn_sla_core_bwd(P(qkv), P(d_tok), F(ctx), F(kstat), F(dctx), P(dqkv), n_img, N, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_sla_core_bwd_ffi, SlaCoreBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_img").Attr<int32_t>("N"));

static ffi::Error SlaFusedFwd(cudaStream_t st, Buf x, Buf w_qkv, Buf w_out, Out out, Out ctx, Out kstat, Out ws,
                              int32_t n_img, int32_t N, int32_t C) {
  return Status(vdn_sla_fused_fwd(P(x), P(w_qkv), P(w_out), P(out), F(ctx), F(kstat), F(ws), n_img, N, C, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_sla_fused_fwd_ffi, SlaFusedFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_img")
        .Attr<int32_t>("N").Attr<int32_t>("C"));

// ---- small layers (unet3d.py:110-133,251; modules.py:30-45,202-208) ----------------------------------------------------
static ffi::Error InitConvFwd(cudaStream_t st, Buf x, Buf w, Buf bias, Out out, int32_t B, int32_t Cin, int32_t F_, int32_t H,
                              int32_t W, int32_t Cout, int32_t ks) {
  return Status(vdn_init_conv_fwd(F(x), F(w), F(bias), P(out), B, Cin, F_, H, W, Cout, ks, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_init_conv_fwd_ffi, InitConvFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("Cin").Attr<int32_t>("F")
        .Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("Cout").Attr<int32_t>("ks"));

static ffi::Error InitConvWgrad(cudaStream_t st, Buf x, Buf dy, Out dw, Out dbias, int32_t B, int32_t Cin, int32_t F_,
                                int32_t H, int32_t W, int32_t Cout, int32_t ks) {
  return Status(vdn_init_conv_wgrad(F(x), P(dy), F(dw), F(dbias), B, Cin, F_, H, W, Cout, ks, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_init_conv_wgrad_ffi, InitConvWgrad,
    STREAM.Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("Cin").Attr<int32_t>("F")
        .Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("Cout").Attr<int32_t>("ks"));

static ffi::Error FinalConvFwd(cudaStream_t st, Buf h, Buf w, Buf bias, Out out, int64_t P_, int32_t C, int32_t Co) {
  return Status(vdn_final_conv_fwd(P(h), F(w), F(bias), F(out), static_cast<long>(P_), C, Co, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_final_conv_fwd_ffi, FinalConvFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int64_t>("P").Attr<int32_t>("C").Attr<int32_t>("Co"));

static ffi::Error FinalConvBwd(cudaStream_t st, Buf h, Buf dout, Buf w, Out dh, Out dw, Out db, int64_t P_, int32_t C,
                               int32_t Co) {
  return Status(vdn_final_conv_bwd(P(h), F(dout), F(w), P(dh), F(dw), F(db), static_cast<long>(P_), C, Co, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_final_conv_bwd_ffi, FinalConvBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int64_t>("P").Attr<int32_t>("C")
        .Attr<int32_t>("Co"));

static ffi::Error TimeMlpFwd(cudaStream_t st, Buf time, Buf w1, Buf b1, Buf w2, Buf b2, Out emb, Out h1, Out t, int32_t B,
                             int32_t dim) {
  return Status(vdn_time_mlp_fwd(static_cast<const int*>(P(time)), F(w1), F(b1), F(w2), F(b2), F(emb), F(h1), F(t), B, dim, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_time_mlp_fwd_ffi, TimeMlpFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("B")
        .Attr<int32_t>("dim"));

static ffi::Error TimeMlpBwd(cudaStream_t st, Buf dt, Buf emb, Buf h1, Buf w2, Out dw1, Out db1, Out dw2, Out db2,
                             Out dh1_ws, int32_t B, int32_t dim) {
  return Status(vdn_time_mlp_bwd(F(dt), F(emb), F(h1), F(w2), F(dw1), F(db1), F(dw2), F(db2), F(dh1_ws), B, dim, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_time_mlp_bwd_ffi, TimeMlpBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>()
        .Attr<int32_t>("B").Attr<int32_t>("dim"));

// heads_dev: a device byte buffer holding the vdn_time_head table (pointers into the parameter buffers)
static ffi::Error TimeHeadsFwd(cudaStream_t st, Buf t, Buf heads_dev, Out e_pre, Out ss, int32_t n_heads, int32_t ss_ld,
                               int32_t B, int32_t td) {
  return Status(vdn_time_heads_fwd(F(t), P(heads_dev), n_heads, F(e_pre), F(ss), ss_ld, B, td, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_time_heads_fwd_ffi, TimeHeadsFwd,
    STREAM.Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_heads").Attr<int32_t>("ss_ld").Attr<int32_t>("B")
        .Attr<int32_t>("td"));

static ffi::Error TimeHeadsBwd(cudaStream_t st, Buf t, Buf heads_dev, Buf e_pre, Buf dss, Out de_ws, Out dt, int32_t n_heads,
                               int32_t ss_ld, int32_t B, int32_t td) {
  return Status(vdn_time_heads_bwd(F(t), P(heads_dev), n_heads, F(e_pre), F(dss), ss_ld, F(de_ws), F(dt), B, td, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_time_heads_bwd_ffi, TimeHeadsBwd,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_heads").Attr<int32_t>("ss_ld")
        .Attr<int32_t>("B").Attr<int32_t>("td"));

// ---- diffusion math (gaussian_diffusion.py:120-261,401-470) --------------------------------------------------------------
static ffi::Error QSample(cudaStream_t st, Buf x, Buf noise, Buf t, Buf sqrt_ac, Buf sqrt_1mac, Out out, int32_t B,
                          int64_t per_sample, int32_t normalize) {
  return Status(vdn_q_sample(F(x), F(noise), static_cast<const int*>(P(t)), F(sqrt_ac), F(sqrt_1mac), F(out), B,
                             static_cast<long>(per_sample), normalize, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_q_sample_ffi, QSample,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int64_t>("per_sample")
        .Attr<int32_t>("normalize"));

static ffi::Error Loss(cudaStream_t st, Buf pred, Buf noise, Out loss, Out dpred, int32_t B, int32_t C, int64_t FHW, int32_t l1) {
  return Status(vdn_loss(F(pred), F(noise), F(loss), F(dpred), B, C, static_cast<long>(FHW), l1, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_loss_ffi, Loss,
    STREAM.Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("B").Attr<int32_t>("C").Attr<int64_t>("FHW")
        .Attr<int32_t>("l1"));

static ffi::Error PSample(cudaStream_t st, Buf x, Buf eps, Buf z, Buf t, Buf recip, Buf recipm1, Buf coef1, Buf coef2,
                          Buf logvar, Out out, int32_t B, int32_t C, int64_t FHW, int32_t clip) {
  return Status(vdn_p_sample(F(x), F(eps), F(z), static_cast<const int*>(P(t)), F(recip), F(recipm1), F(coef1), F(coef2),
                             F(logvar), F(out), B, C, static_cast<long>(FHW), clip, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_p_sample_ffi, PSample,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>()
        .Attr<int32_t>("B").Attr<int32_t>("C").Attr<int64_t>("FHW").Attr<int32_t>("clip"));

static ffi::Error Randn(cudaStream_t st, Out out, int64_t seed, int64_t subseq, int64_t elem_offset) {
  return Status(vdn_randn(F(out), static_cast<long>(out->element_count()), static_cast<unsigned long long>(seed),
                          static_cast<unsigned long long>(subseq), static_cast<unsigned long long>(elem_offset), st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_randn_ffi, Randn,
    STREAM.Ret<Buf>().Attr<int64_t>("seed").Attr<int64_t>("subseq").Attr<int64_t>("elem_offset"));

// ---- training glue (trainer.py:367-382, utils.py:127-152) -------------------------------------------------------------------
// p / m / v / ema are updated in place: alias them to the results (input_output_aliases={0:0, 2:1, 3:2, 4:3})
static ffi::Error AdamEma(cudaStream_t st, Buf p_in, Buf g, Buf m_in, Buf v_in, Buf ema_in, Buf hp, Buf sqnorm, Out p, Out m,
                          Out v, Out ema) {
  (void)p_in; (void)m_in; (void)v_in; (void)ema_in;
  const long n = static_cast<long>(p->element_count());
  return Status(P(sqnorm) ? vdn_adam_ema_clip(F(p), F(g), F(m), F(v), F(ema), F(hp), F(sqnorm), n, st)
                          : vdn_adam_ema(F(p), F(g), F(m), F(v), F(ema), F(hp), n, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_adam_ema_ffi, AdamEma,
    STREAM.Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>()
        .Ret<Buf>());

static ffi::Error GradSqnorm(cudaStream_t st, Buf g, Out out) {
  return Status(vdn_grad_sqnorm(F(g), static_cast<long>(g.element_count()), F(out), st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_grad_sqnorm_ffi, GradSqnorm, STREAM.Arg<Buf>().Ret<Buf>());

static ffi::Error Colsum(cudaStream_t st, Buf dy, Buf db_in, Out db, int64_t P_, int32_t C) {
  (void)db_in;
  return Status(vdn_colsum(P(dy), F(db), static_cast<long>(P_), C, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_colsum_ffi, Colsum, STREAM.Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int64_t>("P").Attr<int32_t>("C"));

static ffi::Error AddBf16(cudaStream_t st, Buf a, Buf b, Out out) {
  return Status(vdn_add_bf16(P(a), P(b), P(out), static_cast<long>(a.element_count()), st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_add_bf16_ffi, AddBf16, STREAM.Arg<Buf>().Arg<Buf>().Ret<Buf>());

// gradient exchange: `comm` is the handle vdn_comm_init returned (created once per process on the Python side and passed
// as an int64 attribute); the bucket is reduced in place (alias operand 0 to result 0)
static ffi::Error AllreduceBucket(cudaStream_t st, Buf buf_in, Out buf, int64_t comm, int32_t dtype) {
  (void)buf_in;
  return Status(vdn_allreduce_bucket(reinterpret_cast<void*>(comm), P(buf), static_cast<long>(buf->element_count()), dtype, st));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(vdn_allreduce_bucket_ffi, AllreduceBucket,
    STREAM.Arg<Buf>().Ret<Buf>().Attr<int64_t>("comm").Attr<int32_t>("dtype"));

extern "C" int vdn_ffi_available(void) { return 1; }

#else  // ---------------------------------------------------------------------------------------------------------------
// No XLA FFI headers on this machine (the development image has no jaxlib): the translation unit still compiles, so
// that the build and the symbol check run everywhere; jax_ffi.register() refuses to proceed when this returns 0.
extern "C" int vdn_ffi_available(void) { return 0; }
#endif

// names of the FFI targets this file defines (NULL-terminated); each is exported as the symbol "<name>_ffi"
extern "C" const char* const* vdn_ffi_targets(void) {
  static const char* const names[] = {
      "vdn_tapgemm", "vdn_wgrad", "vdn_pack_weight", "vdn_gn_silu_fwd", "vdn_resblock_tail_fwd", "vdn_gn_silu_bwd",
      "vdn_ln_bwd", "vdn_mha_core_fwd", "vdn_mha_core_bwd", "vdn_mha_temporal_fwd", "vdn_mha_temporal_core_fwd",
      "vdn_mha_temporal_bwd", "vdn_mha_fold_pack", "vdn_mha_temporal_folded_fwd", "vdn_qkv_headmajor_pack",
      "vdn_mha_core_ext_fwd", "vdn_rel_pos_bias", "vdn_sla_core_fwd", "vdn_sla_core_bwd", "vdn_sla_fused_fwd",
      "vdn_init_conv_fwd", "vdn_init_conv_wgrad", "vdn_final_conv_fwd", "vdn_final_conv_bwd", "vdn_time_mlp_fwd",
      "vdn_time_mlp_bwd", "vdn_time_heads_fwd", "vdn_time_heads_bwd", "vdn_q_sample", "vdn_loss", "vdn_p_sample",
      "vdn_randn", "vdn_adam_ema", "vdn_grad_sqnorm", "vdn_colsum", "vdn_add_bf16", "vdn_allreduce_bucket", nullptr};
  return names;
}
