"""GPU parity of the non-GEMM kernels and of the tcgen05 wgrad kernel against the oracle's layer
semantics (oracle/unet3d_oracle.py) evaluated with torch autograd in fp32 on the same bf16-rounded
operands. Tolerances are for bf16 storage of activations (stated per test)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)


def _bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(torch.bfloat16)


def _replicated(sums):
    """[B][G][2] -> the kernels' [R][B][G][2] replica layout (all mass in replica 3)."""
    from video_diffusion_nnx_b200 import ops

    out = torch.zeros((ops.GN_REPLICAS,) + tuple(sums.shape), device=sums.device)
    out[3] = sums
    return out.contiguous()


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()


# ------------------------------------------------------------------------------------------
# wgrad (tcgen05, MN-major operands) vs reference kernel and torch
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "n_img,H,W,c,cout,n_src",
    [(2, 16, 16, 32, 32, 1), (2, 32, 32, 64, 128, 1), (3, 8, 8, 128, 64, 2), (1, 16, 16, 16, 48, 1),
     (2, 64, 64, 32, 32, 2), (1, 16, 16, 256, 256, 1)],
)
def test_wgrad_unit(n_img, H, W, c, cout, n_src):
    from video_diffusion_nnx_b200 import ops

    _setup()
    srcs = [_bf(n_img, H, W, c) for _ in range(n_src)]
    g = _bf(n_img, H, W, cout)
    for taps, k, pad in ((ops.TAPS_3x3, 3, 1), (ops.TAPS_1x1, 1, 0)):
        x = torch.cat([s.float() for s in srcs], -1).permute(0, 3, 1, 2).requires_grad_(False)
        w = torch.zeros(cout, n_src * c, k, k, device=DEV, requires_grad=True)
        y = F.conv2d(x, w, padding=pad)
        y.backward(g.float().permute(0, 3, 1, 2))
        ref = w.grad.permute(2, 3, 1, 0).reshape(len(taps), n_src * c, cout).contiguous()  # [taps][cin][cout]
        dw_ref = torch.zeros_like(ref)
        ops.wgrad(ops.VDN_TAP_UNIT, srcs, g, dw_ref, taps, ref=True)
        assert _rel(dw_ref, ref) < 1e-4
        dw = torch.zeros_like(ref)
        ops.wgrad(ops.VDN_TAP_UNIT, srcs, g, dw, taps)
        torch.cuda.synchronize()
        assert _rel(dw, ref) < 1e-4, f"taps={len(taps)}"
        # with the fused bias gradient (ones atom in a spare M slot, or the colsum fallback)
        dw2, db = torch.zeros_like(ref), torch.full((cout,), 0.5, device=DEV)
        ops.wgrad(ops.VDN_TAP_UNIT, srcs, g, dw2, taps, dbias=db)
        torch.cuda.synchronize()
        assert _rel(dw2, ref) < 1e-4
        assert _rel(db - 0.5, g.float().sum(dim=(0, 1, 2))) < 1e-4, f"bias taps={len(taps)}"


@pytest.mark.parametrize("n_img,H,W,c", [(2, 32, 32, 32), (2, 16, 16, 128)])
def test_wgrad_down_and_up(n_img, H, W, c):
    from video_diffusion_nnx_b200 import ops

    _setup()
    # strided conv k4 s2: x (2H x 2W) -> y (H x W)
    x = _bf(n_img, 2 * H, 2 * W, c)
    g = _bf(n_img, H, W, c)
    w = torch.zeros(c, c, 4, 4, device=DEV, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=2, padding=1)
    y.backward(g.float().permute(0, 3, 1, 2))
    ref = w.grad.permute(2, 3, 1, 0).reshape(16, c, c).contiguous()
    dw, db = torch.zeros_like(ref), torch.zeros(c, device=DEV)
    ops.wgrad(ops.VDN_TAP_DOWN, [x], g, dw, ops.TAPS_4x4, dbias=db)
    assert _rel(dw, ref) < 1e-4
    assert _rel(db, g.float().sum(dim=(0, 1, 2))) < 1e-4
    # transposed conv (flax, unflipped kernel): x (H x W) -> y (2H x 2W)
    x2 = _bf(n_img, H, W, c)
    g2 = _bf(n_img, 2 * H, 2 * W, c)
    wk = torch.zeros(4, 4, c, c, device=DEV, requires_grad=True)  # flax layout (kh,kw,in,out)
    wt = wk.flip(0, 1).permute(2, 3, 0, 1)
    y2 = F.conv_transpose2d(x2.float().permute(0, 3, 1, 2), wt, stride=2, padding=1)
    y2.backward(g2.float().permute(0, 3, 1, 2))
    ref2 = wk.grad.reshape(16, c, c).contiguous()
    dw2, db2 = torch.zeros_like(ref2), torch.zeros(c, device=DEV)
    ops.wgrad(ops.VDN_TAP_UP, [x2], g2, dw2, ops.TAPS_4x4, dbias=db2)
    assert _rel(dw2, ref2) < 1e-4
    assert _rel(db2, g2.float().sum(dim=(0, 1, 2))) < 1e-4
    dw2r = torch.zeros_like(ref2)
    ops.wgrad(ops.VDN_TAP_UP, [x2], g2, dw2r, ops.TAPS_4x4, ref=True)
    assert _rel(dw2r, ref2) < 1e-4


# ------------------------------------------------------------------------------------------
# GroupNorm + scale/shift + SiLU, ResnetBlock tail, their backward
# ------------------------------------------------------------------------------------------
def _gn_ref(x, gamma, beta, G=8, eps=1e-6):
    B, R, Cc = x.shape
    g = x.reshape(B, R, G, Cc // G)
    mean = g.mean(dim=(1, 3), keepdim=True)
    var = ((g * g).mean(dim=(1, 3), keepdim=True) - mean * mean).clamp_min(0)
    return ((g - mean) * torch.rsqrt(var + eps)).reshape(B, R, Cc) * gamma + beta


def _ln_ref(x, gamma, beta, eps=1e-6):
    mean = x.mean(-1, keepdim=True)
    var = ((x * x).mean(-1, keepdim=True) - mean * mean).clamp_min(0)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


@pytest.fixture(params=["two-kernel", "fused", "large-sample"])
def gn_bwd_mode(request):
    """gn_silu_bwd variants: the default two-kernel version, the opt-in single-launch version with the per-sample
    grid barrier (VDN_GN_FUSED=1), and - through a larger sample - the multi-chunk blocks of the two-kernel version."""
    from video_diffusion_nnx_b200 import _lib

    if request.param == "fused":
        _lib.debug_set("VDN_GN_FUSED", 1)
    yield request.param
    _lib.debug_clear("VDN_GN_FUSED")


@pytest.mark.parametrize("B,R,Cc,with_ss", [(2, 512, 32, True), (3, 160, 64, False), (2, 96, 256, True), (1, 64, 1024, True)])
def test_gn_silu_fwd_bwd(gn_bwd_mode, B, R, Cc, with_ss):
    from video_diffusion_nnx_b200 import ops

    _setup()
    if gn_bwd_mode == "large-sample":
        R *= 2048 if Cc <= 64 else 1024
    x = _bf(B, R, Cc, scale=2.0)
    gamma = (1 + 0.2 * torch.randn(Cc, device=DEV)).requires_grad_(True)
    beta = (0.2 * torch.randn(Cc, device=DEV)).requires_grad_(True)
    ss = (0.3 * torch.randn(B, 2 * Cc, device=DEV)).requires_grad_(True) if with_ss else None
    xf = x.float().requires_grad_(True)
    y = _gn_ref(xf, gamma, beta)
    if with_ss:
        y = y * (ss[:, None, :Cc] + 1) + ss[:, None, Cc:]
    y = F.silu(y)
    dy = _bf(B, R, Cc)
    y.backward(dy.float())
    xg = xf.detach().reshape(B, R, 8, Cc // 8)
    sums = _replicated(torch.stack([xg.sum(dim=(1, 3)), (xg * xg).sum(dim=(1, 3))], -1))
    out = torch.empty_like(x)
    ops.gn_silu_fwd(x, sums, gamma.detach(), beta.detach(), ss.detach() if with_ss else None, out, B, R, Cc)
    assert _rel(out, y) < 1e-2  # bf16 output rounding
    T = torch.empty(B, Cc, 2, device=DEV)
    dx = torch.empty_like(x)
    dgam, dbet = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    dss = torch.zeros(B, 2 * Cc, device=DEV) if with_ss else None
    dcb = torch.zeros(Cc, device=DEV)
    ops.gn_silu_bwd(dy, x, sums, gamma.detach(), beta.detach(), ss.detach() if with_ss else None, T, dx, dgam, dbet,
                    dss, B, R, Cc, dconv_bias=dcb)
    assert _rel(dx, xf.grad) < 2e-2
    assert _rel(dcb, xf.grad.sum(dim=(0, 1))) < 2e-2 or dcb.abs().max() < 1e-2 * xf.grad.abs().sum(dim=(0, 1)).max()
    assert _rel(dgam, gamma.grad) < 1e-3
    assert _rel(dbet, beta.grad) < 1e-3
    if with_ss:
        assert _rel(dss, ss.grad) < 1e-3
    # the caller-zeroed variant (vdn_gn_silu_bwd_acc: the engine zeroes one scratch per step instead of a memset node per
    # layer) computes the same thing when its T_ws slice is zero on entry
    T0 = torch.zeros(B, Cc, 2, device=DEV)
    dx0 = torch.empty_like(x)
    dgam0, dbet0, dcb0 = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    dss0 = torch.zeros(B, 2 * Cc, device=DEV) if with_ss else None
    ops.gn_silu_bwd(dy, x, sums, gamma.detach(), beta.detach(), ss.detach() if with_ss else None, T0, dx0, dgam0, dbet0,
                    dss0, B, R, Cc, dconv_bias=dcb0, prezeroed=True)
    assert _rel(dx0, dx) < 1e-2 and _rel(dgam0, dgam) < 1e-3 and _rel(dbet0, dbet) < 1e-3
    assert _rel(T0, T) < 1e-3  # same sums (atomic order may differ)


@pytest.mark.parametrize("B,R,Cc", [(2, 256, 32), (2, 100, 128), (1, 40, 512), (1, 24, 1024),
                                    (2, 655360, 32), (1, 320000, 128)])  # large: several chunks per block in ln_bwd
def test_resblock_tail_and_ln_bwd(B, R, Cc):
    from video_diffusion_nnx_b200 import ops

    _setup()
    b_raw = _bf(B, R, Cc, scale=1.5)
    s = _bf(B, R, Cc)
    gamma = 1 + 0.2 * torch.randn(Cc, device=DEV)
    beta = 0.2 * torch.randn(Cc, device=DEV)
    lg = (1 + 0.2 * torch.randn(Cc, device=DEV)).requires_grad_(True)
    lb = (0.2 * torch.randn(Cc, device=DEV)).requires_grad_(True)
    sf = s.float().requires_grad_(True)
    ln = _ln_ref(sf, lg, lb)
    ref = F.silu(_gn_ref(b_raw.float(), gamma, beta)) + ln
    xg = b_raw.float().reshape(B, R, 8, Cc // 8)
    sums = _replicated(torch.stack([xg.sum(dim=(1, 3)), (xg * xg).sum(dim=(1, 3))], -1))
    out = torch.empty_like(s)
    ops.resblock_tail_fwd(b_raw, sums, gamma, beta, s, lg.detach(), lb.detach(), out, B, R, Cc)
    assert _rel(out, ref) < 1e-2
    dy = _bf(B, R, Cc)
    ln.backward(dy.float())
    ds = torch.empty_like(s)
    dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    ops.ln_bwd(s, dy, lg.detach(), ds, dg, db, B * R, Cc)
    assert _rel(ds, sf.grad) < 2e-2
    big = B * R > 100000  # fp32 atomics over ~1e6 pixels against torch's own fp32 reduction order
    assert _rel(dg, lg.grad) < (3e-3 if big else 1e-3)
    assert _rel(db, lb.grad) < (3e-3 if big else 1e-3)


# ------------------------------------------------------------------------------------------
# attention cores
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,B,Fr,HW", [(0, 2, 10, 64), (0, 1, 16, 256), (1, 2, 3, 64), (1, 1, 2, 256), (0, 2, 2, 4096),
                                          (0, 1, 10, 4096)])
def test_mha_core(mode, B, Fr, HW):
    from video_diffusion_nnx_b200 import ops

    _setup()
    P = B * Fr * HW
    qkv = _bf(P, 768)
    qf = qkv.float().requires_grad_(True)
    t = qf.reshape(B, Fr, HW, 3, 8, 32)
    if mode == 0:  # sequences over frames: (B, HW, F, ...)
        t = t.permute(0, 2, 1, 3, 4, 5)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]  # (..., S, 8, 32)
    att = torch.einsum("...ihd,...jhd->...hij", q / math.sqrt(32), k).softmax(-1)
    o = torch.einsum("...hij,...jhd->...ihd", att, v)
    if mode == 0:
        o = o.permute(0, 2, 1, 3, 4)
    o_ref = o.reshape(P, 256)
    do = _bf(P, 256)
    o_ref.backward(do.float())
    out = torch.empty(P, 256, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(P, 8, device=DEV)
    ops.mha_core_fwd(qkv, out, lse, mode, B, Fr, HW)
    assert _rel(out, o_ref) < 1e-2
    dqkv = torch.empty_like(qkv)
    Dws = torch.empty(P, 8, device=DEV)
    ops.mha_core_bwd(qkv, out, do, lse, Dws, dqkv, mode, B, Fr, HW)
    assert _rel(dqkv, qf.grad) < 2e-2
    if mode == 0 and Fr <= 16:  # single-kernel temporal backward (smem exchange)
        side = int(round(HW ** 0.5))
        dq2 = torch.zeros_like(qkv)
        ops.mha_temporal_bwd(qkv, out, do, lse, dq2, B, Fr, side, side)
        assert _rel(dq2, qf.grad) < 2e-2
        if Fr <= 16:  # tensor-core backward (warp-level MMAs, any F <= 16)
            dq3 = torch.zeros_like(qkv)
            dbias = torch.zeros(768, device=DEV)
            ops.mha_temporal_tc_bwd(qkv, do, lse, dq3, B, Fr, side, side, dbias=dbias)
            torch.cuda.synchronize()
            assert _rel(dbias, qf.grad.sum(0)) < 2e-2   # fused gradient of the q|k|v projection bias
            for part, name in enumerate("qkv"):
                e = _rel(dq3[:, part * 256:(part + 1) * 256], qf.grad[:, part * 256:(part + 1) * 256])
                assert e < 3e-2, (name, e)   # P and dS are rounded to bf16 before the MMAs


@pytest.mark.parametrize("n_img,N", [(3, 64), (2, 256), (2, 1024), (1, 4096), (2, 100)])
def test_sla_core(n_img, N):
    from video_diffusion_nnx_b200 import ops

    _setup()
    P = n_img * N
    qkv = _bf(P, 768)
    qf = qkv.float().requires_grad_(True)
    t = qf.reshape(n_img, N, 3, 8, 32)
    q = t[:, :, 0].softmax(-1)       # over the 32 features; NOT scaled (modules.py:107-108,118)
    k = t[:, :, 1].softmax(1)        # over the N tokens
    v = t[:, :, 2]
    ctx = torch.einsum("bnhd,bnhe->bhde", k, v)
    o_ref = torch.einsum("bhde,bnhd->bnhe", ctx, q).reshape(P, 256)
    do = _bf(P, 256)
    o_ref.backward(do.float())
    out = torch.empty(P, 256, dtype=torch.bfloat16, device=DEV)
    ctx_b = torch.empty(n_img, 8, 32, 32, device=DEV)
    kstat = torch.empty(n_img, 8, 2, 32, device=DEV)
    ws = torch.empty(ops.sla_workspace_floats(n_img, N), device=DEV)
    ops.sla_core_fwd(qkv, out, ctx_b, kstat, ws, n_img, N)
    # exp(k - m) is rounded to bf16 before the tensor-core product: expected relative error ~2^-9 / sqrt(3) = 1.1e-3
    assert _rel(ctx_b, ctx) < 3e-3
    assert _rel(out, o_ref) < 1e-2
    dctx = torch.empty(n_img, 8, 32, 32, device=DEV)
    dqkv = torch.empty_like(qkv)
    ops.sla_core_bwd(qkv, do, ctx_b, kstat, dctx, dqkv, n_img, N)
    assert _rel(dqkv, qf.grad) < 2e-2


@pytest.mark.parametrize("n_img,H,W", [(3, 8, 8), (2, 16, 16), (1, 64, 64), (2, 10, 10)])
def test_sla_fused_fwd(n_img, H, W):
    """Fused SpatialLinearAttention block forward (x -> out, C = 32) against torch fp32 on the same bf16 operands
    and against the unfused path (projection GEMM + core + to_out GEMM)."""
    from video_diffusion_nnx_b200 import ops

    _setup()
    Cc, N = 32, H * W
    P = n_img * N
    x = _bf(n_img, H, W, Cc)
    wq = _bf(Cc, 768, scale=Cc ** -0.5)      # fused q|k|v kernel [C][768]
    wo = _bf(256, Cc, scale=256 ** -0.5)     # to_out kernel [256][C]
    w_qkv = torch.empty(768, Cc, dtype=torch.bfloat16, device=DEV)
    w_out = torch.empty(Cc, 256, dtype=torch.bfloat16, device=DEV)
    ops.pack_weight(wq.float().reshape(1, Cc, 768).contiguous(), w_qkv, 1, Cc, 768, 0)
    ops.pack_weight(wo.float().reshape(1, 256, Cc).contiguous(), w_out, 1, 256, Cc, 0)
    qkv = (x.float().reshape(P, Cc) @ wq.float()).to(torch.bfloat16).float()
    t = qkv.reshape(n_img, N, 3, 8, 32)
    q = t[:, :, 0].softmax(-1)
    k = t[:, :, 1].softmax(1)
    ctx_ref = torch.einsum("bnhd,bnhe->bhde", k, t[:, :, 2])
    tok = torch.einsum("bhde,bnhd->bnhe", ctx_ref, q).reshape(P, 256)
    ref = tok @ wo.float() + x.float().reshape(P, Cc)
    out = torch.empty(n_img, H, W, Cc, dtype=torch.bfloat16, device=DEV)
    ctx = torch.empty(n_img, 8, 32, 32, device=DEV)
    kstat = torch.empty(n_img, 8, 2, 32, device=DEV)
    ws = torch.empty(ops.sla_workspace_floats(n_img, N), device=DEV)
    ops.sla_fused_fwd(x, w_qkv, w_out, out, ctx, kstat, ws, n_img, N, Cc)
    torch.cuda.synchronize()
    assert _rel(ctx, ctx_ref) < 3e-3
    assert _rel(out.reshape(P, Cc), ref) < 1e-2


@pytest.mark.parametrize("B,Fr,H,W", [(2, 10, 8, 8), (1, 16, 4, 8), (2, 2, 16, 16), (1, 7, 5, 3)])
def test_mha_temporal_folded_fwd(B, Fr, H, W):
    """Folded inference temporal-attention block (x -> x + out_proj(MHA(x)), C = 32) vs torch fp32."""
    from video_diffusion_nnx_b200 import ops

    _setup()
    Cc = 32
    P = B * Fr * H * W
    x = _bf(B, Fr, H, W, Cc)
    wqkv = torch.randn(Cc, 768, device=DEV) / Cc ** 0.5
    bqkv = 0.1 * torch.randn(768, device=DEV)
    wo = torch.randn(256, Cc, device=DEV) / 16.0
    bo = 0.1 * torch.randn(Cc, device=DEV)
    fa, fm = (torch.empty(8, 32, 32, dtype=torch.bfloat16, device=DEV) for _ in range(2))
    fu, fb = torch.empty(8, 32, device=DEV), torch.empty(32, device=DEV)
    ops.mha_fold_pack(wqkv, bqkv, wo, bo, fa, fu, fm, fb)
    out = torch.empty(B, Fr, H, W, Cc, dtype=torch.bfloat16, device=DEV)
    ops.mha_temporal_folded_fwd(x, fa, fu, fm, fb, out, B, Fr, H, W, Cc)
    torch.cuda.synchronize()
    xf = x.float().reshape(P, Cc)
    t = (xf @ wqkv + bqkv).reshape(B, Fr, H * W, 3, 8, 32).permute(0, 2, 1, 3, 4, 5)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
    att = torch.einsum("...ihd,...jhd->...hij", q / math.sqrt(32), k).softmax(-1)
    o = torch.einsum("...hij,...jhd->...ihd", att, v).permute(0, 2, 1, 3, 4).reshape(P, 256)
    ref = o @ wo + bo + xf
    assert _rel(out.reshape(P, Cc), ref) < 2e-2


# ------------------------------------------------------------------------------------------
# small kernels
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,Cin,Fr,H,W,Cout", [(2, 1, 2, 64, 64, 32), (1, 3, 2, 32, 32, 64)])
def test_init_conv(B, Cin, Fr, H, W, Cout):
    from video_diffusion_nnx_b200 import ops

    _setup()
    x = torch.randn(B, Cin, Fr, H, W, device=DEV)
    w = (torch.randn(7, 7, Cin, Cout, device=DEV) / 7).requires_grad_(True)
    bias = torch.randn(Cout, device=DEV).requires_grad_(True)
    xx = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, Cin, H, W)
    ref = F.conv2d(xx, w.permute(3, 2, 0, 1), bias, padding=3).permute(0, 2, 3, 1)
    out = torch.empty(B * Fr, H, W, Cout, dtype=torch.bfloat16, device=DEV)
    ops.init_conv_fwd(x, w.detach().contiguous(), bias.detach(), out, B, Cin, Fr, H, W, Cout, 7)
    assert _rel(out, ref) < 1e-2
    dy = _bf(B * Fr, H, W, Cout)
    ref.backward(dy.float())
    dw, db = torch.zeros_like(w), torch.zeros_like(bias)
    ops.init_conv_wgrad(x, dy, dw, db, B, Cin, Fr, H, W, Cout, 7)
    assert _rel(dw, w.grad) < 1e-4
    assert _rel(db, bias.grad) < 1e-4


@pytest.mark.parametrize("P,Cc,Co", [(4096, 32, 1), (1000, 128, 3)])
def test_final_conv(P, Cc, Co):
    from video_diffusion_nnx_b200 import ops

    _setup()
    h = _bf(P, Cc)
    hf = h.float().requires_grad_(True)
    w = (torch.randn(Cc, Co, device=DEV) / Cc ** 0.5).requires_grad_(True)
    b = torch.randn(Co, device=DEV).requires_grad_(True)
    ref = hf @ w + b
    out = torch.empty(P, Co, device=DEV)
    ops.final_conv_fwd(h, w.detach(), b.detach(), out, P, Cc, Co)
    assert _rel(out, ref) < 1e-5
    dout = torch.randn(P, Co, device=DEV)
    ref.backward(dout)
    dh = torch.empty_like(h)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.final_conv_bwd(h, dout, w.detach(), dh, dw, db, P, Cc, Co)
    assert _rel(dh, hf.grad) < 1e-2
    assert _rel(dw, w.grad) < 1e-4
    assert _rel(db, b.grad) < 1e-4


def test_time_mlp_and_heads():
    from video_diffusion_nnx_b200 import ops

    _setup()
    B, dim = 4, 32
    td = 4 * dim
    time = torch.tensor([0, 17, 500, 999], dtype=torch.int32, device=DEV)
    w1 = (torch.randn(dim, td, device=DEV) / dim ** 0.5).requires_grad_(True)
    b1 = (0.1 * torch.randn(td, device=DEV)).requires_grad_(True)
    w2 = (torch.randn(td, td, device=DEV) / td ** 0.5).requires_grad_(True)
    b2 = (0.1 * torch.randn(td, device=DEV)).requires_grad_(True)
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=DEV, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    ang = time.float()[:, None] * freq[None]
    emb_ref = torch.cat([ang.sin(), ang.cos()], -1)
    t_ref = F.gelu(emb_ref @ w1 + b1, approximate="tanh") @ w2 + b2
    emb, h1, t_out = torch.empty(B, dim, device=DEV), torch.empty(B, td, device=DEV), torch.empty(B, td, device=DEV)
    ops.time_mlp_fwd(time, w1.detach(), b1.detach(), w2.detach(), b2.detach(), emb, h1, t_out, B, dim)
    assert _rel(emb, emb_ref) < 1e-4
    assert _rel(t_out, t_ref) < 1e-4
    # heads
    couts = [32, 64, 256]
    heads, entries, off = [], [], 0
    for c in couts:
        hw = (torch.randn(td, 2 * c, device=DEV) / td ** 0.5).requires_grad_(True)
        hb = (0.1 * torch.randn(2 * c, device=DEV)).requires_grad_(True)
        lg = (1 + 0.1 * torch.randn(2 * c, device=DEV)).requires_grad_(True)
        lb = (0.1 * torch.randn(2 * c, device=DEV)).requires_grad_(True)
        g = [torch.zeros_like(p) for p in (hw, hb, lg, lb)]
        heads.append((hw, hb, lg, lb, g))
        entries.append(dict(w=hw.detach(), b=hb.detach(), ln_g=lg.detach(), ln_b=lb.detach(), dw=g[0], db=g[1],
                            dln_g=g[2], dln_b=g[3], n_out=2 * c, off=off))
        off += 2 * c
    table = ops.make_time_head_table(entries, DEV)
    e_pre, ss = torch.empty(B, off, device=DEV), torch.empty(B, off, device=DEV)
    ops.time_heads_fwd(t_out, table, len(couts), e_pre, ss, B, td)
    t_leaf = t_out.clone().requires_grad_(True)
    refs = []
    for (hw, hb, lg, lb, _), c in zip(heads, couts):
        e = F.silu(t_leaf) @ hw + hb
        mean = e.mean(-1, keepdim=True)
        var = ((e * e).mean(-1, keepdim=True) - mean * mean).clamp_min(0)
        refs.append((e - mean) * torch.rsqrt(var + 1e-6) * lg + lb)
    ss_ref = torch.cat(refs, -1)
    assert _rel(ss, ss_ref) < 1e-4
    dss = torch.randn(B, off, device=DEV)
    ss_ref.backward(dss)
    de_ws, dt = torch.empty(B, off, device=DEV), torch.empty(B, td, device=DEV)
    ops.time_heads_bwd(t_out, table, len(couts), e_pre, dss, de_ws, dt, B, td)
    assert _rel(dt, t_leaf.grad) < 1e-3
    for (hw, hb, lg, lb, g) in heads:
        assert _rel(g[0], hw.grad) < 1e-3 and _rel(g[1], hb.grad) < 1e-3
        assert _rel(g[2], lg.grad) < 1e-3 and _rel(g[3], lb.grad) < 1e-3
    # time mlp backward
    t_ref.backward(t_leaf.grad)
    dw1, db1, dw2, db2 = (torch.zeros_like(p) for p in (w1, b1, w2, b2))
    ws = torch.empty(B, td, device=DEV)
    ops.time_mlp_bwd(dt, emb, h1, w2.detach(), dw1, db1, dw2, db2, ws, B, dim)
    for got, want in ((dw1, w1.grad), (db1, b1.grad), (dw2, w2.grad), (db2, b2.grad)):
        assert _rel(got, want) < 1e-3


def test_diffusion_elementwise_colsum_adam():
    from video_diffusion_nnx_b200 import ops

    _setup()
    B, Cc, Fr, H, W, T = 3, 2, 2, 8, 8, 50
    FHW = Fr * H * W
    x = torch.rand(B, Cc, Fr, H, W, device=DEV)
    noise = torch.randn_like(x)
    t = torch.tensor([0, 7, 49], dtype=torch.int32, device=DEV)
    tabs = [torch.rand(T, device=DEV) for _ in range(5)]
    out = torch.empty_like(x)
    ops.q_sample(x, noise, t, tabs[0], tabs[1], out, True)
    ref = tabs[0][t.long()].view(B, 1, 1, 1, 1) * (2 * x - 1) + tabs[1][t.long()].view(B, 1, 1, 1, 1) * noise
    assert torch.allclose(out, ref, atol=1e-6)
    pred = torch.randn(B, Fr, H, W, Cc, device=DEV)
    for l1 in (False, True):
        pl = pred.clone().requires_grad_(True)
        d = pl.permute(0, 4, 1, 2, 3) - noise
        lref = d.abs().mean() if l1 else (d * d).mean()
        lref.backward()
        loss, dpred = torch.zeros(1, device=DEV), torch.empty_like(pred)
        ops.loss_fwd_bwd(pred, noise, loss, dpred, B, Cc, FHW, l1)
        assert abs(loss.item() - lref.item()) < 1e-5 * max(1.0, abs(lref.item()))
        assert torch.allclose(dpred, pl.grad, atol=1e-7)
    z = torch.randn_like(x)
    o2 = torch.empty_like(x)
    ops.p_sample(x, pred, z, t, tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], o2, B, Cc, FHW, True)
    g = lambda a: a[t.long()].view(B, 1, 1, 1, 1)
    x0 = (g(tabs[0]) * x - g(tabs[1]) * pred.permute(0, 4, 1, 2, 3)).clamp(-1, 1)
    ref2 = g(tabs[2]) * x0 + g(tabs[3]) * x + (t != 0).float().view(B, 1, 1, 1, 1) * torch.exp(0.5 * g(tabs[4])) * z
    assert torch.allclose(o2, ref2, atol=1e-5)
    dy = _bf(1000, 64)
    db = torch.zeros(64, device=DEV)
    ops.colsum(dy, db, 1000, 64)
    assert _rel(db, dy.float().sum(0)) < 1e-4
    a, b = _bf(4096), _bf(4096)
    o = torch.empty_like(a)
    ops.add_bf16(a, b, o)
    assert _rel(o, a.float() + b.float()) < 1e-2
    n = 10001
    p, gr = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    m, v, ema = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), p.clone()
    p0 = p.clone()
    hp = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1 - 0.9, 1 - 0.999, 0.995, 1.0, 1.0], device=DEV)
    ops.adam_ema(p, gr, m, v, ema, hp)
    m_ref, v_ref = 0.1 * gr, 0.001 * gr * gr
    p_ref = p0 - 1e-3 * (m_ref / 0.1) / ((v_ref / 0.001).sqrt() + 1e-8)
    assert torch.allclose(p, p_ref, atol=1e-6)
    assert torch.allclose(ema, 0.995 * p0 + 0.005 * p_ref, atol=1e-6)


@pytest.mark.parametrize("B,Fr,H,W,Cc", [(2, 10, 16, 16, 32), (1, 2, 64, 64, 64), (1, 16, 8, 8, 128), (2, 4, 8, 16, 256),
                                          (1, 3, 4, 4, 32)])
def test_mha_temporal_fused_fwd(B, Fr, H, W, Cc):
    """Fused projection + temporal attention vs torch fp32 on the same bf16-rounded operands, and vs the
    unfused GEMM + core kernels (identical rounding points: q, k, v rounded to bf16)."""
    from video_diffusion_nnx_b200 import ops

    _setup()
    x = _bf(B, Fr, H, W, Cc)
    w = torch.randn(Cc, 768, device=DEV) / Cc ** 0.5
    bias = 0.1 * torch.randn(768, device=DEV)
    w_hm = torch.empty(768, Cc, dtype=torch.bfloat16, device=DEV)
    b_hm = torch.empty(768, device=DEV)
    ops.qkv_headmajor_pack(w.contiguous(), bias, w_hm, b_hm, Cc)
    P = B * Fr * H * W
    o = torch.empty(P, 256, dtype=torch.bfloat16, device=DEV)
    qkv = torch.empty(P, 768, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(P, 8, device=DEV)
    ops.mha_temporal_fused_fwd(x, w_hm, b_hm, o, qkv, lse, B, Fr, H, W, Cc)
    torch.cuda.synchronize()
    wq = w.to(torch.bfloat16).float()
    qkv_ref = (x.float().reshape(P, Cc) @ wq + bias).to(torch.bfloat16)
    assert _rel(qkv, qkv_ref) < 1e-2
    t = qkv_ref.float().reshape(B, Fr, H * W, 3, 8, 32).permute(0, 2, 1, 3, 4, 5)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
    s = torch.einsum("...ihd,...jhd->...hij", q / math.sqrt(32), k)
    att = s.softmax(-1)
    o_ref = torch.einsum("...hij,...jhd->...ihd", att, v).permute(0, 2, 1, 3, 4).reshape(P, 256)
    lse_ref = torch.logsumexp(s, -1).permute(0, 3, 1, 2).reshape(P, 8)   # (B,HW,h,i) -> (B,i,HW,h)
    assert _rel(o, o_ref) < 1e-2
    assert _rel(lse, lse_ref) < 3e-3  # q, k are rounded to bf16 from differently-ordered fp32 sums
    # without the optional outputs
    o2 = torch.empty_like(o)
    ops.mha_temporal_fused_fwd(x, w_hm, b_hm, o2, None, None, B, Fr, H, W, Cc)
    assert _rel(o2, o_ref) < 1e-2  # inference mode keeps q, k, v in fp32 (no bf16 rounding)


@pytest.mark.parametrize("B,Fr,H,W,Cc", [(2, 10, 16, 16, 32), (1, 10, 8, 8, 256), (1, 16, 8, 8, 128), (2, 16, 8, 16, 64),
                                          (1, 10, 64, 64, 32), (1, 2, 8, 8, 32), (2, 16, 8, 8, 32), (2, 3, 4, 4, 32),
                                          (1, 9, 5, 7, 32)])
def test_mha_temporal_tc_fwd(B, Fr, H, W, Cc):
    """All-tensor-core temporal attention (projection, Q K^T and P V on tcgen05) vs torch fp32."""
    from video_diffusion_nnx_b200 import ops

    _setup()
    assert ops.mha_tc_supported(Fr, Cc)
    x = _bf(B, Fr, H, W, Cc)
    w = torch.randn(Cc, 768, device=DEV) / Cc ** 0.5
    bias = 0.1 * torch.randn(768, device=DEV)
    w_hm = torch.empty(768, Cc, dtype=torch.bfloat16, device=DEV)
    b_hm = torch.empty(768, device=DEV)
    ops.qkv_headmajor_pack(w.contiguous(), bias, w_hm, b_hm, Cc)
    P = B * Fr * H * W
    o = torch.zeros(P, 256, dtype=torch.bfloat16, device=DEV)
    qkv = torch.zeros(P, 768, dtype=torch.bfloat16, device=DEV)
    lse = torch.zeros(P, 8, device=DEV)
    ops.mha_temporal_tc_fwd(x, w_hm, b_hm, o, qkv, lse, B, Fr, H, W, Cc)
    torch.cuda.synchronize()
    wq = w.to(torch.bfloat16).float()
    qkv_ref = (x.float().reshape(P, Cc) @ wq + bias).to(torch.bfloat16)
    assert _rel(qkv, qkv_ref) < 1e-2
    t = qkv_ref.float().reshape(B, Fr, H * W, 3, 8, 32).permute(0, 2, 1, 3, 4, 5)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
    s = torch.einsum("...ihd,...jhd->...hij", q / math.sqrt(32), k)
    o_ref = torch.einsum("...hij,...jhd->...ihd", s.softmax(-1), v).permute(0, 2, 1, 3, 4).reshape(P, 256)
    lse_ref = torch.logsumexp(s, -1).permute(0, 3, 1, 2).reshape(P, 8)
    assert _rel(o, o_ref) < 2e-2   # P is rounded to bf16 before the P V product
    assert _rel(lse, lse_ref) < 3e-3
    # inference mode (no qkv / lse outputs) must give the same o
    o2 = torch.zeros_like(o)
    ops.mha_temporal_tc_fwd(x, w_hm, b_hm, o2, None, None, B, Fr, H, W, Cc)
    assert _rel(o2, o_ref) < 2e-2


@pytest.mark.parametrize("B,Fr,H,W", [(2, 10, 16, 16), (1, 16, 8, 32), (3, 2, 8, 8), (1, 7, 4, 4)])
def test_mha_temporal_core_fwd(B, Fr, H, W):
    """Attention core on a materialised q|k|v tensor (training engines at C >= 64) vs torch fp32, any F <= 16;
    lse must be the one the shared backward kernel expects (max + log sum of the scaled logits)."""
    from video_diffusion_nnx_b200 import ops

    _setup()
    P = B * Fr * H * W
    qkv = _bf(P, 768)
    o = torch.zeros(P, 256, dtype=torch.bfloat16, device=DEV)
    lse = torch.zeros(P, 8, device=DEV)
    ops.mha_temporal_core_fwd(qkv, o, lse, B, Fr, H, W)
    torch.cuda.synchronize()
    t = qkv.float().reshape(B, Fr, H * W, 3, 8, 32).permute(0, 2, 1, 3, 4, 5)
    q, k, v = t[..., 0, :, :], t[..., 1, :, :], t[..., 2, :, :]
    s = torch.einsum("...ihd,...jhd->...hij", q / math.sqrt(32), k)
    o_ref = torch.einsum("...hij,...jhd->...ihd", s.softmax(-1), v).permute(0, 2, 1, 3, 4).reshape(P, 256)
    lse_ref = torch.logsumexp(s, -1).permute(0, 3, 1, 2).reshape(P, 8)
    assert _rel(o, o_ref) < 2e-2   # P is rounded to bf16 before the P V product
    assert _rel(lse, lse_ref) < 3e-3
    o2 = torch.zeros_like(o)
    ops.mha_temporal_core_fwd(qkv, o2, None, B, Fr, H, W)
    assert torch.equal(o2, o)
