"""GPU parity of the persistent row-ring (1,3,3) conv kernel (csrc/conv3x3_rows.cu) that vdn_tapgemm
dispatches to for the 64- and 32-pixel-wide levels: against torch fp32 convolutions on the same
bf16-rounded operands and against the generic tap-GEMM (VDN_NO_ROWCONV is read once per process, so the
generic kernel is reached through shapes the row kernel does not take). VDN_RC_GRID forces a tiny grid so
that one CTA walks many tiles: ring wrap-around, image boundaries and GroupNorm sample changes."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def _pack(w, mode=0, perm=None):
    from video_diffusion_nnx_b200 import ops

    taps, cin, cout = w.shape
    rows, k = (cout, taps * cin) if mode == 0 else (cin, taps * cout)
    dst = torch.empty(rows, k, dtype=torch.bfloat16, device="cuda")
    ops.pack_weight(w.contiguous(), dst, taps, cin, cout, mode, perm)
    return dst


def _conv_ref(xs, w):
    x = torch.cat([t.float() for t in xs], dim=-1).permute(0, 3, 1, 2)
    taps, cin, cout = w.shape
    wt = w.view(3, 3, cin, cout).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, wt, padding=1).permute(0, 2, 3, 1).contiguous()


def _rel(a, b):
    return ((a.float() - b).abs().max() / (b.abs().max() + 1e-6)).item()


@pytest.fixture(params=[None, "1", "3", "7"])
def rc_grid(request):
    from video_diffusion_nnx_b200 import _lib

    if request.param is not None:
        _lib.debug_set("VDN_RC_GRID", int(request.param))
    yield request.param
    _lib.debug_clear("VDN_RC_GRID")


@pytest.mark.parametrize("B,Fr,H,W,n_src,cout,c", [(2, 2, 64, 64, 1, 32, 32), (2, 3, 64, 64, 2, 32, 32),
                                                  (1, 5, 32, 32, 1, 64, 32), (3, 1, 32, 32, 2, 32, 32),
                                                  (1, 2, 8, 64, 1, 64, 32), (2, 3, 32, 32, 1, 64, 64),
                                                  (1, 1, 8, 32, 1, 64, 64)])
def test_rows_forward_bias_gn(rc_grid, B, Fr, H, W, n_src, cout, c):
    from video_diffusion_nnx_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    n_img = B * Fr
    xs = [_bf(n_img, H, W, c) for _ in range(n_src)]
    w = _bf(9, n_src * c, cout, scale=(9 * n_src * c) ** -0.5).float()
    bias = torch.randn(cout, device="cuda")
    ref = _conv_ref(xs, w) + bias
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device="cuda")
    out = ops.tapgemm(ops.VDN_TAP_UNIT, xs, _pack(w), ops.TAPS_3x3, bias=bias, gn_sums=sums, gn_groups=8,
                      rows_per_sample=Fr * H * W)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-2
    g = ref.view(B, Fr * H * W, 8, cout // 8)
    tot = sums.sum(0)
    assert torch.allclose(tot[..., 0], g.sum(dim=(1, 3)), rtol=2e-3, atol=0.5)
    assert torch.allclose(tot[..., 1], (g * g).sum(dim=(1, 3)), rtol=2e-3, atol=0.5)


@pytest.mark.parametrize("n_img,H,W", [(5, 64, 64), (7, 32, 32)])
def test_rows_dgrad_split_and_residual(rc_grid, n_img, H, W):
    """dgrad of a concat conv: N = 64 split into two 32-channel outputs, each with an aliased residual."""
    from video_diffusion_nnx_b200 import ops

    torch.manual_seed(2)
    cin, cout = 64, 32
    dy = _bf(n_img, H, W, cout)
    w = _bf(9, cin, cout, scale=(9 * cout) ** -0.5).float()
    wd = _pack(w, mode=1, perm=[8 - t for t in range(9)])
    wt = w.view(3, 3, cin, cout).permute(3, 2, 0, 1).contiguous()
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1).contiguous()
    r1, r2 = _bf(n_img, H, W, 32), _bf(n_img, H, W, 32)
    o1, o2 = r1.clone(), r2.clone()
    ops.tapgemm(ops.VDN_TAP_UNIT, [dy], wd, ops.TAPS_3x3, residual=o1, residual2=o2, out=o1, out2=o2, split_col=32)
    torch.cuda.synchronize()
    assert _rel(o1, ref[..., :32] + r1.float()) < 1e-2
    assert _rel(o2, ref[..., 32:] + r2.float()) < 1e-2
    # unsplit N = 32 dgrad (same-width block) with a separate residual tensor
    w2 = _bf(9, 32, 32, scale=(9 * 32) ** -0.5).float()
    wd2 = _pack(w2, mode=1, perm=[8 - t for t in range(9)])
    wt2 = w2.view(3, 3, 32, 32).permute(3, 2, 0, 1).contiguous()
    ref2 = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt2, padding=1).permute(0, 2, 3, 1).contiguous()
    o3 = ops.tapgemm(ops.VDN_TAP_UNIT, [dy], wd2, ops.TAPS_3x3, residual=r1)
    assert _rel(o3, ref2 + r1.float()) < 1e-2


def test_rows_matches_generic_tapgemm_bitwise_class():
    """Same operands through the generic kernel (fp32 output forces it) and the row kernel (bf16): the bf16
    result must equal the rounded fp32 result up to one bf16 ulp of accumulation-order noise."""
    from video_diffusion_nnx_b200 import ops

    torch.manual_seed(3)
    x = _bf(4, 64, 64, 32)
    w = _bf(9, 32, 32, scale=(9 * 32) ** -0.5).float()
    wp = _pack(w)
    o32 = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, out_dtype=torch.float32)
    obf = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3)
    assert (obf.float() - o32).abs().max().item() <= 2 ** -7 * o32.abs().max().item()
