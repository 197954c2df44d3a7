"""GPU parity of the slab (1,3,3) conv kernel (csrc/conv3x3_slab.cu) that vdn_tapgemm dispatches to for the
wide-channel levels (>= 64 channels per source, enough tiles to fill the GPU): against torch fp32
convolutions on the same bf16-rounded operands and against the generic tap-GEMM. VDN_SLAB_MIN_ITEMS=1 routes
small shapes through the kernel; VDN_SLAB_GRID forces a tiny grid so that one CTA walks many (group, N tile)
items: pipeline wrap-around across items, image boundaries, GroupNorm sample changes, TMEM hand-over."""
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


@pytest.fixture(params=[None, "1", "3"])
def slab_env(request):
    from video_diffusion_nnx_b200 import _lib

    _lib.debug_set("VDN_SLAB_MIN_ITEMS", 1)
    if request.param is not None:
        _lib.debug_set("VDN_SLAB_GRID", int(request.param))
    yield request.param
    _lib.debug_clear("VDN_SLAB_GRID")
    _lib.debug_clear("VDN_SLAB_MIN_ITEMS")


# (B, Fr, H, W, n_src, c, cout): every supported width, R = 4 and R = 2 groups, 64- and 128-column tiles,
# several N tiles, two sources (concat input), groups narrower and wider than 16 channels
SHAPES = [(1, 2, 128, 128, 1, 128, 128), (2, 1, 64, 64, 1, 64, 256), (1, 3, 32, 32, 2, 64, 64),
          (2, 2, 16, 16, 1, 128, 256), (1, 2, 8, 64, 1, 96, 128), (1, 1, 32, 32, 1, 256, 1024),
          (2, 1, 16, 32, 2, 128, 128)]


@pytest.mark.parametrize("B,Fr,H,W,n_src,c,cout", SHAPES)
def test_slab_forward_bias_gn(slab_env, B, Fr, H, W, n_src, c, cout):
    from video_diffusion_nnx_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
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
    assert torch.allclose(tot[..., 0], g.sum(dim=(1, 3)), rtol=2e-3, atol=1.0)
    assert torch.allclose(tot[..., 1], (g * g).sum(dim=(1, 3)), rtol=2e-3, atol=1.0)


@pytest.mark.parametrize("n_img,H,W,cin_half,cout", [(3, 32, 32, 128, 128), (2, 64, 64, 64, 64), (2, 128, 128, 128, 128)])
def test_slab_dgrad_split_and_residual(slab_env, n_img, H, W, cin_half, cout):
    """dgrad of a concat conv: N = 2*cin_half split into two outputs, each accumulated onto an aliased residual;
    then an unsplit dgrad with a separate residual tensor."""
    from video_diffusion_nnx_b200 import ops

    torch.manual_seed(2)
    cin = 2 * cin_half
    dy = _bf(n_img, H, W, cout)
    w = _bf(9, cin, cout, scale=(9 * cout) ** -0.5).float()
    wd = _pack(w, mode=1, perm=[8 - t for t in range(9)])
    wt = w.view(3, 3, cin, cout).permute(3, 2, 0, 1).contiguous()
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1).contiguous()
    r1, r2 = _bf(n_img, H, W, cin_half), _bf(n_img, H, W, cin_half)
    o1, o2 = r1.clone(), r2.clone()
    ops.tapgemm(ops.VDN_TAP_UNIT, [dy], wd, ops.TAPS_3x3, residual=o1, residual2=o2, out=o1, out2=o2,
                split_col=cin_half)
    torch.cuda.synchronize()
    assert _rel(o1, ref[..., :cin_half] + r1.float()) < 1e-2
    assert _rel(o2, ref[..., cin_half:] + r2.float()) < 1e-2
    r3 = _bf(n_img, H, W, cin)
    o3 = ops.tapgemm(ops.VDN_TAP_UNIT, [dy], wd, ops.TAPS_3x3, residual=r3)
    assert _rel(o3, ref + r3.float()) < 1e-2


def test_slab_matches_generic_tapgemm(slab_env):
    """Same operands through the generic kernel (fp32 output forces it) and the slab kernel (bf16): the bf16
    result must equal the rounded fp32 result up to bf16 rounding plus accumulation-order noise."""
    from video_diffusion_nnx_b200 import ops

    torch.manual_seed(3)
    x = _bf(3, 64, 64, 128)
    w = _bf(9, 128, 256, scale=(9 * 128) ** -0.5).float()
    wp = _pack(w)
    o32 = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, out_dtype=torch.float32)
    obf = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3)
    assert (obf.float() - o32).abs().max().item() <= 2 ** -7 * o32.abs().max().item()
