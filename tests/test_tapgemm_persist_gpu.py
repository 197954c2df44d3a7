"""GPU parity of the persistent tap-GEMM (csrc/tapgemm.cu: tapgemm_persist_kernel - TMEM double buffer, producer
running ahead across tiles, TMA-store epilogue) that vdn_tapgemm dispatches to for launches with many tiles per SM
and a short K loop (projections, their dgrads, 1x1 residual convs, stride-2 convs): against torch fp32 on the same
bf16-rounded operands. VDN_PERSIST_MIN_ITEMS=1 routes small shapes through it; VDN_PERSIST_GRID forces a tiny grid
so that one CTA walks many items (ring wrap-around across items, accumulator and staging hand-over)."""
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


def _rel(a, b):
    return ((a.float() - b).abs().max() / (b.abs().max() + 1e-6)).item()


@pytest.fixture(params=[None, "1", "5"])
def persist_env(request):
    from video_diffusion_nnx_b200 import _lib

    _lib.debug_set("VDN_PERSIST_MIN_ITEMS", 1)
    if request.param is not None:
        _lib.debug_set("VDN_PERSIST_GRID", int(request.param))
    yield request.param
    _lib.debug_clear("VDN_PERSIST_GRID")
    _lib.debug_clear("VDN_PERSIST_MIN_ITEMS")


# (n_img, H, W, n_src, c, cout): 256-, 128-, 64- and 32-column tiles (four / two / one 128-byte sub-tiles, one
# 64-byte sub-tile), several N tiles, BK = 64 / 32 / 16, two sources, an M tail (5*8*8 = 320 rows)
SHAPES = [(2, 32, 32, 1, 128, 768), (3, 16, 16, 1, 256, 128), (2, 64, 64, 1, 32, 256), (4, 16, 16, 1, 256, 32),
          (5, 8, 8, 1, 64, 64), (2, 16, 16, 2, 64, 128), (2, 16, 16, 1, 16, 64), (1, 128, 128, 1, 128, 384)]


@pytest.mark.parametrize("n_img,H,W,n_src,c,cout", SHAPES)
@pytest.mark.parametrize("with_res", [False, True])
def test_persist_pointwise(persist_env, n_img, H, W, n_src, c, cout, with_res):
    from video_diffusion_nnx_b200 import ops

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(4)
    xs = [_bf(n_img, H, W, c) for _ in range(n_src)]
    w = _bf(1, n_src * c, cout, scale=(n_src * c) ** -0.5).float()
    bias = torch.randn(cout, device="cuda")
    ref = torch.cat([t.float() for t in xs], dim=-1) @ w[0] + bias
    res = _bf(n_img, H, W, cout) if with_res else None
    if with_res:
        ref = ref + res.float()
    out = ops.tapgemm(ops.VDN_TAP_UNIT, xs, _pack(w), ops.TAPS_1x1, bias=bias, residual=res)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-2


def test_persist_inplace_residual_and_split(persist_env):
    """dgrad of a 1x1 conv over a concat input: N = 2*64 split into two tensors, each accumulated in place."""
    from video_diffusion_nnx_b200 import ops

    torch.manual_seed(5)
    n_img, H, W, half, cout = 3, 32, 32, 64, 128
    dy = _bf(n_img, H, W, cout)
    w = _bf(1, 2 * half, cout, scale=cout ** -0.5).float()
    wd = _pack(w, mode=1, perm=[0])
    ref = dy.float() @ w[0].t()
    r1, r2 = _bf(n_img, H, W, half), _bf(n_img, H, W, half)
    o1, o2 = r1.clone(), r2.clone()
    ops.tapgemm(ops.VDN_TAP_UNIT, [dy], wd, ops.TAPS_1x1, residual=o1, residual2=o2, out=o1, out2=o2, split_col=half)
    torch.cuda.synchronize()
    assert _rel(o1, ref[..., :half] + r1.float()) < 1e-2
    assert _rel(o2, ref[..., half:] + r2.float()) < 1e-2
    p1, p2 = torch.empty_like(r1), torch.empty_like(r2)
    ops.tapgemm(ops.VDN_TAP_UNIT, [dy], wd, ops.TAPS_1x1, out=p1, out2=p2, split_col=half)
    assert _rel(p1, ref[..., :half]) < 1e-2 and _rel(p2, ref[..., half:]) < 1e-2


@pytest.mark.parametrize("n_img,H,W,c", [(2, 32, 32, 32), (3, 16, 16, 64)])
def test_persist_stride2_conv(persist_env, n_img, H, W, c):
    """nnx.Conv(dim, dim, (1,4,4), strides (1,2,2)) SAME (utils.py:125): 16 taps over four parity views."""
    from video_diffusion_nnx_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(6)
    x = _bf(n_img, 2 * H, 2 * W, c)
    w = _bf(16, c, c, scale=(16 * c) ** -0.5).float()
    bias = torch.randn(c, device="cuda")
    wt = w.view(4, 4, c, c).permute(3, 2, 0, 1).contiguous()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, stride=2, padding=1).permute(0, 2, 3, 1) + bias
    out = ops.tapgemm(ops.VDN_TAP_DOWN, [x], _pack(w), ops.TAPS_4x4, bias=bias)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-2
