"""GPU parity of the split-K tap-GEMM (vdn_tapgemm_ws: thread-block clusters along K, reduction through a caller-owned
L2-resident scratch) against torch fp32 convolutions on the same bf16-rounded operands and against the one-tile-per-CTA
kernel behind vdn_tapgemm - the small-M shapes of config_v2_2's 8x8 level and the epilogue variants the engine uses
there (GroupNorm partial sums, residual, fp32 output, concat input, split dgrad output, ragged M)."""
import pytest
import torch

from test_tapgemm_gpu import _check, _conv_ref, _pack, _rand_bf16, _setup

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _splitk_on():
    """The split-K kernel is opt-in (measured slower than the one-tile kernel on config_v2_2)."""
    from video_diffusion_nnx_b200._lib import debug_switches

    with debug_switches(VDN_SPLITK=1):
        yield


def _ws(nbytes):
    assert nbytes > 0, "the launch was expected to split its K loop"
    return torch.full((nbytes,), 0xFF, dtype=torch.uint8, device="cuda")  # NaN pattern: stale scratch must not leak


@pytest.mark.parametrize(
    "n_img,H,W,cin,cout,n_src",
    [
        (40, 8, 8, 256, 256, 1),   # the conv of the 8x8 level: 20 row tiles, BN 128 x 3 K ranges
        (40, 8, 8, 256, 128, 1),   # 256 -> 128: 20 tiles, 4 K ranges
        (40, 8, 8, 128, 256, 2),   # concat input (two 128-channel sources)
        (3, 8, 8, 128, 64, 1),     # ragged M (192 rows), BK = 64, narrow N
        (10, 8, 8, 128, 96, 1),    # N = 96 -> 32-column tiles, 5 row tiles x 3, 18 K steps
        (8, 16, 16, 64, 128, 1),   # 16 row tiles at 16x16
    ],
)
def test_splitk_conv_matches_reference(n_img, H, W, cin, cout, n_src):
    from video_diffusion_nnx_b200 import ops

    _setup()
    xs = [_rand_bf16(n_img, H, W, cin) for _ in range(n_src)]
    w = _rand_bf16(9, n_src * cin, cout, scale=(9 * n_src * cin) ** -0.5).float()
    bias = torch.randn(cout, device="cuda")
    wp = _pack(w)
    ref = _conv_ref(xs, w, 3, 3) + bias
    nbytes = ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, n_img, H, W, n_src, cin, ops.TAPS_3x3, cout)
    ws = _ws(nbytes)
    out32 = ops.tapgemm(ops.VDN_TAP_UNIT, xs, wp, ops.TAPS_3x3, bias=bias, out_dtype=torch.float32, workspace=ws)
    torch.cuda.synchronize()
    _check(out32, ref, 1e-4)
    base = ops.tapgemm(ops.VDN_TAP_UNIT, xs, wp, ops.TAPS_3x3, bias=bias, out_dtype=torch.float32)
    assert (out32 - base).abs().max().item() <= 1e-4 * ref.abs().max().item()
    out = ops.tapgemm(ops.VDN_TAP_UNIT, xs, wp, ops.TAPS_3x3, bias=bias, workspace=ws)
    _check(out, ref, 1e-2)
    # the same scratch again, back to back (stream order is the only protection the contract promises)
    for _ in range(3):
        out_b = ops.tapgemm(ops.VDN_TAP_UNIT, xs, wp, ops.TAPS_3x3, bias=bias, workspace=ws)
    assert torch.equal(out_b, out)


def test_splitk_gn_sums_residual_and_split_output():
    from video_diffusion_nnx_b200 import ops

    _setup()
    B, Fr, H, W, C = 4, 10, 8, 8, 256
    n_img = B * Fr
    x = _rand_bf16(n_img, H, W, C)
    w = _rand_bf16(9, C, C, scale=(9 * C) ** -0.5).float()
    bias = torch.randn(C, device="cuda")
    wp = _pack(w)
    ref = _conv_ref([x], w, 3, 3) + bias
    for groups in (8, 4, 2):
        nbytes = ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, n_img, H, W, 1, C, ops.TAPS_3x3, C, gn_groups=groups,
                                             rows_per_sample=Fr * H * W)
        ws = _ws(nbytes)
        sums = torch.zeros(ops.GN_REPLICAS, B, groups, 2, device="cuda")
        out = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, bias=bias, gn_sums=sums, gn_groups=groups,
                          rows_per_sample=Fr * H * W, workspace=ws)
        _check(out, ref, 1e-2)
        g = ref.view(B, Fr * H * W, groups, C // groups)
        tot = sums.sum(0)
        assert torch.allclose(tot[..., 0], g.sum(dim=(1, 3)), rtol=1e-3, atol=2e-1)
        assert torch.allclose(tot[..., 1], (g * g).sum(dim=(1, 3)), rtol=1e-3, atol=2e-1)
    # residual, in place
    nbytes = ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, n_img, H, W, 1, C, ops.TAPS_3x3, C)
    ws = _ws(nbytes)
    res = _rand_bf16(n_img, H, W, C)
    acc = res.clone()
    ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, bias=bias, residual=acc, out=acc, workspace=ws)
    _check(acc, ref + res.float(), 1e-2)
    # dgrad of a concat input: N = 2 * 128 columns split over two outputs with their own residuals
    half = C // 2
    r0, r1 = _rand_bf16(n_img, H, W, half), _rand_bf16(n_img, H, W, half)
    o0, o1 = torch.empty_like(r0), torch.empty_like(r1)
    nb2 = ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, n_img, H, W, 1, C, ops.TAPS_3x3, C, split_col=half)
    ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, residual=r0, residual2=r1, out=o0, out2=o1, split_col=half,
                workspace=_ws(nb2))
    ref_nb = ref - bias
    _check(o0, ref_nb[..., :half] + r0.float(), 1e-2)
    _check(o1, ref_nb[..., half:] + r1.float(), 1e-2)


def test_splitk_not_used_for_large_m_and_scratch_too_small_is_an_error():
    from video_diffusion_nnx_b200 import ops
    from video_diffusion_nnx_b200._lib import VdnError

    _setup()
    # 40 x 32 x 32 pixels = 320 row tiles: fills the SMs on its own, no scratch wanted
    assert ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, 40, 32, 32, 1, 64, ops.TAPS_3x3, 64) == 0
    # 1x1 projection with 4 K steps: nothing to split
    assert ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, 40, 8, 8, 1, 256, ops.TAPS_1x1, 768) == 0
    x = _rand_bf16(40, 8, 8, 256)
    w = _rand_bf16(9, 256, 256, scale=48.0 ** -1).float()
    wp = _pack(w)
    small = torch.empty(1024, dtype=torch.uint8, device="cuda")
    with pytest.raises(VdnError):
        ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_3x3, workspace=small)
