"""GPU parity of the tcgen05 tap-GEMM against torch fp32 convolutions on identical (bf16-rounded)
operands, and against the test-only CUDA-core kernel behind the same C ABI."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)


def _rand_bf16(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def _pack(w, mode=0, perm=None):
    from video_diffusion_nnx_b200 import ops

    taps, cin, cout = w.shape
    rows, k = (cout, taps * cin) if mode == 0 else (cin, taps * cout)
    dst = torch.empty(rows, k, dtype=torch.bfloat16, device="cuda")
    ops.pack_weight(w.contiguous(), dst, taps, cin, cout, mode, perm)
    return dst


def _conv_ref(xs, w, kh, kw, stride=1, pad=1):
    # xs: list of bf16 NHWC; w fp32 [taps][cin_total][cout] with bf16-representable values
    x = torch.cat([t.float() for t in xs], dim=-1).permute(0, 3, 1, 2)
    taps, cin, cout = w.shape
    wt = w.view(kh, kw, cin, cout).permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(x, wt, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def _check(out, ref, tol=2e-2):
    out = out.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err / scale < tol, f"max abs err {err} vs scale {scale}"


@pytest.mark.parametrize(
    "n_img,H,W,cin,cout",
    [
        (1, 16, 16, 64, 64),     # single tile pair, BK=64
        (2, 64, 64, 32, 32),     # L0 shape class, BK=32 / SW64
        (3, 32, 32, 64, 128),    # BK=64, N=128
        (5, 8, 8, 256, 256),     # bn=2 boxes, M tail (320 rows)
        (2, 16, 16, 16, 48),     # BK=16 / SW32, N not a power of two
        (1, 32, 32, 128, 768),   # wide N (3 N-tiles)
    ],
)
def test_conv133_and_pointwise(n_img, H, W, cin, cout):
    from video_diffusion_nnx_b200 import ops

    _setup()
    x = _rand_bf16(n_img, H, W, cin)
    for taps, kh, pad in ((ops.TAPS_3x3, 3, 1), (ops.TAPS_1x1, 1, 0)):
        w = _rand_bf16(len(taps), cin, cout, scale=(len(taps) * cin) ** -0.5).float()
        bias = torch.randn(cout, device="cuda")
        wp = _pack(w)
        ref = _conv_ref([x], w, kh, kh, 1, pad) + bias
        out_ref_kernel = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, taps, bias=bias, out_dtype=torch.float32, ref=True)
        _check(out_ref_kernel, ref, 1e-4)
        out = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, taps, bias=bias, out_dtype=torch.float32)
        torch.cuda.synchronize()
        _check(out, ref, 1e-4)
        out_bf = ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, taps, bias=bias)
        _check(out_bf, ref, 1e-2)


def test_concat_two_sources_residual_and_gn_sums():
    from video_diffusion_nnx_b200 import ops

    _setup()
    B, Fr, H, W, c, cout = 2, 2, 16, 16, 32, 64
    n_img = B * Fr
    x0, x1 = _rand_bf16(n_img, H, W, c), _rand_bf16(n_img, H, W, c)
    w = _rand_bf16(9, 2 * c, cout, scale=(18 * c) ** -0.5).float()
    bias = torch.randn(cout, device="cuda")
    res = _rand_bf16(n_img, H, W, cout)
    wp = _pack(w)
    ref = _conv_ref([x0, x1], w, 3, 3) + bias
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device="cuda")
    out = ops.tapgemm(ops.VDN_TAP_UNIT, [x0, x1], wp, ops.TAPS_3x3, bias=bias, gn_sums=sums,
                      gn_groups=8, rows_per_sample=Fr * H * W)
    _check(out, ref, 1e-2)
    g = ref.view(B, Fr * H * W, 8, cout // 8)
    s1 = g.sum(dim=(1, 3))
    s2 = (g * g).sum(dim=(1, 3))
    tot = sums.sum(0)
    assert torch.allclose(tot[..., 0], s1, rtol=1e-3, atol=1e-1)
    assert torch.allclose(tot[..., 1], s2, rtol=1e-3, atol=1e-1)
    # residual (bf16, same dtype as the output) through the staged epilogue; in-place aliasing allowed
    out2 = ops.tapgemm(ops.VDN_TAP_UNIT, [x0, x1], wp, ops.TAPS_3x3, bias=bias, residual=res)
    _check(out2, ref + res.float(), 1e-2)
    acc = res.clone()
    ops.tapgemm(ops.VDN_TAP_UNIT, [x0, x1], wp, ops.TAPS_3x3, bias=bias, residual=acc, out=acc)
    _check(acc, ref + res.float(), 1e-2)
    # GroupNorm sums when a 128-row tile spans two samples (rows_per_sample = 96 -> per-warp path)
    xs = _rand_bf16(3, 4, 8, c)
    w1 = _rand_bf16(1, c, cout, scale=c ** -0.5).float()
    sums2 = torch.zeros(ops.GN_REPLICAS, 1, 8, 2, device="cuda")
    o3 = ops.tapgemm(ops.VDN_TAP_UNIT, [xs], _pack(w1), ops.TAPS_1x1, gn_sums=sums2, gn_groups=8,
                     rows_per_sample=96, out_dtype=torch.float32)
    g3 = o3.view(1, 96, 8, cout // 8)
    assert torch.allclose(sums2.sum(0)[..., 0], g3.sum(dim=(1, 3)), rtol=1e-3, atol=1e-2)


def test_split_output_dgrad_of_concat():
    from video_diffusion_nnx_b200 import ops

    _setup()
    n_img, H, W, cin, cout = 2, 16, 16, 64, 32  # "dgrad": K = 9*cout rows, N = cin split in two
    dy = _rand_bf16(n_img, H, W, cout)
    w = _rand_bf16(9, cin, cout, scale=(9 * cout) ** -0.5).float()
    perm = [8 - t for t in range(9)]
    wd = _pack(w, mode=1, perm=perm)
    # reference: conv_transpose == conv with flipped taps and swapped channels
    wt = w.view(3, 3, cin, cout).permute(3, 2, 0, 1).contiguous()  # [cout, cin, kh, kw] as a fwd weight
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1).contiguous()
    o1 = torch.empty(n_img, H, W, cin // 2, dtype=torch.float32, device="cuda")
    o2 = torch.empty_like(o1)
    ops.tapgemm(ops.VDN_TAP_UNIT, [dy], wd, ops.TAPS_3x3, out=o1, out2=o2, split_col=cin // 2,
                out_dtype=torch.float32)
    _check(o1, ref[..., : cin // 2], 1e-4)
    _check(o2, ref[..., cin // 2:], 1e-4)


@pytest.mark.parametrize("n_img,H,W,c", [(2, 64, 64, 32), (3, 16, 16, 128), (2, 32, 32, 64)])
def test_down_conv_k4s2(n_img, H, W, c):
    from video_diffusion_nnx_b200 import ops

    _setup()
    x = _rand_bf16(n_img, H, W, c)
    w = _rand_bf16(16, c, c, scale=(16 * c) ** -0.5).float()
    bias = torch.randn(c, device="cuda")
    wp = _pack(w)
    ref = _conv_ref([x], w, 4, 4, stride=2, pad=1) + bias
    out = ops.tapgemm(ops.VDN_TAP_DOWN, [x], wp, ops.TAPS_4x4, bias=bias, out_dtype=torch.float32)
    _check(out, ref, 1e-4)
    out_r = ops.tapgemm(ops.VDN_TAP_DOWN, [x], wp, ops.TAPS_4x4, bias=bias, out_dtype=torch.float32, ref=True)
    _check(out_r, ref, 1e-4)


@pytest.mark.parametrize("n_img,H,W,c", [(2, 32, 32, 32), (3, 8, 8, 128)])
def test_up_conv_transpose_k4s2(n_img, H, W, c):
    """nnx.ConvTranspose((1,4,4),(1,2,2)) SAME, unflipped kernel == torch conv_transpose2d with the
    spatially flipped kernel, stride 2, padding 1."""
    from video_diffusion_nnx_b200 import ops

    _setup()
    x = _rand_bf16(n_img, H, W, c)
    w = _rand_bf16(16, c, c, scale=(4 * c) ** -0.5).float()
    bias = torch.randn(c, device="cuda")
    w4 = w.view(4, 4, c, c)
    wt = w4.flip(0, 1).permute(2, 3, 0, 1).contiguous()  # [cin, cout, kh, kw], flipped
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt, stride=2, padding=1).permute(0, 2, 3, 1) + bias
    out = torch.empty(n_img, 2 * H, 2 * W, c, dtype=torch.float32, device="cuda")
    for py in range(2):
        for px in range(2):
            shifts, kidx = ops.up_class_taps(py, px)
            wp = torch.empty(c, 4 * c, dtype=torch.bfloat16, device="cuda")
            ops.pack_weight(w, wp, 4, c, c, 0, kidx)
            ops.tapgemm(ops.VDN_TAP_UP, [x], wp, shifts, bias=bias, out=out, py=py, px=px, out_dtype=torch.float32)
    _check(out, ref.contiguous(), 1e-4)


def test_pack_batched_matches_pack_weight():
    """The one-launch tiled repack (fp32 [taps][cin][cout] -> bf16 K-major operands, both modes, ragged tiles,
    tap permutation) against the single-job kernel."""
    from video_diffusion_nnx_b200 import ops

    _setup()
    jobs, refs = [], []
    for taps, cin, cout, mode, perm in [(9, 48, 80, 0, None), (9, 48, 80, 1, [8 - t for t in range(9)]),
                                        (1, 32, 768, 0, None), (4, 64, 64, 1, [5, 7, 13, 15]), (16, 16, 16, 0, None)]:
        n_src_taps = 16 if perm and max(perm) > taps - 1 else taps
        w = torch.randn(n_src_taps, cin, cout, device="cuda")
        rows, k = (cout, taps * cin) if mode == 0 else (cin, taps * cout)
        dst = torch.zeros(rows, k, dtype=torch.bfloat16, device="cuda")
        ref = torch.zeros_like(dst)
        ops.pack_weight(w, ref, taps, cin, cout, mode, perm)
        jobs.append((w, dst, taps, cin, cout, mode, perm))
        refs.append(ref)
    table, n, total = ops.make_pack_table(jobs, "cuda")
    ops.pack_batched(table, n, total)
    torch.cuda.synchronize()
    for (_, dst, *_), ref in zip(jobs, refs):
        assert torch.equal(dst, ref)
