"""Stand-alone MultiheadAttention (optional focus_present_mask / pos_bias) and RelativePositionBias on the GPU against
the oracle, called the way the reference's own tests call them (test_modules.py:242-293): same shapes (b, h, w, f, c)
with 4 heads x 8, a random per-batch mask, the all-focus early return, and a (heads, f, f) bias added AFTER the softmax
(modules.py:291-321). Bucket ids are integer work: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_params(mod, prefix="m"):
    return {f"{prefix}.{k}": torch.from_numpy(v) for k, v in mod.state_dict().items()}


@pytest.mark.parametrize("heads,dim,C,shape", [(4, 8, 32, (2, 6, 7, 5)), (8, 32, 64, (3, 4, 4, 10)), (8, 32, 32, (2, 1, 9, 16))])
def test_multihead_attention_plain_mask_bias(heads, dim, C, shape):
    from oracle import unet3d_oracle as U
    from video_diffusion_nnx_b200.modules import MultiheadAttention

    b, h, w, f = shape
    rng = np.random.default_rng(10)
    x = torch.from_numpy(rng.standard_normal((b, h, w, f, C)).astype(np.float32))
    mod = MultiheadAttention(in_features=C, dim=dim, num_heads=heads, rngs=0)
    st = mod.state_dict()
    for n in ("q", "k", "v"):  # exercise the biases (zero at init)
        st[f"{n}.bias"] = (0.1 * rng.standard_normal(st[f"{n}.bias"].shape)).astype(np.float32)
    st["out.bias"] = (0.1 * rng.standard_normal(st["out.bias"].shape)).astype(np.float32)
    mod.load_state_dict(st)
    p = _oracle_params(mod)
    old_heads, old_dim = U.HEADS, U.DIM_HEAD
    U.DIM_HEAD = dim  # the oracle's scale uses its module constant
    try:
        # plain
        out = mod(x.cuda()).cpu()
        ref = U.multihead_attention(p, "m", x)
        assert out.shape == x.shape and out.dtype == torch.float32
        assert ((out - ref).norm() / ref.norm()).item() < 2e-2  # bf16 projections, fp32 core
        # post-softmax bias (modules.py:320-321)
        bias = torch.from_numpy(rng.standard_normal((heads, f, f)).astype(np.float32))
        out_b = mod(x.cuda(), pos_bias=bias.cuda()).cpu()
        ref_b = U.multihead_attention(p, "m", x, pos_bias=bias)
        assert ((out_b - ref_b).norm() / ref_b.norm()).item() < 2e-2
        assert ((out_b - out).norm() / out.norm()).item() > 0.1  # the bias really took part
        # all-focus early return out(v) (modules.py:291-292)
        ones = torch.ones(b, dtype=torch.bool)
        out_v = mod(x.cuda(), focus_present_mask=ones.cuda()).cpu()
        ref_v = U.multihead_attention(p, "m", x, focus_present_mask=ones)
        assert ((out_v - ref_v).norm() / ref_v.norm()).item() < 2e-2
        # mixed mask (modules.py:307-316): unmasked samples are untouched; masked samples carry finfo.min * v sums,
        # i.e. huge / non-finite values exactly like the reference - compare what is comparable
        mask = torch.zeros(b, dtype=torch.bool)
        mask[0] = True
        out_m = mod(x.cuda(), focus_present_mask=mask.cuda()).cpu()
        ref_m = U.multihead_attention(p, "m", x, focus_present_mask=mask)
        assert ((out_m[1:] - ref_m[1:]).norm() / ref_m[1:].norm()).item() < 2e-2
        if f > 1:
            big = (~torch.isfinite(ref_m[0])) | (ref_m[0].abs() > 1e30)
            assert big.float().mean() > 0.9
            ours_big = (~torch.isfinite(out_m[0])) | (out_m[0].abs() > 1e30)
            assert (ours_big | ~big).float().mean() > 0.99
    finally:
        U.HEADS, U.DIM_HEAD = old_heads, old_dim


@pytest.mark.parametrize("n", [1, 2, 10, 16, 33, 200])
def test_relative_position_bias_buckets_bit_exact_and_gather(n):
    from oracle import unet3d_oracle as U
    from video_diffusion_nnx_b200.modules import RelativePositionBias

    mod = RelativePositionBias(rngs=3, heads=8, num_buckets=32, max_distance=32)  # ctor args are ignored by __call__
    out = mod(n).cpu()
    assert out.shape == (8, n, n) and out.dtype == torch.float32
    pos = torch.arange(n, dtype=torch.int32)
    want = U.relative_position_bucket(pos[:, None] - pos[None, :])
    got = mod.buckets(n).cpu()
    assert got.dtype == torch.int32 and torch.equal(got, want.to(torch.int32))
    p = {"t.relative_attention_bias.embedding": torch.from_numpy(mod.state_dict()["relative_attention_bias.embedding"])}
    assert torch.equal(out, U.relative_position_bias(p, "t", n))
