"""From-definition numpy checks of the flax/jax layer semantics the oracle hard-codes,
the committed golden vector (pins the oracle against drift), and the integer bucket table."""
import math
import os

import numpy as np
import torch

from oracle import unet3d_oracle as U

GOLD = os.path.join(os.path.dirname(__file__), "golden", "v1_0_b2_seed3.npz")


def test_conv_transpose_same_unflipped_matches_lhs_dilation_definition():
    """nnx.ConvTranspose((1,4,4),(1,2,2)), SAME, transpose_kernel=False == conv_general_dilated with
    lhs_dilation 2, padding (2,2), stride 1 and the UNFLIPPED kernel (utils.py:113)."""
    rng = np.random.default_rng(0)
    H, W, ci, co = 3, 4, 2, 3
    x = rng.standard_normal((1, 1, H, W, ci))
    k = rng.standard_normal((1, 4, 4, ci, co))
    b = rng.standard_normal(co)
    xd = np.zeros((2 * H - 1, 2 * W - 1, ci))
    xd[::2, ::2] = x[0, 0]
    xp = np.pad(xd, ((2, 2), (2, 2), (0, 0)))
    out = np.zeros((2 * H, 2 * W, co))
    for oy in range(2 * H):
        for ox in range(2 * W):
            for a in range(4):
                for c in range(4):
                    out[oy, ox] += xp[oy + a, ox + c] @ k[0, a, c]
    out += b
    got = U.conv_transpose_k4s2(torch.from_numpy(x), torch.from_numpy(k), torch.from_numpy(b))[0, 0].numpy()
    np.testing.assert_allclose(got, out, atol=1e-10)
    # and the parity-class decomposition the CUDA path uses (ops.up_class_taps)
    from video_diffusion_nnx_b200.ops import up_class_taps

    out2 = np.zeros_like(out)
    for py in range(2):
        for px in range(2):
            shifts, kidx = up_class_taps(py, px)
            for y in range(H):
                for xx in range(W):
                    for (dy, dx), ki in zip(shifts, kidx):
                        sy, sx = y + dy, xx + dx
                        if 0 <= sy < H and 0 <= sx < W:
                            out2[2 * y + py, 2 * xx + px] += x[0, 0, sy, sx] @ k[0, ki // 4, ki % 4]
    np.testing.assert_allclose(out2 + b, out, atol=1e-10)


def test_strided_conv_same_padding_is_1_1():
    """nnx.Conv((1,4,4),(1,2,2)) SAME on even H: pad (1,1), out = H/2 (utils.py:125)."""
    rng = np.random.default_rng(1)
    H, ci, co = 6, 2, 2
    x = rng.standard_normal((1, 1, H, H, ci))
    k = rng.standard_normal((1, 4, 4, ci, co))
    xp = np.pad(x[0, 0], ((1, 1), (1, 1), (0, 0)))
    out = np.zeros((H // 2, H // 2, co))
    for oy in range(H // 2):
        for ox in range(H // 2):
            for a in range(4):
                for c in range(4):
                    out[oy, ox] += xp[2 * oy + a, 2 * ox + c] @ k[0, a, c]
    got = U.conv_khw(torch.from_numpy(x), torch.from_numpy(k), None, stride=2)[0, 0].numpy()
    np.testing.assert_allclose(got, out, atol=1e-10)


def test_group_norm_statistics_span_frames_and_space_per_sample():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 3, 4, 4, 16))
    scale, bias = rng.standard_normal(16), rng.standard_normal(16)
    want = np.zeros_like(x)
    for b in range(2):
        for g in range(8):
            blk = x[b, ..., 2 * g:2 * g + 2]
            mean = blk.mean()
            var = max(0.0, (blk * blk).mean() - mean * mean)
            want[b, ..., 2 * g:2 * g + 2] = (blk - mean) / math.sqrt(var + 1e-6)
    want = want * scale + bias
    got = U.group_norm(torch.from_numpy(x), torch.from_numpy(scale), torch.from_numpy(bias)).numpy()
    np.testing.assert_allclose(got, want, atol=1e-9)


def test_relative_position_buckets_bit_exact():
    """modules.py:351-378 with the static defaults 32/128:
    ret = (j-i<0)*16 + where(|n|<8, |n|, min(15, 8 + int(log(|n|/8)/log(16)*8)))."""
    n = 40
    pos = torch.arange(n, dtype=torch.int32)
    got = U.relative_position_bucket(pos[:, None] - pos[None, :]).numpy()
    want = np.zeros((n, n), np.int32)
    for i in range(n):
        for j in range(n):
            m = -(i - j)
            r = 16 if m < 0 else 0
            a = abs(m)
            if a < 8:
                r += a
            else:
                r += min(15, 8 + int(np.float32(np.log(np.float32(a) / np.float32(8))) / math.log(16) * 8))
            want[i, j] = r
    assert np.array_equal(got, want)
    assert got.min() >= 0 and got.max() <= 31


def test_sinusoidal_embedding_definition():
    t = torch.tensor([0, 3, 999])
    e = U.sinusoidal_pos_emb(t, 32, torch.float64).numpy()
    freq = np.exp(np.arange(16) * -(math.log(10000) / 15))
    np.testing.assert_allclose(e[:, :16], np.sin(t.numpy()[:, None] * freq), atol=1e-12)
    np.testing.assert_allclose(e[:, 16:], np.cos(t.numpy()[:, None] * freq), atol=1e-12)


def test_param_count_matches_survey():
    assert sum(int(np.prod(s)) for s in U.param_shapes(32, 1).values()) == 9_993_409
    assert sum(int(np.prod(s)) for s in U.param_shapes(128, 1).values()) == 134_552_833


def test_prenorm_is_dead_and_attention_sees_raw_input():
    """PreNorm computes the LayerNorm and discards it (modules.py:146-148): norm params do not affect
    the output and receive zero gradient."""
    torch.manual_seed(0)
    p = U.init_params(32, 1, seed=1, perturb=0.05)
    x = torch.randn(1, 2, 8, 8, 32)
    a = U.temporal_attention(p, "init_temporal_attn", x)
    p2 = dict(p)
    p2["init_temporal_attn.fn.norm.scale"] = p["init_temporal_attn.fn.norm.scale"] * 3 + 1
    assert torch.equal(a, U.temporal_attention(p2, "init_temporal_attn", x))


def test_oracle_reproduces_committed_golden():
    g = np.load(GOLD)
    p = U.init_params(32, 1, seed=3, perturb=0.05)
    eps = U.unet3d_forward(p, torch.from_numpy(g["x_noisy"]), torch.from_numpy(g["t"]), 32).numpy()
    err = np.abs(eps - g["eps_f32"]).max() / np.abs(g["eps_f32"]).max()
    assert err < 1e-4, err
