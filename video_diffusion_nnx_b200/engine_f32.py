"""fp32-grade forward engine (BASELINE.json north_star: loss and predicted noise within 1e-3 relative in fp32).

The reference computes in float32 throughout (modules.py has no dtype= anywhere). UnetEngine (engine.py) is the
throughput path: bf16 activations, bf16 tcgen05 operands, 9e-3 on the predicted noise. This engine runs the SAME graph
(unet3d.py:262-387, Appendix A.1 of SURVEY.md) with every activation in fp32 and every GEMM as a split-bf16 product on
the same tcgen05 tap-GEMM kernels (csrc/fp32_path.cu explains the arithmetic):

    x w ~= [x_hi | x_lo] [w_hi ; w_hi] + x_hi w_lo        two vdn_tapgemm launches, fp32 accumulate + fp32 output

Forward only: it serves `Unet3D.forward_fp32`, `GaussianDiffusion.p_losses(..., precision="fp32")` and the `parity`
block of bench.py. The time-embedding path (already fp32) and the module structure are shared with an inference
UnetEngine of the same shape. There is no CPU fallback."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from ._lib import VDN_F32, check, lib, ptr, stream_ptr
from .engine import GROUPS, HD, HEADS, UnetEngine
from .ops import TAPS_1x1, TAPS_3x3, TAPS_4x4, VDN_TAP_DOWN, VDN_TAP_UNIT, VDN_TAP_UP, up_class_taps

F32, BF16 = torch.float32, torch.bfloat16


class SplitGemm:
    """One conv / projection of the fp32-grade path: packed (w_hi duplicated for the (x_hi | x_lo) source pair, w_lo)."""

    def __init__(self, eng: "F32Engine", w: torch.Tensor, bias: Optional[torch.Tensor], taps, cin: int, cout: int,
                 kind: int = VDN_TAP_UNIT, tap_ids=None, py: int = 0, px: int = 0):
        self.eng, self.w, self.bias, self.taps, self.cin, self.cout, self.kind = eng, w, bias, list(taps), cin, cout, kind
        self.tap_ids = list(tap_ids) if tap_ids is not None else list(range(len(self.taps)))
        self.py, self.px = py, px
        nt, dev = len(self.taps), w.device
        self.wp2 = torch.empty(cout, nt * 2 * cin, dtype=BF16, device=dev)  # K = (tap, {hi,lo} source, channel)
        self.wphi = torch.empty(cout, nt * cin, dtype=BF16, device=dev)
        self.wplo = torch.empty(cout, nt * cin, dtype=BF16, device=dev)
        eng.gemms.append(self)

    def pack(self):
        n = self.w.numel()
        hi, lo = torch.empty(n, dtype=F32, device=self.w.device), torch.empty(n, dtype=F32, device=self.w.device)
        check(lib.vdn_f32_hilo(ptr(self.w), ptr(hi), ptr(lo), n, stream_ptr()), "vdn_f32_hilo")
        cin, cout = self.cin, self.cout
        for t, kid in enumerate(self.tap_ids):
            ops.pack_weight(hi, self.wp2, 1, cin, cout, 0, perm=[kid], k_off=t * 2 * cin)
            ops.pack_weight(hi, self.wp2, 1, cin, cout, 0, perm=[kid], k_off=t * 2 * cin + cin)
        ops.pack_weight(hi, self.wphi, len(self.tap_ids), cin, cout, 0, perm=self.tap_ids)
        ops.pack_weight(lo, self.wplo, len(self.tap_ids), cin, cout, 0, perm=self.tap_ids)

    def __call__(self, hi, lo, out, residual=None):
        k = dict(py=self.py, px=self.px, out_dtype=F32)
        if self.kind == VDN_TAP_DOWN:  # one source only: three accumulating launches
            ops.tapgemm(VDN_TAP_DOWN, [hi], self.wphi, self.taps, bias=self.bias, residual=residual, out=out, **k)
            ops.tapgemm(VDN_TAP_DOWN, [lo], self.wphi, self.taps, residual=out, out=out, **k)
        else:
            ops.tapgemm(self.kind, [hi, lo], self.wp2, self.taps, bias=self.bias, residual=residual, out=out, **k)
        ops.tapgemm(self.kind, [hi], self.wplo, self.taps, residual=out, out=out, **k)
        return out


class F32Engine:
    def __init__(self, net, B: int, Fr: int, H: int, W: int):
        self.net = net
        self.base: UnetEngine = net.engine(B, Fr, H, W, training=False)  # structure, time-embedding path, offsets
        self.store = net.store
        self.B, self.F, self.H, self.W = B, Fr, H, W
        self.dev = self.store.flat.device
        self.groups = self.base.groups
        self.gemms: List[SplitGemm] = []
        st, dim = self.store, net.dim
        v = st.view
        dims = [dim] + [dim * m for m in net.dim_mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        n = len(in_out)

        def conv3(prefix, cin, cout):
            return SplitGemm(self, v(prefix + ".proj.kernel"), v(prefix + ".proj.bias"), TAPS_3x3, cin, cout)

        def res(prefix, cin, cout):
            d = dict(prefix=prefix, cin=cin, cout=cout, c1=conv3(prefix + ".block_1", cin, cout),
                     c2=conv3(prefix + ".block_2", cout, cout),
                     rc=SplitGemm(self, v(prefix + ".res_conv.kernel"), v(prefix + ".res_conv.bias"), TAPS_1x1, cin, cout)
                     if cin != cout else None)
            return d

        def mha(prefix, C):
            return dict(prefix=prefix, C=C,
                        qkv=SplitGemm(self, v(prefix + ".qkv.kernel"), v(prefix + ".qkv.bias"), TAPS_1x1, C, 3 * HD),
                        out=SplitGemm(self, v(prefix + ".out.kernel"), v(prefix + ".out.bias"), TAPS_1x1, HD, C))

        def sla(prefix, C):
            if not self.base.use_sla:
                return None
            return dict(prefix=prefix, C=C, qkv=SplitGemm(self, v(prefix + ".qkv.kernel"), None, TAPS_1x1, C, 3 * HD),
                        out=SplitGemm(self, v(prefix + ".to_out.kernel"), None, TAPS_1x1, HD, C))

        self.init_attn = mha("init_temporal_attn", dim)
        self.downs, self.ups = [], []
        for l, (ci, co) in enumerate(in_out):
            down = SplitGemm(self, v(f"downs.{l}.4.kernel"), v(f"downs.{l}.4.bias"), TAPS_4x4, co, co, VDN_TAP_DOWN) if l < n - 1 else None
            self.downs.append((res(f"downs.{l}.0", ci, co), res(f"downs.{l}.1", co, co), sla(f"downs.{l}.2", co),
                               mha(f"downs.{l}.3", co), down))
        mid = dims[-1]
        self.mid1, self.mid2 = res("mid_block1", mid, mid), res("mid_block2", mid, mid)
        self.mid_s, self.mid_t = mha("mid_spatial_attn", mid), mha("mid_temporal_attn", mid)
        for i, (ci, co) in enumerate(reversed(in_out)):
            up = None
            if i < n - 1:
                up = []
                for py in range(2):
                    for px in range(2):
                        shifts, kidx = up_class_taps(py, px)
                        up.append(SplitGemm(self, v(f"ups.{i}.4.kernel"), v(f"ups.{i}.4.bias"), shifts, ci, ci, VDN_TAP_UP,
                                            tap_ids=kidx, py=py, px=px))
            self.ups.append((res(f"ups.{i}.0", 2 * co, ci), res(f"ups.{i}.1", ci, ci), sla(f"ups.{i}.2", ci),
                             mha(f"ups.{i}.3", ci), up))
        self.final = res("final_conv.0", 2 * dim, dim)
        self.packed_version = -1

    # -- helpers -------------------------------------------------------------------------------
    def sync_weights(self):
        if self.packed_version != self.store.version:
            for g in self.gemms:
                g.pack()
            self.packed_version = self.store.version

    def _new(self, shape, dtype=F32):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def _split(self, srcs):
        a = srcs[0]
        b = srcs[1] if len(srcs) > 1 else None
        n_img, H, W, C0 = a.shape
        C1 = b.shape[-1] if b is not None else 0
        hi, lo = self._new((n_img, H, W, C0 + C1), BF16), self._new((n_img, H, W, C0 + C1), BF16)
        check(lib.vdn_f32_split(ptr(a), ptr(b), ptr(hi), ptr(lo), n_img * H * W, C0, C1, stream_ptr()), "vdn_f32_split")
        return hi, lo

    def _gn_sums(self, x, rows):
        C = x.shape[-1]
        sums = self._new((self.B, self.groups, 2), torch.float64)
        check(lib.vdn_f32_gn_stats(ptr(x), ptr(sums), self.B, rows, C, self.groups, stream_ptr()), "vdn_f32_gn_stats")
        return sums

    def _res(self, blk, base_blk, srcs):
        """ResnetBlock (modules.py:226-243)."""
        n_img, H, W, _ = srcs[0].shape
        cout, rows = blk["cout"], (n_img // self.B) * H * W
        st, p = self.store, blk["prefix"]
        hi, lo = self._split(srcs)
        a_raw = blk["c1"](hi, lo, self._new((n_img, H, W, cout)))
        ss = None
        if base_blk.ss_off is not None:
            ss = self.base.ss[:, base_blk.ss_off:base_blk.ss_off + 2 * cout]
        a = self._new(a_raw.shape)
        check(lib.vdn_f32_gn_silu(ptr(a_raw), ptr(self._gn_sums(a_raw, rows)), ptr(st.view(p + ".block_1.norm.scale")),
                                  ptr(st.view(p + ".block_1.norm.bias")), ptr(ss), ss.stride(0) if ss is not None else 0,
                                  ptr(a), self.B, rows, cout, self.groups, stream_ptr()), "vdn_f32_gn_silu")
        hi2, lo2 = self._split([a])
        b_raw = blk["c2"](hi2, lo2, self._new(a_raw.shape))
        if blk["rc"] is not None:
            s = blk["rc"](hi, lo, self._new(a_raw.shape))
        else:
            s = srcs[0]
        out = self._new(a_raw.shape)
        check(lib.vdn_f32_tail(ptr(b_raw), ptr(self._gn_sums(b_raw, rows)), ptr(st.view(p + ".block_2.norm.scale")),
                               ptr(st.view(p + ".block_2.norm.bias")), ptr(s), ptr(st.view(p + ".norm_2.scale")),
                               ptr(st.view(p + ".norm_2.bias")), ptr(out), self.B, rows, cout, self.groups, stream_ptr()),
              "vdn_f32_tail")
        return out

    def _mha(self, blk, x, spatial: bool):
        """x + MultiheadAttention(x) (modules.py:285-326; PreNorm's LayerNorm is discarded, modules.py:146-148)."""
        n_img, H, W, C = x.shape
        hi, lo = self._split([x])
        qkv = blk["qkv"](hi, lo, self._new((n_img, H, W, 3 * HD)))
        o = self._new((n_img, H, W, HD))
        Fr = n_img // self.B
        if spatial:   # 'b f h w c -> b f (h w) c': sequences are the pixels of a frame
            n_seq, S, inner = n_img, H * W, 1
        else:         # 'b f h w c -> b (h w) f c': one sequence per pixel, tokens H*W rows apart
            n_seq, S, inner = self.B * H * W, Fr, H * W
        check(lib.vdn_mha_core_ext_fwd(ptr(qkv), ptr(o), VDN_F32, HEADS, 32, n_seq, S, inner, None, 1, None, 0, stream_ptr()),
              "vdn_mha_core_ext_fwd")
        ohi, olo = self._split([o])
        return blk["out"](ohi, olo, self._new(x.shape), residual=x)

    def _sla(self, blk, x):
        """x + SpatialLinearAttention(x) (modules.py:99-129)."""
        n_img, H, W, C = x.shape
        hi, lo = self._split([x])
        qkv = blk["qkv"](hi, lo, self._new((n_img, H, W, 3 * HD)))
        tok, ctx = self._new((n_img, H, W, HD)), self._new((n_img, HEADS, 32, 32))
        check(lib.vdn_f32_sla_core(ptr(qkv), ptr(tok), ptr(ctx), n_img, H * W, stream_ptr()), "vdn_f32_sla_core")
        thi, tlo = self._split([tok])
        return blk["out"](thi, tlo, self._new(x.shape), residual=x)

    # -- forward -------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """x fp32 (B,C,F,H,W), time int32 (B,) -> fp32 (B,F,H,W,C) (unet3d.py:262-387)."""
        self.sync_weights()
        base, st = self.base, self.store
        B, Fr, H, W, dim = self.B, self.F, self.H, self.W, self.net.dim
        n_img = B * Fr
        ops.time_mlp_fwd(time, st.view("time_mlp.1.kernel"), st.view("time_mlp.1.bias"), st.view("time_mlp.3.kernel"),
                         st.view("time_mlp.3.bias"), base.emb, base.h1, base.t_emb, B, dim)
        ops.time_heads_fwd(base.t_emb, base.head_table, base.n_heads, base.e_pre, base.ss, B, base.td)
        h = self._new((n_img, H, W, dim))
        check(lib.vdn_f32_init_conv(ptr(x), ptr(st.view("init_conv.kernel")), ptr(st.view("init_conv.bias")), ptr(h), B,
                                    self.net.channels, Fr, H, W, dim, self.net.init_kernel_size, stream_ptr()), "vdn_f32_init_conv")
        h = self._mha(self.init_attn, h, False)
        r = h
        skips = []
        for (b1, b2, sla, mha, down), (e1, e2, _, _, _) in zip(self.downs, base.downs):
            h = self._res(b1, e1, [h])
            h = self._res(b2, e2, [h])
            if sla is not None:
                h = self._sla(sla, h)
            h = self._mha(mha, h, False)
            skips.append(h)
            if down is not None:
                hi, lo = self._split([h])
                h = down(hi, lo, self._new((h.shape[0], h.shape[1] // 2, h.shape[2] // 2, h.shape[3])))
        h = self._res(self.mid1, base.mid1, [h])
        h = self._mha(self.mid_s, h, True)
        h = self._mha(self.mid_t, h, False)
        h = self._res(self.mid2, base.mid2, [h])
        for (b1, b2, sla, mha, up), (e1, e2, _, _, _) in zip(self.ups, base.ups):
            h = self._res(b1, e1, [h, skips.pop()])
            h = self._res(b2, e2, [h])
            if sla is not None:
                h = self._sla(sla, h)
            h = self._mha(mha, h, False)
            if up is not None:
                hi, lo = self._split([h])
                out = self._new((h.shape[0], 2 * h.shape[1], 2 * h.shape[2], h.shape[3]))
                for g in up:
                    g(hi, lo, out)
                h = out
        h = self._res(self.final, base.final_block, [h, r])
        out = self._new((B, Fr, H, W, self.net.out_dim))
        check(lib.vdn_f32_final_conv(ptr(h), ptr(st.view("final_conv.1.kernel")), ptr(st.view("final_conv.1.bias")), ptr(out),
                                     n_img * H * W, dim, self.net.out_dim, stream_ptr()), "vdn_f32_final_conv")
        return out
