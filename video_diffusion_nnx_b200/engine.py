"""Host-side execution engine for the Unet3D hot path.

Mirrors the module tree of the reference (unet3d.py:58-252, modules.py) as a static plan for one
(B, F, H, W): every activation / saved tensor is allocated once (stable addresses, so a whole
forward+backward can be captured in a CUDA graph), and forward / backward are sequences of C-ABI
kernel launches (include/vdn.h) on the current CUDA stream. torch is used for device memory only.

Effective graph = SURVEY.md Appendix A.1 (PreNorm's LayerNorm result is discarded, attention sees
the un-normalised input and no mask / bias: modules.py:146-148).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from .ops import TAPS_1x1, TAPS_3x3, TAPS_4x4, VDN_TAP_DOWN, VDN_TAP_UNIT, VDN_TAP_UP

HEADS = 8
DIM_HEAD = 32
HD = HEADS * DIM_HEAD  # 256; also SLA's heads * D (unet3d.py:174,225)
GROUPS = 8
BF16 = torch.bfloat16
F32 = torch.float32


# ------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------
def internal_param_spec(dim: int, channels: int, dim_mults=(1, 2, 4, 8), init_kernel_size: int = 7,
                        out_dim: Optional[int] = None, use_sla: bool = True) -> List[Tuple[str, tuple]]:
    """Internal flat layout. q/k/v kernels are stored fused as [C][768] (q | k | v) so that one
    GEMM produces qkv; state_dict()/load_state_dict() translate to the reference's nnx paths."""
    s: List[Tuple[str, tuple]] = []
    td = dim * 4
    out_dim = channels if out_dim is None else out_dim

    def mha(p, c):
        s.extend([(p + ".norm.scale", (c,)), (p + ".norm.bias", (c,)), (p + ".qkv.kernel", (c, 3 * HD)),
                  (p + ".qkv.bias", (3 * HD,)), (p + ".out.kernel", (HD, c)), (p + ".out.bias", (c,))])

    def sla(p, c):
        if not use_sla:  # use_sparse_linear_attn=False: the slot holds an Identity (unet3d.py:179-181,230-231)
            return
        s.extend([(p + ".norm.scale", (c,)), (p + ".norm.bias", (c,)), (p + ".qkv.kernel", (c, 3 * HD)),
                  (p + ".to_out.kernel", (HD, c))])

    def resnet(p, cin, cout, time=True):
        if time:
            s.extend([(p + ".mlp.kernel", (td, 2 * cout)), (p + ".mlp.bias", (2 * cout,))])
        s.extend([(p + ".norm_1.scale", (2 * cout,)), (p + ".norm_1.bias", (2 * cout,))])
        for b, ci in (("block_1", cin), ("block_2", cout)):
            s.extend([(f"{p}.{b}.proj.kernel", (9, ci, cout)), (f"{p}.{b}.proj.bias", (cout,)),
                      (f"{p}.{b}.norm.scale", (cout,)), (f"{p}.{b}.norm.bias", (cout,))])
        if cin != cout:
            s.extend([(p + ".res_conv.kernel", (1, cin, cout)), (p + ".res_conv.bias", (cout,))])
        s.extend([(p + ".norm_2.scale", (cout,)), (p + ".norm_2.bias", (cout,))])

    k = init_kernel_size
    s.append(("time_rel_pos_bias.embedding", (32, HEADS)))
    s.extend([("init_conv.kernel", (k * k, channels, dim)), ("init_conv.bias", (dim,))])
    mha("init_temporal_attn", dim)
    s.extend([("time_mlp.1.kernel", (dim, td)), ("time_mlp.1.bias", (td,)),
              ("time_mlp.3.kernel", (td, td)), ("time_mlp.3.bias", (td,))])
    dims = [dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    n = len(in_out)
    for l, (ci, co) in enumerate(in_out):
        resnet(f"downs.{l}.0", ci, co)
        resnet(f"downs.{l}.1", co, co)
        sla(f"downs.{l}.2", co)
        mha(f"downs.{l}.3", co)
        if l < n - 1:
            s.extend([(f"downs.{l}.4.kernel", (16, co, co)), (f"downs.{l}.4.bias", (co,))])
    mid = dims[-1]
    resnet("mid_block1", mid, mid)
    mha("mid_spatial_attn", mid)
    mha("mid_temporal_attn", mid)
    resnet("mid_block2", mid, mid)
    for i, (ci, co) in enumerate(reversed(in_out)):
        resnet(f"ups.{i}.0", co * 2, ci)
        resnet(f"ups.{i}.1", ci, ci)
        sla(f"ups.{i}.2", ci)
        mha(f"ups.{i}.3", ci)
        if i < n - 1:
            s.extend([(f"ups.{i}.4.kernel", (16, ci, ci)), (f"ups.{i}.4.bias", (ci,))])
    resnet("final_conv.0", dim * 2, dim, time=False)
    s.extend([("final_conv.1.kernel", (dim, out_dim)), ("final_conv.1.bias", (out_dim,))])

    # "late" parameters (their gradients are completed by the LAST backward stage) go first, the rest
    # stays in execution order: see UnetEngine.backward_stages / grad_slices.
    return [e for e in s if is_late_param(e[0])] + [e for e in s if not is_late_param(e[0])]


def is_late_param(name: str) -> bool:
    return name.startswith(("time_", "init_")) or ".mlp." in name or ".norm_1." in name


class ParamStore:
    """Flat fp32 master parameters (+ gradients) with named views; 16-byte aligned segments."""

    def __init__(self, spec: Sequence[Tuple[str, tuple]], device, with_grad: bool):
        self.spec = list(spec)
        self.offsets: Dict[str, Tuple[int, tuple]] = {}
        off = 0
        for name, shape in self.spec:
            n = 1
            for d in shape:
                n *= d
            self.offsets[name] = (off, shape)
            off += (n + 3) // 4 * 4
        self.total = off
        self.flat = torch.zeros(off, dtype=F32, device=device)
        self.grad = torch.zeros(off, dtype=F32, device=device) if with_grad else None
        # bumped by every writer of `flat` (optimizer step, state upload): engines compare it with the version their
        # packed bf16 GEMM operands were built from and repack lazily (UnetEngine.sync_weights)
        self.version = 0

    def ensure_grad(self) -> None:
        """Adds the gradient buffer in place (the store object, and with it every engine / captured graph that holds
        views of `flat`, stays valid)."""
        if self.grad is None:
            self.grad = torch.zeros(self.total, dtype=F32, device=self.flat.device)

    def bump(self) -> None:
        self.version += 1

    def view(self, name: str) -> torch.Tensor:
        off, shape = self.offsets[name]
        n = 1
        for d in shape:
            n *= d
        return self.flat[off:off + n].view(shape)

    def gview(self, name: str) -> torch.Tensor:
        off, shape = self.offsets[name]
        n = 1
        for d in shape:
            n *= d
        return self.grad[off:off + n].view(shape)


# ------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------
class _Pool:
    """Exact-shape scratch pool. The get/put sequence of a step is static, so after the first eager
    step no allocation happens and buffer addresses are stable (CUDA-graph safe)."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[tuple, List[torch.Tensor]] = {}
        self.deferred: Optional[List[torch.Tensor]] = None

    def get(self, shape, dtype=BF16) -> torch.Tensor:
        key = (tuple(shape), dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        return torch.empty(shape, dtype=dtype, device=self.device)

    def put(self, t: torch.Tensor) -> None:
        if self.deferred is not None:  # inside a backward stage with deferred joins: recycled at the end of the stage
            self.deferred.append(t)
            return
        self.free.setdefault((tuple(t.shape), t.dtype), []).append(t)

    def release_deferred(self) -> None:
        bufs, self.deferred = self.deferred or [], None
        for t in bufs:
            self.put(t)


class GemmConv:
    """A conv / projection with its packed tensor-core operands (forward and dgrad)."""

    def __init__(self, eng: "UnetEngine", wname: str, bname: Optional[str], taps, n_src: int, c_src: int, cout: int):
        self.eng, self.taps, self.n_src, self.c_src, self.cout = eng, list(taps), n_src, c_src, cout
        self.cin = n_src * c_src
        st = eng.store
        self.w = st.view(wname)
        self.bias = st.view(bname) if bname else None
        nt = len(self.taps)
        self.wp = torch.empty(cout, nt * self.cin, dtype=BF16, device=eng.device)
        self.wd = torch.empty(self.cin, nt * cout, dtype=BF16, device=eng.device) if eng.training else None
        if eng.training:
            self.dw = st.gview(wname)
            self.dbias = st.gview(bname) if bname else None
        self.flip = [nt - 1 - t for t in range(nt)]
        self._ws_bytes = {}
        eng.pack_jobs.append((self.w, self.wp, nt, self.cin, self.cout, 0, None))
        if self.wd is not None:
            eng.pack_jobs.append((self.w, self.wd, nt, self.cin, self.cout, 1, self.flip))

    def _ws(self, x, n_src, n_out, split_col=0, gn_groups=0, rows_per_sample=0):
        """Split-K scratch of this launch (vdn_tapgemm_ws) from the engine's per-stream-lane buffer, or None."""
        n_img, H, W, C = x.shape
        key = (n_img, H, W, C, n_src, n_out, split_col, gn_groups, rows_per_sample)
        nbytes = self._ws_bytes.get(key)
        if nbytes is None:
            nbytes = ops.tapgemm_workspace_bytes(VDN_TAP_UNIT, n_img, H, W, n_src, C, self.taps, n_out,
                                                 split_col=split_col, gn_groups=gn_groups,
                                                 rows_per_sample=rows_per_sample)
            self._ws_bytes[key] = nbytes
        return self.eng.splitk_ws(nbytes) if nbytes else None

    def fwd(self, srcs, out, residual=None, gn_sums=None, rows_per_sample=0, out_dtype=BF16):
        groups = self.eng.groups if gn_sums is not None else 0
        return ops.tapgemm(VDN_TAP_UNIT, srcs, self.wp, self.taps, bias=self.bias, residual=residual, out=out,
                           gn_sums=gn_sums, gn_groups=groups, rows_per_sample=rows_per_sample, out_dtype=out_dtype,
                           workspace=self._ws(srcs[0], len(srcs), self.cout, gn_groups=groups,
                                              rows_per_sample=rows_per_sample))

    def dgrad(self, dy, outs, residuals=None):
        """dsrc(s) = dy (*) W^T (+ residuals). outs: 1 or 2 tensors (concat split)."""
        r = residuals or [None] * len(outs)
        if len(outs) == 1:
            ops.tapgemm(VDN_TAP_UNIT, [dy], self.wd, self.taps, residual=r[0], out=outs[0],
                        workspace=self._ws(dy, 1, self.cin))
        else:
            ops.tapgemm(VDN_TAP_UNIT, [dy], self.wd, self.taps, residual=r[0], residual2=r[1], out=outs[0],
                        out2=outs[1], split_col=self.c_src, workspace=self._ws(dy, 1, self.cin, split_col=self.c_src))

    def wgrad(self, srcs, dy, bias_done=False):
        # the bias gradient (column sums of dy) rides on the weight-gradient GEMM (spare "ones" M atom)
        ops.wgrad(VDN_TAP_UNIT, srcs, dy, self.dw, self.taps,
                  dbias=self.dbias if (self.dbias is not None and not bias_done) else None)


class ResBlock:
    """ResnetBlock (modules.py:182-243)."""

    def __init__(self, eng, prefix, n_src, c_src, cout, n_img, H, W, time=True):
        self.eng, self.prefix = eng, prefix
        self.n_src, self.c_src, self.cout, self.n_img, self.H, self.W = n_src, c_src, cout, n_img, H, W
        cin = n_src * c_src
        st = eng.store
        self.conv1 = GemmConv(eng, prefix + ".block_1.proj.kernel", prefix + ".block_1.proj.bias", TAPS_3x3, n_src, c_src, cout)
        self.conv2 = GemmConv(eng, prefix + ".block_2.proj.kernel", prefix + ".block_2.proj.bias", TAPS_3x3, 1, cout, cout)
        self.res = (GemmConv(eng, prefix + ".res_conv.kernel", prefix + ".res_conv.bias", TAPS_1x1, n_src, c_src, cout)
                    if cin != cout else None)
        self.p = {k: st.view(f"{prefix}.{k}") for k in ("block_1.norm.scale", "block_1.norm.bias", "block_2.norm.scale",
                                                         "block_2.norm.bias", "norm_2.scale", "norm_2.bias")}
        if eng.training:
            self.g = {k: st.gview(f"{prefix}.{k}") for k in self.p}
            self.t_off = eng.reserve_zeroed(2 * eng.B * cout * 2)  # the two GroupNorm backward passes' [B][C][2] sums
        shape = (n_img, H, W, cout)
        self.a_raw, self.a, self.b_raw, self.out = (eng.new(shape) for _ in range(4))
        self.s = eng.new(shape) if self.res is not None else None
        self.sums1, self.sums2 = eng.new_gn_sums(), eng.new_gn_sums()
        self.rows = n_img // eng.B * H * W
        self.ss_off = eng.register_time_head(prefix, cout) if time else None
        self.srcs = None

    def _ss(self):
        if self.ss_off is None:
            return None
        return self.eng.ss[:, self.ss_off:self.ss_off + 2 * self.cout]

    def forward(self, srcs):
        eng, B = self.eng, self.eng.B
        self.srcs = list(srcs)
        hres = None
        if self.res is not None:  # the 1x1 residual projection is independent of the conv chain: side stream
            hres = eng.side(lambda: self.res.fwd(srcs, self.s))
            s = self.s
        else:
            s = srcs[0]
        self.conv1.fwd(srcs, self.a_raw, gn_sums=self.sums1, rows_per_sample=self.rows)
        ops.gn_silu_fwd(self.a_raw, self.sums1, self.p["block_1.norm.scale"], self.p["block_1.norm.bias"], self._ss(),
                        self.a, B, self.rows, self.cout, G=eng.groups)
        self.conv2.fwd([self.a], self.b_raw, gn_sums=self.sums2, rows_per_sample=self.rows)
        eng.join(hres)
        ops.resblock_tail_fwd(self.b_raw, self.sums2, self.p["block_2.norm.scale"], self.p["block_2.norm.bias"], s,
                              self.p["norm_2.scale"], self.p["norm_2.bias"], self.out, B, self.rows, self.cout, G=eng.groups)
        return self.out

    def backward(self, dout, extra=None):
        """Returns the list of source gradients (freshly pooled buffers; caller releases). `extra`
        (optional) is added to the first source's gradient."""
        eng, B, C = self.eng, self.eng.B, self.cout
        pool = eng.pool
        shape = (self.n_img, self.H, self.W, C)
        P = self.n_img * self.H * self.W
        s = self.s if self.res is not None else self.srcs[0]
        ds = pool.get(shape)
        # LayerNorm backward of the residual branch does not depend on the GroupNorm chain: side stream
        hln = eng.side(lambda: ops.ln_bwd(s, dout, self.p["norm_2.scale"], ds, self.g["norm_2.scale"],
                                          self.g["norm_2.bias"], P, C), lane=1)
        T2, T1 = eng.zeroed(self.t_off, B * C * 2).view(B, C, 2), eng.zeroed(self.t_off + B * C * 2, B * C * 2).view(B, C, 2)
        db_raw = pool.get(shape)
        ops.gn_silu_bwd(dout, self.b_raw, self.sums2, self.p["block_2.norm.scale"], self.p["block_2.norm.bias"], None,
                        T2, db_raw, self.g["block_2.norm.scale"], self.g["block_2.norm.bias"], None, B, self.rows, C,
                        G=eng.groups, dconv_bias=self.conv2.dbias, prezeroed=True)
        # weight gradients are off the critical path: they run on the side stream, overlapped with the
        # data-gradient chain below, and are joined before their operands go back to the pool
        h2 = eng.side(lambda: self.conv2.wgrad([self.a], db_raw, bias_done=True))
        da = pool.get(shape)
        self.conv2.dgrad(db_raw, [da])
        da_raw = pool.get(shape)
        dss = eng.dss[:, self.ss_off:self.ss_off + 2 * C] if self.ss_off is not None else None
        ops.gn_silu_bwd(da, self.a_raw, self.sums1, self.p["block_1.norm.scale"], self.p["block_1.norm.bias"],
                        self._ss(), T1, da_raw, self.g["block_1.norm.scale"], self.g["block_1.norm.bias"], dss, B,
                        self.rows, C, G=eng.groups, dconv_bias=self.conv1.dbias, prezeroed=True)
        pool.put(da)
        h1 = eng.side(lambda: self.conv1.wgrad(self.srcs, da_raw, bias_done=True))
        sshape = (self.n_img, self.H, self.W, self.c_src)
        dsrc = [pool.get(sshape) for _ in range(self.n_src)]
        hr = None
        if self.res is not None:
            hr = eng.side(lambda: self.res.wgrad(self.srcs, ds), lane=1)  # same stream as ln_bwd: ordered after it
            self.conv1.dgrad(da_raw, dsrc)
            eng.join(hln)
            self.res.dgrad(ds, dsrc, residuals=dsrc)  # in-place accumulate
        else:
            eng.join(hln)
            self.conv1.dgrad(da_raw, dsrc, residuals=[ds])
        eng.join_later(h2, h1, hr)
        pool.put(db_raw)
        pool.put(ds)
        pool.put(da_raw)
        if extra is not None:
            ops.add_bf16(dsrc[0], extra, dsrc[0])
        return dsrc


class MHABlock:
    """x + MultiheadAttention(x) over frames (mode 0) or over pixels of a frame (mode 1)
    (unet3d.py:86-96,118-120,196-208; modules.py:247-326 without mask / bias)."""

    def __init__(self, eng, prefix, C, n_img, H, W, mode):
        self.eng, self.C, self.n_img, self.H, self.W, self.mode = eng, C, n_img, H, W, mode
        self.qkv_proj = GemmConv(eng, prefix + ".qkv.kernel", prefix + ".qkv.bias", TAPS_1x1, 1, C, 3 * HD)
        self.out_proj = GemmConv(eng, prefix + ".out.kernel", prefix + ".out.bias", TAPS_1x1, 1, HD, C)
        # temporal attention runs the fused projection + core kernel; qkv is only materialised when the
        # (unfused) backward needs it, i.e. in training engines
        self.fused = mode == 0 and (n_img // eng.B) <= 16
        # inference engines at C = 32: the whole block (projections, core, out projection, residual) is one
        # kernel on folded weights (A_h = W_q W_k^T, M_h = W_v W_o): no q/k/v/o tensors
        self.folded = self.fused and (not eng.training) and C == 32
        # training engines at C >= 64: the projection runs as a tap-GEMM (full tensor-core rate, qkv has to be
        # materialised for the backward anyway) and the F x F core as a register-resident warp-MMA kernel
        # (VDN_MHA_SPLIT_FWD=0: the single-kernel tcgen05 version, which is a latency chain per head)
        self.split_fwd = (self.fused and eng.training and C >= 64 and _lib.host_flag("VDN_MHA_SPLIT_FWD", "1") != "0")
        need_qkv = eng.training or not self.fused
        self.qkv = eng.new((n_img, H, W, 3 * HD)) if need_qkv else None
        self.o = None if self.folded else eng.new((n_img, H, W, HD))
        self.lse = eng.new((n_img * H * W, HEADS), F32) if need_qkv else None
        self.out = eng.new((n_img, H, W, C))
        if self.folded:
            dev = eng.device
            self.fa, self.fm = (torch.empty(HEADS, 32, 32, dtype=BF16, device=dev) for _ in range(2))
            self.fu, self.fb = torch.empty(HEADS, 32, dtype=F32, device=dev), torch.empty(32, dtype=F32, device=dev)
            self._wq, self._bq = eng.store.view(prefix + ".qkv.kernel"), eng.store.view(prefix + ".qkv.bias")
            self._wo, self._bo = eng.store.view(prefix + ".out.kernel"), eng.store.view(prefix + ".out.bias")
            eng.extra_packers.append(self._repack_folded)
        elif self.fused and not self.split_fwd:
            self.w_hm = torch.empty(3 * HD, C, dtype=BF16, device=eng.device)
            self.b_hm = torch.empty(3 * HD, dtype=F32, device=eng.device)
            self._wq, self._bq = eng.store.view(prefix + ".qkv.kernel"), eng.store.view(prefix + ".qkv.bias")
            eng.extra_packers.append(self._repack_hm)
        self.x = None

    def _repack_folded(self):
        ops.mha_fold_pack(self._wq, self._bq, self._wo, self._bo, self.fa, self.fu, self.fm, self.fb)

    def _repack_hm(self):
        ops.qkv_headmajor_pack(self._wq, self._bq, self.w_hm, self.b_hm, self.C)

    def forward(self, x):
        eng = self.eng
        self.x = x
        Fr = self.n_img // eng.B
        if self.folded:
            ops.mha_temporal_folded_fwd(x, self.fa, self.fu, self.fm, self.fb, self.out, eng.B, Fr, self.H, self.W,
                                        self.C)
            return self.out
        if self.split_fwd:
            self.qkv_proj.fwd([x], self.qkv)
            ops.mha_temporal_core_fwd(self.qkv, self.o, self.lse, eng.B, Fr, self.H, self.W)
        elif self.fused:
            fn = ops.mha_temporal_tc_fwd if ops.mha_tc_supported(Fr, self.C) else ops.mha_temporal_fused_fwd
            fn(x, self.w_hm, self.b_hm, self.o, self.qkv, self.lse, eng.B, Fr, self.H, self.W, self.C)
        else:
            self.qkv_proj.fwd([x], self.qkv)
            ops.mha_core_fwd(self.qkv, self.o, self.lse, self.mode, eng.B, Fr, self.H * self.W)
        self.out_proj.fwd([self.o], self.out, residual=x)
        return self.out

    def backward(self, dout):
        eng, pool = self.eng, self.eng.pool
        ho = eng.side(lambda: self.out_proj.wgrad([self.o], dout))
        do = pool.get(self.o.shape)
        self.out_proj.dgrad(dout, [do])
        dqkv = pool.get(self.qkv.shape)
        Fr = self.n_img // eng.B
        if self.fused and ops.mha_tc_supported(Fr, self.C):
            ops.mha_temporal_tc_bwd(self.qkv, do, self.lse, dqkv, eng.B, Fr, self.H, self.W)
        elif self.fused:
            ops.mha_temporal_bwd(self.qkv, self.o, do, self.lse, dqkv, eng.B, Fr, self.H, self.W)
        else:
            D = pool.get(self.lse.shape, F32)
            ops.mha_core_bwd(self.qkv, self.o, do, self.lse, D, dqkv, self.mode, eng.B, self.n_img // eng.B,
                             self.H * self.W)
            pool.put(D)
        pool.put(do)
        hq = eng.side(lambda: self.qkv_proj.wgrad([self.x], dqkv))
        dx = pool.get(self.x.shape)
        self.qkv_proj.dgrad(dqkv, [dx], residuals=[dout])
        eng.join_later(ho, hq)  # dout belongs to the caller, dqkv to the pool
        pool.put(dqkv)
        return dx


class SLABlock:
    """x + SpatialLinearAttention(x) (modules.py:64-129; unet3d.py:170-178)."""

    def __init__(self, eng, prefix, C, n_img, H, W):
        self.eng, self.C, self.n_img, self.H, self.W = eng, C, n_img, H, W
        self.qkv_proj = GemmConv(eng, prefix + ".qkv.kernel", None, TAPS_1x1, 1, C, 3 * HD)
        self.out_proj = GemmConv(eng, prefix + ".to_out.kernel", None, TAPS_1x1, 1, HD, C)
        N = H * W
        # inference engines at C = 32 run the fused x -> out kernels: no q/k/v/tok tensors at all
        self.fused = (not eng.training) and C == 32
        self.qkv = None if self.fused else eng.new((n_img, H, W, 3 * HD))
        self.tok = None if self.fused else eng.new((n_img, H, W, HD))
        self.ctx = eng.new((n_img, HEADS, 32, 32), F32)
        if eng.training:
            self.dctx_off = eng.reserve_zeroed(n_img * HEADS * 32 * 32)
        self.kstat = eng.new((n_img, HEADS, 2, 32), F32)
        self.ws = eng.shared_ws(ops.sla_workspace_floats(n_img, N))
        self.out = eng.new((n_img, H, W, C))
        self.x = None

    def forward(self, x):
        self.x = x
        if self.fused:
            ops.sla_fused_fwd(x, self.qkv_proj.wp, self.out_proj.wp, self.out, self.ctx, self.kstat, self.ws,
                              self.n_img, self.H * self.W, self.C)
            return self.out
        self.qkv_proj.fwd([x], self.qkv)
        ops.sla_core_fwd(self.qkv, self.tok, self.ctx, self.kstat, self.ws, self.n_img, self.H * self.W)
        self.out_proj.fwd([self.tok], self.out, residual=x)
        return self.out

    def backward(self, dout):
        eng, pool = self.eng, self.eng.pool
        ho = eng.side(lambda: self.out_proj.wgrad([self.tok], dout))
        dtok = pool.get(self.tok.shape)
        self.out_proj.dgrad(dout, [dtok])
        dqkv = pool.get(self.qkv.shape)
        dctx = eng.zeroed(self.dctx_off, self.ctx.numel()).view(self.ctx.shape)
        ops.sla_core_bwd(self.qkv, dtok, self.ctx, self.kstat, dctx, dqkv, self.n_img, self.H * self.W, prezeroed=True)
        pool.put(dtok)
        hq = eng.side(lambda: self.qkv_proj.wgrad([self.x], dqkv))
        dx = pool.get(self.x.shape)
        self.qkv_proj.dgrad(dqkv, [dx], residuals=[dout])
        eng.join_later(ho, hq)
        pool.put(dqkv)
        return dx


class DownConv:
    """nnx.Conv(dim,dim,(1,4,4),(1,2,2)) SAME (utils.py:125)."""

    def __init__(self, eng, prefix, C, n_img, H, W):
        self.eng, self.C, self.n_img, self.H, self.W = eng, C, n_img, H, W
        st = eng.store
        self.w, self.bias = st.view(prefix + ".kernel"), st.view(prefix + ".bias")
        self.wp = torch.empty(C, 16 * C, dtype=BF16, device=eng.device)
        self.out = eng.new((n_img, H // 2, W // 2, C))
        self.cls = []
        if eng.training:
            self.dw, self.dbias = st.gview(prefix + ".kernel"), st.gview(prefix + ".bias")
            for py in range(2):
                for px in range(2):
                    kys, kxs = (1 - py, 3 - py), (1 - px, 3 - px)
                    shifts = [((py + 1 - ky) // 2, (px + 1 - kx) // 2) for ky in kys for kx in kxs]
                    kidx = [ky * 4 + kx for ky in kys for kx in kxs]
                    self.cls.append((py, px, shifts, kidx, torch.empty(C, 4 * C, dtype=BF16, device=eng.device)))
        eng.pack_jobs.append((self.w, self.wp, 16, C, C, 0, None))
        for _, _, _, kidx, wd in self.cls:
            eng.pack_jobs.append((self.w, wd, 4, C, C, 1, kidx))
        self.x = None

    def forward(self, x):
        self.x = x
        ops.tapgemm(VDN_TAP_DOWN, [x], self.wp, TAPS_4x4, bias=self.bias, out=self.out)
        return self.out

    def backward(self, dy, acc):
        """acc (same shape as x) += dgrad(dy); returns acc."""
        hw = self.eng.side(lambda: ops.wgrad(VDN_TAP_DOWN, [self.x], dy, self.dw, TAPS_4x4, dbias=self.dbias))
        for py, px, shifts, _, wd in self.cls:
            ops.tapgemm(VDN_TAP_UP, [dy], wd, shifts, residual=acc, out=acc, py=py, px=px)
        self.eng.join_later(hw)
        return acc


class UpConv:
    """nnx.ConvTranspose(dim,dim,(1,4,4),(1,2,2)) SAME, unflipped kernel (utils.py:113)."""

    def __init__(self, eng, prefix, C, n_img, H, W):
        self.eng, self.C, self.n_img, self.H, self.W = eng, C, n_img, H, W
        st = eng.store
        self.w, self.bias = st.view(prefix + ".kernel"), st.view(prefix + ".bias")
        self.out = eng.new((n_img, 2 * H, 2 * W, C))
        self.cls = []
        for py in range(2):
            for px in range(2):
                shifts, kidx = ops.up_class_taps(py, px)
                self.cls.append((py, px, shifts, kidx, torch.empty(C, 4 * C, dtype=BF16, device=eng.device)))
        self.wd = None
        if eng.training:
            self.dw, self.dbias = st.gview(prefix + ".kernel"), st.gview(prefix + ".bias")
            self.wd = torch.empty(C, 16 * C, dtype=BF16, device=eng.device)
        for _, _, _, kidx, wp in self.cls:
            eng.pack_jobs.append((self.w, wp, 4, C, C, 0, kidx))
        if self.wd is not None:
            eng.pack_jobs.append((self.w, self.wd, 16, C, C, 1, [15 - k for k in range(16)]))
        self.x = None

    def forward(self, x):
        self.x = x
        for py, px, shifts, _, wp in self.cls:
            ops.tapgemm(VDN_TAP_UP, [x], wp, shifts, bias=self.bias, out=self.out, py=py, px=px)
        return self.out

    def backward(self, dy):
        pool = self.eng.pool
        hw = self.eng.side(lambda: ops.wgrad(VDN_TAP_UP, [self.x], dy, self.dw, TAPS_4x4, dbias=self.dbias))
        dx = pool.get(self.x.shape)
        ops.tapgemm(VDN_TAP_DOWN, [dy], self.wd, TAPS_4x4, out=dx)
        self.eng.join_later(hw)
        return dx


# ------------------------------------------------------------------------------------------
# the engine
# ------------------------------------------------------------------------------------------
class UnetEngine:
    def side(self, fn, lane: int = 0):
        """Runs fn() on side stream `lane`, ordered after everything enqueued so far on the current stream; returns
        an event to join(). Works eagerly and under CUDA-graph capture (fork / join edges of the graph)."""
        if self.side_stream is None:
            fn()
            return None
        stream = self.side_stream if lane == 0 else self.side_stream2
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        stream.wait_event(ev)
        prev_lane, self._lane = self._lane, lane + 1
        try:
            with torch.cuda.stream(stream):
                fn()
                done = torch.cuda.Event()
                done.record(stream)
        finally:
            self._lane = prev_lane
        return done

    def splitk_ws(self, nbytes: int) -> torch.Tensor:
        """Scratch for split-K tap-GEMM launches of the CURRENT stream lane (main / side / side 2): launches of one
        lane are stream-ordered, so they can share one L2-resident buffer; lanes run concurrently and must not."""
        buf = self._splitk_ws.get(self._lane)
        if buf is None or buf.numel() < nbytes:
            assert not torch.cuda.is_current_stream_capturing(), "split-K scratch must exist before graph capture"
            buf = torch.empty(max(nbytes, 8 << 20), dtype=torch.uint8, device=self.device)
            self._splitk_ws[self._lane] = buf
        return buf

    def join(self, *handles):
        main = torch.cuda.current_stream()
        for h in handles:
            if h is not None:
                main.wait_event(h)

    def join_later(self, *handles):
        """Join of side-stream work whose RESULTS the dependency chain does not need (weight gradients): inside a
        backward stage the wait moves to the end of the stage, together with the recycling of every scratch buffer the
        stage released - the chain no longer stalls behind a weight-gradient GEMM at the end of every block (180 GB of
        HBM pay for the extra live buffers). Outside a stage it is a plain join."""
        if self.pool.deferred is None:
            self.join(*handles)
        else:
            self._pending.extend(h for h in handles if h is not None)

    def reserve_zeroed(self, n_floats: int) -> int:
        """Reserves n fp32 of the per-step zeroed scratch (accumulators that kernels add into: GroupNorm backward sums,
        SpatialLinearAttention dctx). One memset per step at the start of the backward instead of a memset node in front
        of every consumer on the dependency chain."""
        off = self._zeroed_total
        self._zeroed_total += (n_floats + 63) // 64 * 64
        return off

    def zeroed(self, off: int, n_floats: int) -> torch.Tensor:
        if self._zeroed is None or self._zeroed.numel() < self._zeroed_total:
            assert not torch.cuda.is_current_stream_capturing(), "zeroed scratch must exist before graph capture"
            self._zeroed = torch.zeros(self._zeroed_total, dtype=F32, device=self.device)
        return self._zeroed[off:off + n_floats]

    def begin_stage(self):
        if self.defer_joins:
            self.pool.deferred = []

    def end_stage(self):
        if self.pool.deferred is not None:
            self.join(*self._pending)
            self._pending.clear()
            self.pool.release_deferred()

    def __init__(self, store: ParamStore, *, dim: int, channels: int, dim_mults=(1, 2, 4, 8), init_kernel_size=7,
                 B: int, F: int, H: int, W: int, training: bool, out_dim: Optional[int] = None, groups: int = GROUPS,
                 use_sla: bool = True):
        self.store, self.device, self.training = store, store.flat.device, training
        self._lane, self._splitk_ws = 0, {}
        self._pending = []
        self._zeroed, self._zeroed_total = None, 0
        self.defer_joins = _lib.host_flag("VDN_DEFER_JOINS", "1") != "0"
        self.dim, self.channels, self.B, self.F, self.H, self.W = dim, channels, B, F, H, W
        self.out_dim = channels if out_dim is None else out_dim
        self.ks = init_kernel_size
        self.groups, self.use_sla = groups, use_sla
        self.pack_jobs = []
        self.extra_packers = []
        self.pool = _Pool(self.device)
        # second stream for the weight-gradient GEMMs (see side() / join()); VDN_NO_OVERLAP=1 serialises them
        self.side_stream = (torch.cuda.Stream(device=self.device)
                            if self.device.type == "cuda" and not _lib.host_flag("VDN_NO_OVERLAP") else None)
        self.side_stream2 = torch.cuda.Stream(device=self.device) if self.side_stream is not None else None
        self._gn_slots: List[torch.Tensor] = []
        self._gn_count = 0
        self._heads: List[dict] = []
        self._ss_total = 0
        self._ws = None
        self._ws_floats = 0
        n_img = B * F
        self.n_img = n_img
        td = dim * 4
        self.td = td
        dims = [dim] + [dim * m for m in dim_mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        n = len(in_out)
        self.n_res = n
        # count GroupNorm slots first (2 per ResnetBlock) so one buffer can be zeroed per forward
        n_blocks = 4 * n + 2 + 1
        self.gn_all = torch.zeros(n_blocks * 2, ops.GN_REPLICAS, B, groups, 2, dtype=F32, device=self.device)

        self.h0 = self.new((n_img, H, W, dim))
        self.init_attn = MHABlock(self, "init_temporal_attn", dim, n_img, H, W, 0)
        self.downs, self.ups = [], []
        h, w = H, W
        for l, (ci, co) in enumerate(in_out):
            blk = [ResBlock(self, f"downs.{l}.0", 1, ci, co, n_img, h, w),
                   ResBlock(self, f"downs.{l}.1", 1, co, co, n_img, h, w),
                   SLABlock(self, f"downs.{l}.2", co, n_img, h, w) if use_sla else None,
                   MHABlock(self, f"downs.{l}.3", co, n_img, h, w, 0),
                   DownConv(self, f"downs.{l}.4", co, n_img, h, w) if l < n - 1 else None]
            self.downs.append(blk)
            if l < n - 1:
                h, w = h // 2, w // 2
        mid = dims[-1]
        self.mid1 = ResBlock(self, "mid_block1", 1, mid, mid, n_img, h, w)
        self.mid_sattn = MHABlock(self, "mid_spatial_attn", mid, n_img, h, w, 1)
        self.mid_tattn = MHABlock(self, "mid_temporal_attn", mid, n_img, h, w, 0)
        self.mid2 = ResBlock(self, "mid_block2", 1, mid, mid, n_img, h, w)
        for i, (ci, co) in enumerate(reversed(in_out)):
            blk = [ResBlock(self, f"ups.{i}.0", 2, co, ci, n_img, h, w),
                   ResBlock(self, f"ups.{i}.1", 1, ci, ci, n_img, h, w),
                   SLABlock(self, f"ups.{i}.2", ci, n_img, h, w) if use_sla else None,
                   MHABlock(self, f"ups.{i}.3", ci, n_img, h, w, 0),
                   UpConv(self, f"ups.{i}.4", ci, n_img, h, w) if i < n - 1 else None]
            self.ups.append(blk)
            if i < n - 1:
                h, w = h * 2, w * 2
        self.final_block = ResBlock(self, "final_conv.0", 2, dim, dim, n_img, H, W, time=False)
        self.out = self.new((B, F, H, W, self.out_dim), F32)

        # time embedding buffers + head table
        st = store
        self.emb = self.new((B, dim), F32)
        self.h1 = self.new((B, td), F32)
        self.t_emb = self.new((B, td), F32)
        self.ss = self.new((B, max(self._ss_total, 4)), F32)
        self.e_pre = self.new((B, max(self._ss_total, 4)), F32)
        if training:
            self.dss = torch.zeros((B, max(self._ss_total, 4)), dtype=F32, device=self.device)
            self.de_ws = self.new((B, max(self._ss_total, 4)), F32)
            self.dt = self.new((B, td), F32)
            self.dh1_ws = self.new((B, td), F32)
        entries = []
        for hd in self._heads:
            p = hd["prefix"]
            e = dict(w=st.view(p + ".mlp.kernel"), b=st.view(p + ".mlp.bias"), ln_g=st.view(p + ".norm_1.scale"),
                     ln_b=st.view(p + ".norm_1.bias"), n_out=2 * hd["cout"], off=hd["off"])
            if training:
                e.update(dw=st.gview(p + ".mlp.kernel"), db=st.gview(p + ".mlp.bias"),
                         dln_g=st.gview(p + ".norm_1.scale"), dln_b=st.gview(p + ".norm_1.bias"))
            entries.append(e)
        self.head_table = ops.make_time_head_table(entries, self.device)
        self.n_heads = len(entries)
        self.x_in = None
        self.pack_table, self.n_pack_jobs, self.pack_total = ops.make_pack_table(self.pack_jobs, self.device)
        self.packed_version = -1
        self.repack()

    # -- allocation helpers ------------------------------------------------------------------
    def new(self, shape, dtype=BF16):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def new_gn_sums(self):
        t = self.gn_all[self._gn_count]
        self._gn_count += 1
        return t

    def shared_ws(self, n_floats):
        if self._ws is None or n_floats > self._ws_floats:
            self._ws_floats = max(n_floats, self._ws_floats)
            self._ws = torch.empty(self._ws_floats, dtype=F32, device=self.device)
            self._ws_owner = True
        return _WsRef(self)

    def register_time_head(self, prefix, cout):
        off = self._ss_total
        self._heads.append(dict(prefix=prefix, cout=cout, off=off))
        self._ss_total += 2 * cout
        return off

    def repack(self):
        """Refresh every packed bf16 GEMM operand from the fp32 master weights (one batched launch)."""
        ops.pack_batched(self.pack_table, self.n_pack_jobs, self.pack_total)
        for f in self.extra_packers:
            f()
        self.packed_version = self.store.version

    def sync_weights(self) -> bool:
        """Repack if the master weights changed since this engine's operands were packed (another engine's optimizer
        step, a state upload). Call before launching a forward eagerly or replaying a captured graph of one."""
        if self.packed_version != self.store.version:
            self.repack()
            return True
        return False

    # -- forward -----------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """x fp32 (B,C,F,H,W), time int32 (B,) -> fp32 (B,F,H,W,C) (unet3d.py:262-387)."""
        st = self.store
        B, Fr, H, W = self.B, self.F, self.H, self.W
        self.x_in = x
        if not torch.cuda.is_current_stream_capturing():
            self.sync_weights()
        self.gn_all.zero_()

        def time_path():  # latency-bound tiny kernels: overlapped with the init conv / init attention
            ops.time_mlp_fwd(time, st.view("time_mlp.1.kernel"), st.view("time_mlp.1.bias"),
                             st.view("time_mlp.3.kernel"), st.view("time_mlp.3.bias"), self.emb, self.h1, self.t_emb,
                             B, self.dim)
            ops.time_heads_fwd(self.t_emb, self.head_table, self.n_heads, self.e_pre, self.ss, B, self.td)

        ht = self.side(time_path)
        ops.init_conv_fwd(x, st.view("init_conv.kernel"), st.view("init_conv.bias"), self.h0, B, self.channels, Fr, H,
                          W, self.dim, self.ks)
        h = self.init_attn.forward(self.h0)
        r = h
        self.join(ht)
        skips = []
        for b1, b2, sla, mha, down in self.downs:
            h = b1.forward([h])
            h = b2.forward([h])
            if sla is not None:
                h = sla.forward(h)
            h = mha.forward(h)
            skips.append(h)
            if down is not None:
                h = down.forward(h)
        h = self.mid1.forward([h])
        h = self.mid_sattn.forward(h)
        h = self.mid_tattn.forward(h)
        h = self.mid2.forward([h])
        for b1, b2, sla, mha, up in self.ups:
            h = b1.forward([h, skips.pop()])
            h = b2.forward([h])
            if sla is not None:
                h = sla.forward(h)
            h = mha.forward(h)
            if up is not None:
                h = up.forward(h)
        h = self.final_block.forward([h, r])
        self.h_last = h
        P = self.n_img * H * W
        ops.final_conv_fwd(h, st.view("final_conv.1.kernel"), st.view("final_conv.1.bias"), self.out, P, self.dim,
                           self.out_dim)
        return self.out

    # -- backward ----------------------------------------------------------------------------
    # Gradient completion order (reverse execution): final -> ups.{n-1..0} -> mid -> downs.{n-1..0} -> late
    # ("late" = init conv / init attention / time MLP / all time heads, finished by the last stage).
    # The flat parameter layout is [late | downs.0.. | mid | ups.0.. | final], so each stage completes one
    # contiguous slice of store.grad; trainer.py reduces those slices while later stages still run.
    def backward_stages(self):
        """List of (stage_name, fn). fn() enqueues that stage's kernels; stages must run in order."""
        assert self.training
        st, pool = self.store, self.pool
        P = self.n_img * self.H * self.W
        S = {}
        n = self.n_res

        def step(mod_backward, *a, **k):
            d2 = mod_backward(S["d"], *a, **k)
            if isinstance(d2, (list, tuple)):
                (d2,) = d2
            pool.put(S["d"])
            S["d"] = d2

        def st_final():
            self.dss.zero_()
            if self._zeroed_total:
                self.zeroed(0, 1)  # allocate on first use
                self._zeroed.zero_()
            dh = pool.get(self.h_last.shape)
            ops.final_conv_bwd(self.h_last, S["dout"], st.view("final_conv.1.kernel"), dh,
                               st.gview("final_conv.1.kernel"), st.gview("final_conv.1.bias"), P, self.dim, self.out_dim)
            S["d"], S["dr"] = self.final_block.backward(dh)
            pool.put(dh)
            S["dskips"] = [None] * n

        def make_up(i):
            def f():
                b1, b2, sla, mha, up = self.ups[i]
                if up is not None:
                    step(up.backward)
                step(mha.backward)
                if sla is not None:
                    step(sla.backward)
                step(b2.backward)
                d2, dsk = b1.backward(S["d"])
                pool.put(S["d"])
                S["d"] = d2
                S["dskips"][n - 1 - i] = dsk
            return f

        def st_mid():
            step(self.mid2.backward)
            step(self.mid_tattn.backward)
            step(self.mid_sattn.backward)
            step(self.mid1.backward, extra=S["dskips"][n - 1])
            pool.put(S["dskips"][n - 1])

        def make_down(l):
            def f():
                b1, b2, sla, mha, down = self.downs[l]
                if down is not None:
                    acc = down.backward(S["d"], S["dskips"][l])
                    pool.put(S["d"])
                    S["d"] = acc
                step(mha.backward)
                if sla is not None:
                    step(sla.backward)
                step(b2.backward)
                step(b1.backward, extra=S["dr"] if l == 0 else None)
            return f

        def st_late():
            pool.put(S["dr"])

            def time_path_bwd():  # dss is complete once downs.0 has run: overlap with the init attention backward
                ops.time_heads_bwd(self.t_emb, self.head_table, self.n_heads, self.e_pre, self.dss, self.de_ws,
                                   self.dt, self.B, self.td)
                ops.time_mlp_bwd(self.dt, self.emb, self.h1, st.view("time_mlp.3.kernel"),
                                 st.gview("time_mlp.1.kernel"), st.gview("time_mlp.1.bias"),
                                 st.gview("time_mlp.3.kernel"), st.gview("time_mlp.3.bias"), self.dh1_ws, self.B,
                                 self.dim)

            ht = self.side(time_path_bwd)
            step(self.init_attn.backward)
            ops.init_conv_wgrad(self.x_in, S["d"], st.gview("init_conv.kernel"), st.gview("init_conv.bias"), self.B,
                                self.channels, self.F, self.H, self.W, self.dim, self.ks)
            pool.put(S["d"])
            self.join(ht)

        stages = [("final", st_final)]
        stages += [(f"ups.{i}", make_up(i)) for i in reversed(range(n))]
        stages += [("mid", st_mid)]
        stages += [(f"downs.{l}", make_down(l)) for l in reversed(range(n))]
        stages += [("late", st_late)]
        def staged(fn):
            def run():
                self.begin_stage()
                try:
                    fn()
                finally:
                    self.end_stage()
            return run

        stages = [(nm, staged(fn)) for nm, fn in stages]
        self._bw_state = S
        return stages

    def backward(self, dout: torch.Tensor) -> None:
        """dout fp32 (B,F,H,W,C): accumulates every parameter gradient into store.grad (+=)."""
        if not hasattr(self, "_stages"):
            self._stages = self.backward_stages()
        self._bw_state["dout"] = dout
        for _, fn in self._stages:
            fn()

    def grad_slices(self):
        """{stage_name: (begin, end)} offsets into store.grad completed by each backward stage."""
        offs = self.store.offsets
        names = [nm for nm, _ in self.store.spec if not is_late_param(nm)]

        def first(prefix):
            for nm in names:
                if nm.startswith(prefix):
                    return offs[nm][0]
            raise KeyError(prefix)

        n = self.n_res
        marks = [("late", 0)]
        marks += [(f"downs.{l}", first(f"downs.{l}.")) for l in range(n)]
        marks += [("mid", first("mid_block1."))]
        marks += [(f"ups.{i}", first(f"ups.{i}.")) for i in range(n)]
        marks += [("final", first("final_conv."))]
        out = {}
        for j, (nm, b) in enumerate(marks):
            e = marks[j + 1][1] if j + 1 < len(marks) else self.store.total
            out[nm] = (b, e)
        return out


class _WsRef:
    """Late-bound reference to the engine's shared workspace (it may grow while the plan is built)."""

    def __init__(self, eng):
        self.eng = eng

    def data_ptr(self):
        return self.eng._ws.data_ptr()
