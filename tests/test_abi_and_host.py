"""CPU-side checks: the C-ABI library loads and exports every symbol include/vdn.h declares (no compute
calls), and the host logic (parameter mapping, LR schedule, keys, bucket planning) behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vdn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vdn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, "video_diffusion_nnx_b200", "libvdn.so")
    if not os.path.exists(lib_path):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(lib_path)
    syms = _declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.vdn_version.restype = ctypes.c_int
    assert lib.vdn_version() >= 100


def test_error_path_without_gpu_reports_a_message():
    """Argument validation happens before any CUDA call: a bad descriptor returns VDN_E_SHAPE."""
    from video_diffusion_nnx_b200 import _lib

    d = _lib.TapGemmDesc()
    d.kind = 7
    rc = _lib.lib.vdn_tapgemm(ctypes.byref(d), None, None, None, None, None, None, None, None, None, None)
    assert rc == -1
    assert b"kind" in _lib.lib.vdn_last_error()


def test_state_dict_mapping_round_trip_on_host():
    from video_diffusion_nnx_b200.unet3d import Unet3D

    net = Unet3D(dim=32, channels=1, rngs=0)
    shapes = net.reference_param_shapes()
    assert sum(int(np.prod(s)) for s in shapes.values()) == 9_993_409
    internal = dict(net.spec)
    seen = {}
    for k in shapes:
        name, qi = net._to_internal(k)
        assert name in internal, (k, name)
        seen.setdefault(name, []).append(qi)
    for name, qs in seen.items():
        assert qs == [None] or sorted(qs) == [0, 1, 2], (name, qs)
    assert set(seen) == set(internal)
    # flax init distributions: zero biases, unit scales, lecun-normal kernels
    st = net.state_dict()
    assert np.all(st["init_conv.bias"] == 0) and np.all(st["downs.0.0.norm_2.scale"] == 1)
    k = st["mid_block1.block_1.proj.kernel"]
    assert abs(k.std() - (1.0 / (9 * 256)) ** 0.5) / (1.0 / (9 * 256)) ** 0.5 < 0.05


def test_unsupported_options_raise():
    from video_diffusion_nnx_b200.unet3d import Unet3D

    with pytest.raises(NotImplementedError):
        Unet3D(dim=32, cond_dim=16)
    with pytest.raises(NotImplementedError):
        Unet3D(dim=32, use_bert_text_cond=True)


def test_piecewise_cosine_lr_matches_optax_definition():
    from video_diffusion_nnx_b200.trainer import piecewise_cosine_lr as lr

    assert lr(0, 1e-4, 20000, 80000, 0.1) == 1e-4
    assert lr(20000, 1e-4, 20000, 80000, 0.1) == 1e-4
    assert abs(lr(60000, 1e-4, 20000, 80000, 0.1) - (1e-5 + (1e-4 - 1e-5) / 2)) < 1e-12
    assert abs(lr(100000, 1e-4, 20000, 80000, 0.1) - 1e-5) < 1e-15
    assert abs(lr(10 ** 6, 1e-4, 20000, 80000, 0.1) - 1e-5) < 1e-15
    assert lr(5, 1e-4, 0, 0, 1.0) == 1e-4  # Trainer defaults (trainer.py:116-118)


def test_bucket_plan_covers_the_flat_gradient_exactly_once():
    from video_diffusion_nnx_b200.trainer import plan_buckets

    slices = {"late": (0, 100), "downs.0": (100, 300), "downs.1": (300, 700), "mid": (700, 1500),
              "ups.0": (1500, 2500), "ups.1": (2500, 2600), "final": (2600, 2700)}
    order = ["final", "ups.1", "ups.0", "mid", "downs.1", "downs.0", "late"]
    segs = plan_buckets(order, slices, bucket_elems=900)
    cover = sorted(r for _, r in segs)
    assert cover[0][0] == 0 and cover[-1][1] == 2700
    assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert [n for names, _ in segs for n in names] == order
    assert segs[-1][0][-1] == "late"


def test_keys_are_deterministic_and_distinct():
    from video_diffusion_nnx_b200.gaussian_diffusion import Key

    a, b, c = Key(5).split(3)
    assert (a.seed, a.stream) != (b.seed, b.stream) != (c.seed, c.stream)
    a2, _, _ = Key(5).split(3)
    assert (a.seed, a.stream) == (a2.seed, a2.stream)


def test_checkpoint_bridge_round_trip(tmp_path):
    """{'model', 'ema_params'} tree (utils.py:445-455) <-> flat .npz <-> Unet3D.load_state_dict, on the host."""
    from video_diffusion_nnx_b200 import checkpoint as ck
    from video_diffusion_nnx_b200.unet3d import Unet3D

    net = Unet3D(dim=32, channels=1, rngs=1)
    model = net.state_dict()
    ema = {k: v + 1.0 for k, v in model.items()}
    path = str(tmp_path / "ckpt_7.npz")
    ck.save_checkpoint(path, model, ema, step=7)
    got, step = ck.load_checkpoint(path)
    got_ema, _ = ck.load_checkpoint(path, load_ema_params=True)
    assert step == 7 and set(got) == set(model)
    assert all(np.array_equal(got[k], model[k]) for k in model)
    assert all(np.array_equal(got_ema[k], ema[k]) for k in model)
    # nested (orbax / nnx.State.to_pure_dict) form and back
    nested = ck.nest(model)
    assert "proj" in nested["downs"]["0"]["0"]["block_1"] and ck.unnest(nested).keys() == model.keys()
    other = Unet3D(dim=32, channels=1, rngs=2)
    other.load_state_dict(got_ema)
    assert np.array_equal(other.state_dict()["init_conv.bias"], ema["init_conv.bias"])


def _mnist_file(tmp_path, frames=15, seqs=5, hw=32):
    rng = np.random.default_rng(0)
    path = str(tmp_path / "mnist.npy")
    np.save(path, rng.integers(0, 256, (frames, seqs, hw, hw)).astype(np.uint8))
    return path


def test_moving_mnist_matches_reference_behaviour(tmp_path):
    """test_datasets.py:35-103 carried over: length, (c, f, h, w) float32 items at the ORIGINAL size (the transform is
    never applied), zero padding / truncation / pass-through of the frame axis, raw 0..255 values (no rescaling)."""
    from video_diffusion_nnx_b200.data import MovingMNIST

    path = _mnist_file(tmp_path)
    raw = np.load(path)
    ds = MovingMNIST(file_path=path, image_size=64, num_frames=20, channels=1, force_num_frames=True)
    assert len(ds) == 5 and ds.image_size == 64 and ds.channnels == 1
    assert ds.cast_num_frames_fn.keywords["frames"] == 20
    item = ds[0]
    assert isinstance(item, np.ndarray) and item.shape == (1, 20, 32, 32) and item.dtype == np.float32
    assert np.array_equal(item[0, :15], raw[:, 0].astype(np.float32)) and not item[0, 15:].any()
    assert item.max() > 1.0  # raw pixel values: the reference does not divide by 255
    assert MovingMNIST(path, 64, num_frames=25)[0].shape[1] == 25
    assert MovingMNIST(path, 64, num_frames=10)[3].shape[1] == 10
    assert np.array_equal(MovingMNIST(path, 64, num_frames=10)[3][0], raw[:10, 3].astype(np.float32))
    assert MovingMNIST(path, 64, num_frames=20, force_num_frames=False)[0].shape[1] == 15


def test_training_batches_shard_the_global_batch(tmp_path):
    """Two ranks with the same seed see disjoint halves of the same shuffled global batches (trainer.py:307-309)."""
    from video_diffusion_nnx_b200.data import MovingMNIST, training_batches

    ds = MovingMNIST(_mnist_file(tmp_path, seqs=12), 32, num_frames=4)
    it0 = training_batches(ds, 2, seed=5, rank=0, world=2, device=None)
    it1 = training_batches(ds, 2, seed=5, rank=1, world=2, device=None)
    whole = training_batches(ds, 4, seed=5, rank=0, world=1, device=None)
    for _ in range(7):  # crosses an epoch boundary (3 global batches per epoch)
        a, b, w = next(it0), next(it1), next(whole)
        assert a.shape == (2, 1, 4, 32, 32) and torch_equal(torch_cat(a, b), w)


def torch_cat(a, b):
    import torch

    return torch.cat([a, b])


def torch_equal(a, b):
    import torch

    return torch.equal(a, b)


def test_checkpoint_tree_is_the_diffusion_state(tmp_path):
    """The reference checkpoints nnx.split(GaussianDiffusion) (trainer.py:136, 600; utils.py:486): `denoise_fn/...`
    leaves plus the ten schedule tables. Save -> load -> GaussianDiffusion.load_state_dict round trip on the host,
    including changed (trained) tables and the '/value' leaf spelling."""
    from video_diffusion_nnx_b200 import checkpoint as ck
    from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion
    from video_diffusion_nnx_b200.unet3d import Unet3D

    net = Unet3D(dim=32, channels=1, rngs=1)
    gd = GaussianDiffusion(net, image_size=64, num_frames=2, channels=1, timesteps=50, loss_type="l2")
    state = gd.state_dict()
    assert len(state) == len(net.reference_param_shapes()) + 10
    assert "denoise_fn.downs.0.0.block_1.proj.kernel" in state and state["sqrt_alphas_cumprod"].shape == (50,)
    trained = dict(state)
    trained["sqrt_alphas_cumprod"] = state["sqrt_alphas_cumprod"] * 0.5  # as if Adam had moved the table (C9)
    ema = {k: v + 1.0 for k, v in trained.items()}
    for suffix in (False, True):
        path = str(tmp_path / f"ck{int(suffix)}.npz")
        ck.save_checkpoint(path, trained, ema, step=3, value_suffix=suffix)
        with np.load(path) as z:
            keys = set(z.files)
        want = "model/denoise_fn/init_conv/kernel" + ("/value" if suffix else "")
        assert want in keys and ("ema_params/posterior_variance" + ("/value" if suffix else "")) in keys
        got, step = ck.load_checkpoint(path)
        got_ema, _ = ck.load_checkpoint(path, load_ema_params=True)
        assert step == 3 and set(got) == set(trained)
        net2 = Unet3D(dim=32, channels=1, rngs=9)
        gd2 = GaussianDiffusion(net2, image_size=64, num_frames=2, channels=1, timesteps=50, loss_type="l2")
        gd2.load_state_dict(got_ema)
        back = gd2.state_dict()
        assert all(np.array_equal(back[k], ema[k]) for k in ema)
        net2.load_state_dict(got)  # a bare Unet3D accepts the prefixed tree and ignores the tables
        assert np.array_equal(net2.state_dict()["init_conv.kernel"], trained["denoise_fn.init_conv.kernel"])


def test_ffi_shim_builds_and_names_real_entry_points():
    """ffi/vdn_ffi.cc compiles everywhere (stub form without jaxlib's headers); every FFI target it announces maps to
    a C-ABI entry point that include/vdn.h declares and libvdn.so exports."""
    from video_diffusion_nnx_b200 import _lib, jax_ffi

    assert os.path.exists(jax_ffi.FFI_LIB_PATH), "run build() first"
    lib = ctypes.CDLL(jax_ffi.FFI_LIB_PATH)
    names = jax_ffi._targets(lib)
    assert len(names) >= 35 and len(set(names)) == len(names)
    alias = {"vdn_wgrad": "vdn_wgrad_bias", "vdn_mha_temporal_fwd": "vdn_mha_temporal_tc_fwd",
             "vdn_mha_temporal_bwd": "vdn_mha_temporal_tc_bwd"}
    for n in names:
        assert alias.get(n, n) in _lib.PROTOTYPES, n
    src = open(os.path.join(ROOT, "ffi", "vdn_ffi.cc")).read()
    for n in names:
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({n}_ffi" in src, n
    lib.vdn_ffi_available.restype = ctypes.c_int
    if not lib.vdn_ffi_available():  # no jaxlib here: register() must refuse loudly, not fall back
        with pytest.raises(RuntimeError):
            jax_ffi.register()


def test_every_prototype_has_argtypes():
    from video_diffusion_nnx_b200 import _lib

    assert len(_lib.PROTOTYPES) >= 60
    for name, (ret, argt) in _lib.PROTOTYPES.items():
        fn = getattr(_lib.lib, name)
        assert fn.argtypes is not None and list(fn.argtypes) == argt, name
    with pytest.raises(ctypes.ArgumentError):
        _lib.lib.vdn_randn(None, "not a number", 0, 0, 0, None)


def test_tapgemm_workspace_is_opt_in():
    """vdn_tapgemm_workspace is host arithmetic (no GPU): 0 for every launch by default - the cluster split-K kernel
    behind vdn_tapgemm_ws is opt-in (measured slower) - and the partial-tile scratch of the plan
    (splits x row tiles x 128 x N fp32) once it is switched on."""
    from video_diffusion_nnx_b200 import ops
    from video_diffusion_nnx_b200._lib import debug_switches

    args = (ops.VDN_TAP_UNIT, 40, 8, 8, 1, 256, ops.TAPS_3x3, 256)  # the (1,3,3) 256 -> 256 conv of the 8x8 level
    assert ops.tapgemm_workspace_bytes(*args) == 0
    with debug_switches(VDN_SPLITK=1):
        nbytes = ops.tapgemm_workspace_bytes(*args)
        assert nbytes == 3 * 20 * 128 * 256 * 4  # 20 row tiles x two 128-column tiles -> 3 K ranges per tile
        # launches that fill the SMs on their own, or have too few K steps to split, never ask for scratch
        assert ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, 40, 32, 32, 1, 64, ops.TAPS_3x3, 64) == 0
        assert ops.tapgemm_workspace_bytes(ops.VDN_TAP_UNIT, 40, 8, 8, 1, 256, ops.TAPS_1x1, 768) == 0
    assert ops.tapgemm_workspace_bytes(*args) == 0
