#!/usr/bin/env python
"""Benchmark of the B200-native Unet3D / GaussianDiffusion hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Default workload: configs/config_v2_2.yaml data-parallel TRAINING step (Unet3D dim 32, 1 channel,
10 frames, 64x64, T=1000, L2 loss, Adam + EMA), per-GPU batch 4 (the config's train_batch_size),
synthetic clips, random-init weights, bf16 tensor-core compute with fp32 master weights.
One JSON line is printed by rank 0. `value` = clips/s with inputs resident in HBM (device-timed,
max over ranks); `e2e` = the same step through the public TrainStep.step API with the batch coming
from pinned host memory and the loss read back every step. The sampling metric (frames/s of the
full T-step p_sample_loop) is reported under "sampling".

--impl reference times the reference's algorithm on the host CPU cores (the oracle port of the JAX
code: JAX/flax are not installable in this image), one clip per step as a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(dim=32, channels=1, frames=10, size=64, timesteps=1000, loss="l2", lr=1e-4, lr_decay_start_step=20000,
           lr_decay_steps=80000, lr_decay_coeff=0.1, step_start_ema=2000, update_ema_every=10, ema_decay=0.9999,
           per_gpu_batch=4)
TRAIN_GFLOP_PER_CLIP = 159.3   # 3 x 53.1 GFLOP forward (SURVEY.md 8d / BASELINE.md section 4)
FWD_GFLOP_PER_CLIP = 53.1
WORKLOAD_NAME = "configs/config_v2_2.yaml"


def select_workload(name):
    """v2_2 is the headline workload (BASELINE.json configs[1]); v2_3x is BASELINE.json configs[3]: config_v2_3
    scaled up to 16 frames at 128x128, dim 128 (the conv- and temporal-attention-bound case)."""
    global TRAIN_GFLOP_PER_CLIP, FWD_GFLOP_PER_CLIP, WORKLOAD_NAME
    if name == "v2_3x":
        CFG.update(dim=128, frames=16, size=128)
        TRAIN_GFLOP_PER_CLIP, FWD_GFLOP_PER_CLIP = 9954.0, 3318.0
        WORKLOAD_NAME = "configs/config_v2_3.yaml scaled up (16 frames, 128x128, dim 128)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs: NVML polled every 5 ms (nvidia_ml_py), or
    nvidia-smi every 0.2 s when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index=0):
        self.idx, self.rows, self._stop, self._th = gpu_index, [], threading.Event(), None
        self.sm, self.mx, self.reasons, self.how = [], None, set(), "nvidia-smi"
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].strip().isdigit() else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml, self.how = pynvml, "nvml"
        except Exception:
            self._nvml = None

    def _run_nvml(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for n, bit in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.005)

    def _run_smi(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    r = [c.strip() for c in out.split(",")]
                    if r[0].replace(".", "").isdigit():
                        self.sm.append(float(r[0]))
                    if len(r) > 1 and r[1].replace(".", "").isdigit():
                        self.mx = float(r[1])
                    for n, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(sm), "how": self.how}


def _cpu_train_sampler(batch, steps, warmup, sample_steps):
    """The reference's algorithm on the host cores (oracle port of the JAX code; JAX is not installable here):
    `steps` timed training steps of `batch` clips = p_losses value_and_grad + optax-style Adam + the EMA rule
    (trainer.py:337-382), then `sample_steps` timed p_sample steps of one clip (gaussian_diffusion.py:231-261).
    Returns (clips/s, s per training step, frames/s of a full T-step loop, cores)."""
    import torch

    from oracle import diffusion_oracle as D
    from oracle import unet3d_oracle as U

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    p = U.init_params(CFG["dim"], CFG["channels"])
    for v in p.values():
        v.requires_grad_(True)
    gd = D.GaussianDiffusionOracle(lambda xx, tt: U.unet3d_forward(p, xx, tt, CFG["dim"]), image_size=CFG["size"],
                                   num_frames=CFG["frames"], channels=CFG["channels"], timesteps=CFG["timesteps"],
                                   loss_type=CFG["loss"])
    opt = torch.optim.Adam(list(p.values()), lr=CFG["lr"], betas=(0.9, 0.999), eps=1e-8)
    ema = {k: v.detach().clone() for k, v in p.items()}
    x = torch.rand(batch, 1, CFG["frames"], CFG["size"], CFG["size"])
    n = [0]

    def step():
        opt.zero_grad()
        t = torch.randint(0, CFG["timesteps"], (batch,), dtype=torch.int32)
        loss = gd(x, t, torch.randn_like(x))
        loss.backward()
        for v in p.values():
            if v.grad is None:
                v.grad = torch.zeros_like(v)
        opt.step()
        n[0] += 1
        if n[0] % CFG["update_ema_every"] == 0:
            with torch.no_grad():
                for k in ema:
                    ema[k].mul_(CFG["ema_decay"]).add_(p[k].detach(), alpha=1 - CFG["ema_decay"])
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    fps = None
    if sample_steps > 0:
        with torch.no_grad():
            img = torch.randn(1, 1, CFG["frames"], CFG["size"], CFG["size"])
            tt = torch.full((1,), CFG["timesteps"] - 1, dtype=torch.int32)
            gd.p_sample(img, tt, torch.randn_like(img))
            t1 = time.perf_counter()
            for _ in range(sample_steps):
                img = gd.p_sample(img, tt, torch.randn_like(img))
            ds = (time.perf_counter() - t1) / sample_steps
        fps = CFG["frames"] / (ds * CFG["timesteps"])
    return batch / dt, dt, fps, cores


def run_reference(args):
    """--impl reference: the reference's own algorithm on the box's host cores (oracle port; rank 0 alone)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)
    B = CFG["per_gpu_batch"]
    val, dt, fps, cores = _cpu_train_sampler(B, args.steps, args.warmup, 2)
    sample = (f"{args.steps} steps x {B} clips (one rank's shard of the workload), p_losses fwd+bwd + Adam + EMA, torch fp32 "
              f"oracle port of the JAX reference on {cores} host threads")
    line = {"impl": "reference", "metric": metric_name("v2_2"), "value": val, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, args.gpus),
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample,
                             "sampling": {"value": fps, "unit": "frames/s",
                                          "sample": "2 p_sample steps of 1 clip, extrapolated to the T=1000 loop"}},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "JAX / flax are not installable in this image (DESIGN.md section 6): the arm is the CPU restatement of the "
                    "reference (kind = port), each step a bounded sample of the workload: one rank's batch of "
                    f"{B} clips with the optimizer"}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess():
    """The CPU baseline leg, run in a fresh process with a clean threading environment BEFORE any process group
    exists (torchrun exports OMP_NUM_THREADS=1, and ranks waiting in a NCCL barrier would spin on their GPUs)."""
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "WORLD_SIZE",
                                                            "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-worker"], env=env,
                             capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
        return json.loads(out[-1])
    except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
        return {"value": None, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}


def cpu_baseline_worker():
    B = CFG["per_gpu_batch"]
    val, dt, fps, cores = _cpu_train_sampler(B, 3, 1, 2)
    print(json.dumps({"value": val, "unit": "clips/s", "cores": cores, "kind": "port",
                      "sample": f"3 training steps of {B} clips after 1 warm-up (~{4 * dt:.0f} s of CPU work): p_losses "
                                "fwd+bwd + Adam + EMA, torch fp32 oracle port of the JAX reference on all host cores",
                      "sampling": {"value": fps, "unit": "frames/s",
                                   "sample": "2 p_sample steps of 1 clip, extrapolated to the T=1000 loop"}}), flush=True)


def metric_name(workload):
    return f"train clips/sec (Unet3D config_{workload} p_losses fwd+bwd+allreduce+Adam/EMA, device-timed)"


def workload_config(B, world):
    return {"workload": f"{WORKLOAD_NAME} training step: Unet3D dim {CFG['dim']}, 1 ch, {CFG['frames']} frames, "
                        f"{CFG['size']}x{CFG['size']}, T=1000, L2, Adam+EMA; per-GPU batch {B}, global batch {B * world}",
            "parallelism": f"dp{world}", "global_batch": B * world,
            "l2": "no explicit flush: one step streams > 2 GB of activations / saved tensors, far beyond the 126 MB L2",
            "timing": "CUDA events on the launch stream around K CUDA-graph-replayed steps, max over ranks"}


def conv_roofline(torch, ops, pk, level=0, with_traffic=False):
    """Times one (1,3,3) implicit-GEMM conv + bias + GroupNorm partial sums of the step at resolution level `level`
    (0: 64x64 / dim channels ... 3: 8x8 / 8*dim channels for config_v2_2) in isolation with CUDA events, replayed from a
    CUDA graph that rotates over enough distinct (input, weights, output) sets to exceed the 126 MB L2."""
    dev = "cuda"
    n_img = CFG["per_gpu_batch"] * CFG["frames"]
    H = W = CFG["size"] >> level
    C = CFG["dim"] << level
    per_set = n_img * H * W * C * 2 * 2 + 9 * C * C * 2
    nbuf = max(4, min(128, int(260e6 // per_set) + 1))  # > 2 x L2 in rotation
    xs = [torch.randn(n_img, H, W, C, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    outs = [torch.empty(n_img, H, W, C, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    wps = []
    for _ in range(nbuf):
        w = torch.randn(9, C, C, device=dev) * (9 * C) ** -0.5
        wp = torch.empty(C, 9 * C, dtype=torch.bfloat16, device=dev)
        ops.pack_weight(w, wp, 9, C, C, 0)
        wps.append(wp)
    bias = torch.zeros(C, device=dev)
    sums = torch.zeros(ops.GN_REPLICAS, CFG["per_gpu_batch"], 8, 2, device=dev)
    rows = CFG["frames"] * H * W

    def run(i):
        ops.tapgemm(ops.VDN_TAP_UNIT, [xs[i % nbuf]], wps[i % nbuf], ops.TAPS_3x3, bias=bias, out=outs[i % nbuf],
                    gn_sums=sums, gn_groups=8, rows_per_sample=rows)

    n0 = ops.lib.vdn_launch_count()
    run(0)
    assert ops.lib.vdn_launch_count() == n0 + 1
    # the launches are replayed from a CUDA graph: per-call host work (ctypes) is longer than the kernel
    us = _graph_time_us(torch, run, n=max(64, nbuf), warm=8)
    ms = us * 1e-3
    M = n_img * H * W
    flops = 2.0 * M * C * 9 * C
    bytes_alg = 2.0 * M * C * 2 + 9 * C * C * 2
    tf = flops / (ms * 1e-3) / 1e12
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    # below the ridge (AI = flops/bytes < ~250 FLOP/B: the dim-32 level, AI 144) HBM is the binding roofline; the
    # small-M levels (AI 570-1000) and every level of the v2_3x workload are bound by the tensor pipe
    ridge = pk["tf_burst"] * 1e3 / pk["hbm"]
    tensor_bound = flops / bytes_alg > ridge
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if with_traffic and os.path.exists(tp):  # dram bytes per launch of this kernel from the committed `ncu --set full` capture
        td = json.load(open(tp))
        key = f"conv_l{level}_c{C}"
        if key in td:
            traffic = td[key]["dram_bytes_read"] + td[key]["dram_bytes_write"]
    kern = {32: "conv3x3_rows_kernel<32,32,1>", 64: "conv3x3_rows_kernel<64,64,1>"}.get(C, "tapgemm_kernel<64,8>")
    if CFG["dim"] >= 128:
        kern = "conv3x3_slab_kernel<32>"
    name = f"conv(1,3,3) {C}->{C} @{H}x{W} (M={M},N={C},K={9 * C}) {kern}"
    r = {"kernel": name, "bound": "tensor" if tensor_bound else "hbm"}
    if tensor_bound:
        r.update(achieved=tf, peak=pk["tf_burst"], unit="TFLOP/s", frac=tf / pk["tf_burst"], traffic=traffic,
                 algorithmic_flops=flops, algorithmic_bytes=bytes_alg, hbm_gbs=gbs)
    else:
        r.update(achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], traffic=traffic,
                 algorithmic_bytes=bytes_alg, tensor_tflops=tf, tensor_frac_of_burst=tf / pk["tf_burst"])
    r.update(us_per_launch=ms * 1e3, peak_source=pk["src"], buffers_in_rotation=nbuf,
             l2="inputs, weights and outputs rotate over sets larger than L2 between launches")
    return r


def wgrad_roofline(torch, ops, pk):
    """The step's dominant kernel by the committed launch list (profiles/r2_final_launches_train_v2_2_b4.summary.txt:
    wgrad_kernel<0>, 90 of 486 launches = 20 % of the kernel time) on the instance that takes most of its time: the weight
    (+ bias) gradient of a q|k|v projection at the full-resolution level, dW[C][768] = x^T dqkv over all P pixels -
    one pass over the 768-channel gradient tensor, HBM bound. Timed live, dqkv rotating over sets larger than L2."""
    dev, bf = "cuda", torch.bfloat16
    B, Fr, S, C = CFG["per_gpu_batch"], CFG["frames"], CFG["size"], CFG["dim"]
    n_img, P = B * Fr, B * Fr * S * S
    nb = max(2, int(300e6 // (P * 768 * 2)) + 1)
    x = torch.randn(n_img, S, S, C, device=dev).to(bf)
    gs = [torch.randn(n_img, S, S, 768, device=dev).to(bf) for _ in range(nb)]
    dw = torch.zeros(1, C, 768, device=dev)
    db = torch.zeros(768, device=dev)
    us = _graph_time_us(torch, lambda i: ops.wgrad(ops.VDN_TAP_UNIT, [x], gs[i % nb], dw, ops.TAPS_1x1, dbias=db), n=16)
    bytes_alg = 2.0 * P * (C + 768) + 2 * 4.0 * C * 768
    flops = 2.0 * P * C * 768
    gbs = bytes_alg / (us * 1e-6) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        td = json.load(open(tp)).get("wgrad_qkv_l0")
        if td:
            traffic = td["dram_bytes_read"] + td["dram_bytes_write"]
    return {"kernel": f"q|k|v projection weight + bias gradient {C}x768 over {P} pixels (wgrad_kernel<0>, split-K over 24 pixel ranges)",
            "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": traffic,
            "algorithmic_bytes": bytes_alg, "algorithmic_flops": flops, "tensor_tflops": flops / (us * 1e-6) / 1e12,
            "us_per_launch": us, "peak_source": pk["src"], "buffers_in_rotation": nb,
            "share_of_step_kernel_time": 0.20, "launch_list": "profiles/r2_final_launches_train_v2_2_b4.summary.txt",
            "l2": "the gradient tensors rotate over sets larger than L2 between launches",
            "note": "dominant by kernel name; the kernel FURTHEST below its roofline is the generic tap-GEMM on the "
                    "small-M convs (second by share): first entry of roofline_kernels"}


def _graph_time_us(torch, run, n=32, warm=4):
    """us per call of run(i), replayed from a CUDA graph of n calls (host work per call exceeds these kernels)."""
    for i in range(warm):
        run(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(n):
                run(i)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def extra_rooflines(torch, ops, pk):
    """Three more kernels of the step at the workload's full-resolution level, timed live like the conv above
    (rotating over buffers larger than L2): the q|k|v projection (persistent tap-GEMM, bound by its output
    write), the GroupNorm+SiLU backward pair and the temporal-attention core backward (HBM streams)."""
    dev = "cuda"
    B, Fr, S, C = CFG["per_gpu_batch"], CFG["frames"], CFG["size"], CFG["dim"]
    n_img, P = B * Fr, B * Fr * S * S
    bf = torch.bfloat16
    out = []
    nb = max(3, int(300e6 // (P * 768 * 2)) + 1)  # q|k|v tensors: > 2 x L2 in rotation

    def entry(name, us, bytes_alg, flops=0.0):
        gbs = bytes_alg / (us * 1e-6) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                "algorithmic_bytes": bytes_alg, "us_per_launch": us, "tensor_tflops": flops / (us * 1e-6) / 1e12}

    # q|k|v projection: x [P][C] -> [P][768]
    x = torch.randn(n_img, S, S, C, device=dev).to(bf)
    w = torch.randn(1, C, 768, device=dev) * C ** -0.5
    wp = torch.empty(768, C, dtype=bf, device=dev)
    ops.pack_weight(w, wp, 1, C, 768, 0)
    qkvs = [torch.empty(n_img, S, S, 768, dtype=bf, device=dev) for _ in range(nb)]
    us = _graph_time_us(torch, lambda i: ops.tapgemm(ops.VDN_TAP_UNIT, [x], wp, ops.TAPS_1x1, out=qkvs[i % nb]))
    out.append(entry(f"q|k|v projection {C}->768 (tapgemm_persist_kernel, M={P})", us, 2.0 * P * (C + 768) + 2 * C * 768,
                     2.0 * P * C * 768))
    # temporal attention core backward: reads qkv + dO + lse, writes dqkv
    for q in qkvs:
        q.normal_()
    d_o = torch.randn(P, 256, device=dev).to(bf)
    lse = torch.zeros(P, 8, device=dev)
    ops.mha_temporal_core_fwd(qkvs[0].view(P, 768), torch.empty(P, 256, dtype=bf, device=dev), lse, B, Fr, S, S)
    dq = [torch.empty(P, 768, dtype=bf, device=dev) for _ in range(2)]
    us = _graph_time_us(torch, lambda i: ops.mha_temporal_tc_bwd(qkvs[i % nb].view(P, 768), d_o, lse, dq[i % 2], B, Fr, S, S),
                        n=16)
    out.append(entry(f"temporal attention core backward (mha_temporal_mma_bwd_kernel, {B * S * S} pixels x {Fr} frames)", us,
                     2.0 * P * (768 + 256 + 768) + 4.0 * P * 8, 2.0 * P * Fr * 32 * 8 * 5))
    del qkvs, dq
    # GroupNorm + SiLU backward (reduce + apply): reads dy, x twice, writes dx
    nbn = max(3, int(300e6 // (P * C * 2 * 3)) + 1)
    xr = [torch.randn(n_img, S, S, C, device=dev).to(bf) for _ in range(nbn)]
    dy = [torch.randn(n_img, S, S, C, device=dev).to(bf) for _ in range(nbn)]
    dx = [torch.empty(n_img, S, S, C, dtype=bf, device=dev) for _ in range(nbn)]
    rows = Fr * S * S
    sums = torch.zeros(ops.GN_REPLICAS, B, 8, 2, device=dev)
    sums[0, :, :, 1] = rows * (C // 8)  # unit variance, zero mean
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    T = torch.zeros(B, C, 2, device=dev)
    dg, db, dcb = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    us = _graph_time_us(torch, lambda i: ops.gn_silu_bwd(dy[i % nbn], xr[i % nbn], sums, gamma, beta, None, T, dx[i % nbn], dg, db,
                                                         None, B, rows, C, dconv_bias=dcb))
    out.append(entry(f"GroupNorm+SiLU backward, 2 passes (gn_bwd_reduce_kernel + gn_bwd_apply_kernel, {P} pixels x {C} ch)", us,
                     2.0 * P * C * 5))
    return out


def parity_block(torch, ops, gd, net, B):
    """Loss and predicted noise of the benchmarked model (the weights the timed steps left behind) against the CPU
    oracle on one batch of the benchmark's own shape, in the same job. bf16 = the timed tensor-core path; fp32 = the
    fp32-grade operand path (bf16 x 3 split operands, fp32 activations; gate 1e-3, BASELINE.json north_star)."""
    import numpy as np

    from oracle import diffusion_oracle as D
    from oracle import unet3d_oracle as U

    rng = np.random.default_rng(99)
    shape = (B, CFG["channels"], CFG["frames"], CFG["size"], CFG["size"])
    x = torch.from_numpy(rng.random(shape, dtype=np.float32))
    t = torch.from_numpy(rng.integers(0, CFG["timesteps"], (B,)).astype(np.int32))
    noise = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
    p = {k: torch.from_numpy(v) for k, v in net.state_dict().items()}
    cap = {}

    def fwd(xx, tt):
        cap["eps"] = U.unet3d_forward(p, xx, tt, CFG["dim"])
        return cap["eps"]

    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        gdo = D.GaussianDiffusionOracle(fwd, image_size=CFG["size"], num_frames=CFG["frames"], channels=CFG["channels"],
                                        timesteps=CFG["timesteps"], loss_type=CFG["loss"])
        loss_ref = float(gdo(x, t, noise).item())
    eps_ref = cap["eps"].double()

    def rel(a):
        return float(((a.double().cpu() - eps_ref).norm() / eps_ref.norm()).item())

    out = {"oracle": "torch fp32 CPU restatement of the reference (oracle/), same weights / clips / t / noise",
           "batch": B, "loss_oracle": loss_ref}
    net.train(False)
    xn = gd.q_sample(x.cuda() * 2 - 1, t.cuda(), noise=noise.cuda())
    eps = net(xn, t.cuda())
    loss = float(gd.p_losses(x.cuda() * 2 - 1, t.cuda(), noise=noise.cuda()).item())
    out["bf16"] = {"loss_rel_err": abs(loss - loss_ref) / loss_ref, "eps_rel_l2": rel(eps), "tolerance": 2e-2}
    if hasattr(net, "forward_fp32"):
        eps32 = net.forward_fp32(xn, t.cuda())
        d = eps32.permute(0, 4, 1, 2, 3) - noise.cuda()
        loss32 = float((d * d).mean().item())
        out["fp32"] = {"loss_rel_err": abs(loss32 - loss_ref) / loss_ref, "eps_rel_l2": rel(eps32), "tolerance": 1e-3}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=CFG["per_gpu_batch"], help="per-GPU batch")
    ap.add_argument("--no-sampling", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-rooflines", action="store_true")
    ap.add_argument("--sample-timesteps", type=int, default=CFG["timesteps"])
    ap.add_argument("--sample-batch", type=int, default=16,
                    help="samples per GPU in the sampling metric (16 = the reference default, gaussian_diffusion.py:323)")
    ap.add_argument("--workload", default="v2_2", choices=["v2_2", "v2_3x"])
    ap.add_argument("--headline", default="train", choices=["train", "sampling"],
                    help="which of BASELINE.json's two metrics is the line's top-level metric / value")
    ap.add_argument("--bucket-mb", type=float, default=None, help="gradient bucket size (MB) of the DP exchange")
    ap.add_argument("--comm-ctas", type=int, default=None, help="cap on the CTAs NCCL may use for the gradient exchange")
    ap.add_argument("--cpu-baseline-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    select_workload(args.workload)
    if args.cpu_baseline_worker:
        return cpu_baseline_worker()
    if args.workload == "v2_3x":  # big model: short sampling run (extrapolated to T), small sample batch
        args.sample_timesteps = min(args.sample_timesteps, 20)
        args.sample_batch = min(args.sample_batch, 4)
        args.no_cpu_baseline = True
        args.no_parity = True
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first: rank 0, in a fresh process with all host cores, before CUDA / NCCL exist in this job (the
    # other ranks wait in the process-group rendezvous, asleep on a socket - not spinning on their GPUs)
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_subprocess()

    import datetime

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=20))
        pg = dist.group.WORLD

    from video_diffusion_nnx_b200 import ops
    from video_diffusion_nnx_b200.gaussian_diffusion import GaussianDiffusion
    from video_diffusion_nnx_b200.trainer import TrainStep
    from video_diffusion_nnx_b200.unet3d import Unet3D

    pk = peaks()
    B = args.batch
    dev = torch.device("cuda", local)
    net = Unet3D(dim=CFG["dim"], channels=CFG["channels"], rngs=0, device=dev)
    gd = GaussianDiffusion(net, image_size=CFG["size"], num_frames=CFG["frames"], channels=CFG["channels"],
                           timesteps=CFG["timesteps"], loss_type=CFG["loss"])
    ts = TrainStep(gd, batch_size=B, train_lr=CFG["lr"], lr_decay_start_step=CFG["lr_decay_start_step"],
                   lr_decay_steps=CFG["lr_decay_steps"], lr_decay_coeff=CFG["lr_decay_coeff"],
                   step_start_ema=0, update_ema_every=CFG["update_ema_every"], ema_decay=CFG["ema_decay"],
                   use_graph=True, process_group=pg,
                   **({"bucket_bytes": int(args.bucket_mb * (1 << 20))} if args.bucket_mb else {}),
                   **({"comm_max_ctas": args.comm_ctas} if args.comm_ctas is not None else {}))
    shape = (B, CFG["channels"], CFG["frames"], CFG["size"], CFG["size"])
    g = torch.Generator().manual_seed(1234 + rank)
    K, Wm = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        tv = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv.item())

    # ---------------- device-resident timing (`value`) ----------------
    x_dev = torch.rand(shape, generator=g).to(dev)
    t_all = torch.randint(0, CFG["timesteps"], (K + Wm + 2, B), generator=g, dtype=torch.int32).to(dev)
    ts.x.copy_(x_dev)
    ts.t.copy_(t_all[0])
    ops.randn(ts.noise, 7 + rank, 0)
    ts.step_device(0)  # eager warm-up pass (state restored) + graph capture + one replayed step
    # kernels of libvdn in one replayed step (counted by the library while the step's graph was captured) + the
    # noise kernel launched per step below
    launches_per_step = ts.launches_per_step + 1
    for i in range(Wm):
        ts.t.copy_(t_all[1 + i])
        ops.randn(ts.noise, 7 + rank, 1 + i)
        ts.step_device(1 + i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            ts.t.copy_(t_all[1 + Wm + i])
            ops.randn(ts.noise, 7 + rank, 1 + Wm + i)
            ts.step_device(1 + Wm + i)
        e1.record()
        barrier()
    loss_last = float(ts.loss.item())
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / K
    value = world * B / (ms_per_step * 1e-3)

    # ---------------- end-to-end through the public API (`e2e`) ----------------
    hosts = [torch.rand(shape, generator=g).pin_memory() for _ in range(4)]
    for i in range(2):
        float(ts.step(hosts[i % 4], 100 + i, 2000 + i).item())
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        loss = ts.step(hosts[i % 4], 1000 + i, 3000 + i)
        _ = float(loss.item())  # device -> host read of the step's result, every step
    barrier()
    e2e_value = world * B * K / max_over_ranks(time.perf_counter() - t0)
    h2d = hosts[0].numel() * 4 + B * 4 + 16 * 4
    d2h = 4

    line = None
    if rank == 0:
        line = {
            "metric": metric_name(args.workload), "value": value, "unit": "clips/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(B, world),
            "loss_last": loss_last,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "how": "TrainStep.step(pinned host batch, key, step) + loss.item() every step, wall clock"},
            "gpu_launches": int(launches_per_step * K),
            "launches_per_step": int(launches_per_step),
            "graph_launches_per_step": ts.graph_launches_per_step,
            "clocks": clk.summary(),
            "train_tflops": value * TRAIN_GFLOP_PER_CLIP / 1e3,
            "train_frac_of_sustained_bf16_peak": value / world * TRAIN_GFLOP_PER_CLIP / 1e3 / pk["tf_sust"],
        }

    # ---------------- sampling metric (frames/s of the full T-step loop) ----------------
    if not args.no_sampling:
        net.train(False)
        sb = args.sample_batch
        T = args.sample_timesteps
        gd.p_sample_loop((sb,), 11, sample_offset=rank * sb, timesteps=4)  # warm-up + capture path
        barrier()
        n0 = ops.lib.vdn_launch_count()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as sclk:
            s0.record()
            vid = gd.p_sample_loop((sb,), 12, sample_offset=rank * sb, timesteps=T)
            s1.record()
            barrier()
        sms = max_over_ranks(s0.elapsed_time(s1))
        # end to end through the public API: the loop above plus the device -> host read of the generated clips
        barrier()
        w0 = time.perf_counter()
        vid_host = gd.p_sample_loop((sb,), 13, sample_offset=rank * sb, timesteps=T).cpu()
        s_e2e = max_over_ranks(time.perf_counter() - w0)
        if rank == 0:
            scale = CFG["timesteps"] / T
            fps_full = world * sb * CFG["frames"] / (sms * 1e-3 * scale)
            eng = gd._samplers[sb]["eng"]
            n_eager = int(ops.lib.vdn_launch_count() - n0)  # init noise etc.; the timestep graph replays are on top
            line["sampling"] = {"metric": f"sampling frames/sec (p_sample_loop, T=1000, config_{args.workload})", "value": fps_full,
                                "unit": "frames/s", "timesteps_run": T, "ms_per_timestep": sms / T,
                                "sample_batch_per_gpu": sb, "finite": bool(torch.isfinite(vid).all().item()),
                                "e2e": {"value": world * sb * CFG["frames"] / (s_e2e * scale), "unit": "frames/s",
                                        "h2d_bytes_per_step": 24, "d2h_bytes_per_step": vid_host.numel() * 4},
                                "sampling_tflops": fps_full * FWD_GFLOP_PER_CLIP / CFG["frames"] * CFG["timesteps"] / 1e3,
                                "frac_of_sustained_bf16_peak": fps_full / world * FWD_GFLOP_PER_CLIP / CFG["frames"]
                                * CFG["timesteps"] / 1e3 / pk["tf_sust"], "clocks": sclk.summary(),
                                "eager_launches": n_eager}
            del eng

    if world > 1:  # the rest is rank 0's alone: leave the process group first so that nobody spins in a barrier
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    if args.headline == "sampling" and "sampling" in line:
        smp = line["sampling"]
        train = {k: line[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "launches_per_step",
                                      "clocks", "loss_last")}
        line.update(metric=smp["metric"], value=smp["value"], unit=smp["unit"], ms_per_step=smp["ms_per_timestep"],
                    e2e=smp["e2e"], clocks=smp["clocks"], steps=smp["timesteps_run"], training=train)
        line["config"]["workload"] = (f"{WORKLOAD_NAME} p_sample_loop, T=1000 ({smp['timesteps_run']} timesteps run): Unet3D dim "
                                      f"{CFG['dim']}, {CFG['frames']} frames, {CFG['size']}x{CFG['size']}; {smp['sample_batch_per_gpu']} samples per GPU")
    if not args.no_rooflines:
        small = args.workload == "v2_2"
        # the dominant kernel of the step by the committed launch list (profiles/): the small-M tap-GEMM convs of the
        # 16x16 / 8x8 levels (tapgemm_kernel<64>); the full-resolution conv and three more kernels follow
        if small:
            # dominant kernel of the step by name: wgrad_kernel<0>; then the conv of the small-M levels (the generic
            # tap-GEMM, second by share and the kernel furthest below its roofline), the 16x16 and 64x64 convs, ...
            line["roofline"] = wgrad_roofline(torch, ops, pk)
            extra = [conv_roofline(torch, ops, pk, level=3, with_traffic=True), conv_roofline(torch, ops, pk, level=2),
                     conv_roofline(torch, ops, pk, level=0, with_traffic=True)]
        else:
            line["roofline"] = conv_roofline(torch, ops, pk, level=0, with_traffic=True)
            extra = []
        line["roofline_kernels"] = extra + extra_rooflines(torch, ops, pk)
    if not args.no_parity:
        line["parity"] = parity_block(torch, ops, gd, net, B)
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
