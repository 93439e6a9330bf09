#!/usr/bin/env python
"""bench.py — headline benchmark of the ConsistencySolver sampling hot path on B200.

Workload (BASELINE.json configs[1], solver path): SD1.5-shape latents 4x64x64 fp32, 8-step trailing schedule,
CFG=3, order_dim=4, scaler_dim=0, production policy (2->256->256->33), batch 64 per GPU.  The denoiser is NOT
the product: its outputs are resident synthetic CFG pairs (the "random eps stand-in for the U-Net" of configs[0]);
`with_denoiser` reports the same loop with a random-init SD1.5-architecture U-Net in the middle when requested.

A "step" = one 8-step preview of one batch of 64 latents: 8 x (Exp(1) draw + policy kernel + fused step kernel).
  value      previews/s, inputs resident in HBM, rotating pool of batches larger than L2, CUDA-event timed,
             max over ranks (weak scaling: every rank runs its own shard of prompts/seeds, no collective)
  e2e        previews/s through the public scheduler API with HOST buffers: every step uploads the initial
             latents and the 8 CFG pairs from pinned memory and reads the final latents + rollout record back
  roofline   the fused step kernel (dominant kernel) at this workload's launch shape, timed live with CUDA events
             on its launch stream; `roofline_sweep` repeats it for larger batches (BASELINE config 2)
  cpu_baseline / --impl reference   the UNMODIFIED reference PPOScheduler (oracle/_ref, staged by oracle/stage_ref.py)
             on the host cores; the torch-CPU oracle port as a second row (and as the fallback when oracle/_ref is absent)
  torch_eager_gpu / torch_compile_gpu   the same reference classes on cuda:0, stock torch (SURVEY §2.2's bar)
  ppo_rollout  BASELINE configs[4]: batched rollouts + PPO update with the flat-buffer gradient all-reduce

Timed regions.  The driver calls `--steps 20 --warmup 5`; 20 previews are 0.8 ms of GPU time, shorter than the
fill/drain of the four-deep preview pipeline.  Every timed leg therefore repeats its `steps`-block R times back to
back with the pipeline kept full (R from a calibration pass, so that the region lasts >= TARGET_REGION_S) and reports
ms_per_step = elapsed / (R * steps), `reps`, `timed_region_s` and the measured fill/drain cost of one isolated block.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SD_CFG = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000,
              steps_offset=1, timestep_spacing="trailing", order_dim=4, scaler_dim=0, use_conv=False)
FN_KW = dict(embedding_dim=64, hidden_dim=256, num_actions=11)
SHAPE = (4, 64, 64)
N_STEPS = 8
GUIDANCE = 3.0
L2_BYTES = 126 * 2 ** 20
TARGET_REGION_S = 0.4                                     # minimum length of a timed region (see module docstring)
WORKLOAD = ("BASELINE configs[1] solver path: SD1.5 latents 4x64x64 fp32, 8-step trailing, CFG=3, order_dim=4, "
            "scaler_dim=0, policy 2-256-256-33, batch 64/GPU, denoiser = resident synthetic CFG pairs")
HIST_DEPTHS = [1, 2, 3, 4, 4, 4, 4, 4]                     # n_hist per step of an 8-step preview
TENSORS_PER_PREVIEW = sum(n + 4 for n in HIST_DEPTHS)       # 58 latent-sized transfers / sample


def workload_config(B, world):
    """`config` of the JSON line — the SAME dict in both arms (ours and --impl reference), so the driver's
    same_config check compares like with like; arm-specific launch details go to `run`."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "solver_steps": N_STEPS, "guidance": GUIDANCE,
            "sharding": f"dp{world} by prompt/seed, no collective",
            "l2_policy": "GPU arm: inputs larger than L2 (rotating pool of resident batches, 17x the 126 MB L2 by default); "
                         "CPU arm: one resident batch (host caches are not the bound there)"}


def policy_state_dict(seed=0):
    """Random-init stand-in for the published checkpoint: default nn.Linear init for layers 0/2,
    N(0, 0.05^2) last-layer weight, zero bias, all under one seed."""
    g = torch.Generator().manual_seed(seed)
    H, AK = FN_KW["hidden_dim"], 3 * FN_KW["num_actions"]
    u = lambda *s, fan: (torch.rand(*s, generator=g) * 2 - 1) / fan ** 0.5  # noqa: E731
    from consolver_b200.factor_net import _action_values

    return {"action_values": _action_values("sd", FN_KW["num_actions"], 4, 0, 0),
            "mlp.0.weight": u(H, 2, fan=2), "mlp.0.bias": u(H, fan=2),
            "mlp.2.weight": u(H, H, fan=H), "mlp.2.bias": u(H, fan=H),
            "mlp.4.weight": torch.randn(AK, H, generator=g) * 0.05, "mlp.4.bias": torch.zeros(AK)}


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) — runs during the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self._stop, self.ok = [], set(), threading.Event(), False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.max = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()
            if not self.samples:      # a region shorter than the thread's start-up: sample now, the GPU is still busy/hot
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                except Exception:  # noqa: BLE001
                    pass

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def make_scheduler(device, sd):
    import consolver_b200 as cb

    s = cb.PPOScheduler(factor_net_kwargs=dict(FN_KW), **SD_CFG)
    s.factor_net.load_state_dict(sd)
    s.factor_net.to(device)
    return s


def synth_batch(B, seed, device, pin=False):
    """Synthetic inputs of one preview batch: x_T ~ N(0,1) and one N(0,1) CFG pair per step."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, *SHAPE, generator=g)
    pairs = torch.randn(N_STEPS, 2 * B, *SHAPE, generator=g)
    if device is None:
        return (x.pin_memory(), pairs.pin_memory()) if pin else (x, pairs)
    return x.to(device), pairs.to(device)


def time_step_kernel(B, n_hist, device, iters=64, pool_bytes=3 * L2_BYTES, flags=0):
    """Average device time of ONE fused-step launch (CFG pair, n_hist deep) over `iters` launches on rotating
    buffer sets larger than L2, captured in a CUDA graph so host launch gaps do not enter; CUDA events on the
    launch stream.  `flags` = CONSOLVER_FLAG_CHAIN times the launch form the product's replay graphs use
    (each step a programmatic dependent launch of the previous one, so one launch's tail overlaps the next one's
    ramp); the launches here are independent, which is what the flag requires of everything but x / hist[0]."""
    from consolver_b200 import _lib

    lib = _lib.load()
    N = SHAPE[0] * SHAPE[1] * SHAPE[2]
    per_launch = (n_hist + 4) * B * N * 4
    nsets = int(max(2, min(64, -(-pool_bytes // per_launch))))
    mk = lambda: torch.randn(B, N, device=device)  # noqa: E731
    sets = [dict(u=mk(), c=mk(), x=mk(), h=[mk() for _ in range(n_hist - 1)], o=torch.empty(B, N, device=device),
                 s=torch.empty(B, N, device=device)) for _ in range(nsets)]
    coef = torch.randn(B, 6, device=device)

    def launch(st, stream):
        rc = lib.consolver_step_sd(0, st["u"].data_ptr(), st["c"].data_ptr(), GUIDANCE, st["s"].data_ptr(),
                                   _lib.ptr_array([t.data_ptr() for t in st["h"]]), n_hist, st["x"].data_ptr(),
                                   st["o"].data_ptr(), None, 0, coef.data_ptr(), 6, 4, 0.8378, 0.5460, 0.9151, 0.4033,
                                   flags, B, N, stream)
        assert rc == 0, rc

    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        for i in range(max(3, nsets)):
            launch(sets[i % nsets], side.cuda_stream)
    torch.cuda.current_stream(device).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        st = torch.cuda.current_stream(device).cuda_stream
        for i in range(iters):
            launch(sets[i % nsets], st)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(device)
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize(device)
        ts.append(a.elapsed_time(b) * 1e3 / iters)
    ts.sort()
    return ts[len(ts) // 2], per_launch, nsets


def time_fm_kernel(B, device, n_hist=2, iters=32, pool_bytes=3 * L2_BYTES):
    """FM (FLUX-Kontext shape, BASELINE configs[3]) fused step: packed latents [B,4096,64] bf16, order_dim=2 steady
    state — reads v, x and one older slot, writes x' = 4 tensors of 512 KiB per sample."""
    from consolver_b200 import _lib

    lib = _lib.load()
    N = 4096 * 64
    per_launch = (n_hist + 2) * B * N * 2
    nsets = int(max(2, min(64, -(-pool_bytes // per_launch))))
    mk = lambda: torch.randn(B, N, device=device).bfloat16()  # noqa: E731
    sets = [dict(v=mk(), x=mk(), h=[mk() for _ in range(n_hist - 1)],
                 o=torch.empty(B, N, device=device, dtype=torch.bfloat16)) for _ in range(nsets)]
    coef = torch.randn(B, 4, device=device)

    def launch(st, stream):
        rc = lib.consolver_step_fm(2, 2, st["v"].data_ptr(), None, _lib.ptr_array([t.data_ptr() for t in st["h"]]),
                                   n_hist, st["x"].data_ptr(), st["o"].data_ptr(), None, 0, coef.data_ptr(), 4, 2,
                                   -0.0433, 0,
                                   B, N, stream)
        assert rc == 0, rc

    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        for i in range(max(3, nsets)):
            launch(sets[i % nsets], side.cuda_stream)
    torch.cuda.current_stream(device).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        st = torch.cuda.current_stream(device).cuda_stream
        for i in range(iters):
            launch(sets[i % nsets], st)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(device)
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize(device)
        ts.append(a.elapsed_time(b) * 1e3 / iters)
    ts.sort()
    return ts[len(ts) // 2], per_launch


def with_denoiser(B, device, sd, previews=3):
    """The same 8-step CFG loop with a random-init SD1.5-architecture U-Net (bf16, channels_last, SDPA) producing the
    model outputs — timed, not the product.  Reports previews/s and the solver's share of the loop."""
    from consolver_b200.denoise import denoise_loop
    from consolver_b200.standins import SD15UNet

    torch.manual_seed(0)
    unet = SD15UNet().to(device=device, dtype=torch.bfloat16).to(memory_format=torch.channels_last).eval()
    ctx = torch.randn(2 * B, 77, 768, device=device, dtype=torch.bfloat16)
    sched = make_scheduler(device, sd)
    solver_events = []
    orig = sched.step_cfg

    def timed_step(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **k)
        e1.record()
        solver_events.append((e0, e1))
        return r

    sched.step_cfg = timed_step

    @torch.no_grad()
    def den(x, t, i):
        return unet(x.to(dtype=torch.bfloat16, memory_format=torch.channels_last), t, ctx).float()

    noise = torch.randn(B, *SHAPE, device=device)
    denoise_loop(sched, den, noise, cfg=GUIDANCE, num_inference_steps=N_STEPS)          # warm-up (cuDNN autotune)
    torch.cuda.synchronize(device)
    solver_events.clear()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(previews):
        denoise_loop(sched, den, noise, cfg=GUIDANCE, num_inference_steps=N_STEPS)
    b.record()
    torch.cuda.synchronize(device)
    ms = a.elapsed_time(b) / previews
    solver_ms = sum(x.elapsed_time(y) for x, y in solver_events) / previews
    # ---- interactive preview latency (batch 1): eager loop vs the whole loop, denoiser included, in one CUDA graph ----
    latency = None
    try:
        from consolver_b200.denoise import GraphedDenoiseLoop

        sched.step_cfg = orig
        ctx1 = ctx[:2].contiguous()
        t_dev = {}

        @torch.no_grad()
        def den1(x, t, i):
            if i not in t_dev:
                t_dev[i] = torch.tensor([int(t)], device=device)
            return unet(x.to(dtype=torch.bfloat16, memory_format=torch.channels_last), t_dev[i], ctx1).float()

        n1 = torch.randn(1, *SHAPE, device=device)
        for _ in range(2):
            denoise_loop(sched, den1, n1, cfg=GUIDANCE, num_inference_steps=N_STEPS)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(3):
            denoise_loop(sched, den1, n1, cfg=GUIDANCE, num_inference_steps=N_STEPS)
        torch.cuda.synchronize(device)
        eager_ms = (time.perf_counter() - t0) / 3 * 1e3
        gl = GraphedDenoiseLoop(sched, den1, n1, GUIDANCE, N_STEPS)
        gl.replay()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(5):
            gl.replay()
        torch.cuda.synchronize(device)
        graph_ms = (time.perf_counter() - t0) / 5 * 1e3
        latency = {"batch": 1, "eager_ms_per_preview": round(eager_ms, 2), "graph_ms_per_preview": round(graph_ms, 2),
                   "note": "8 steps, CFG (2 rows per U-Net call); graph = GraphedDenoiseLoop: U-Net forwards + solver "
                           "kernels in one CUDA graph"}
        del gl
    except Exception as e:  # noqa: BLE001
        latency = {"error": repr(e)[:200]}
    del unet
    torch.cuda.empty_cache()
    return {"value": round(B / (ms / 1e3), 2), "unit": "previews/s", "denoiser": "random-init SD1.5-architecture U-Net "
            "(859.5 M params), bf16 channels_last, 2B rows per call (CFG), stock PyTorch — timed, not the product",
            "ms_per_preview_batch": round(ms, 2), "solver_ms_per_preview_batch": round(solver_ms, 4),
            "solver_share": round(solver_ms / ms, 6), "previews_timed": previews, "batch_per_gpu": B,
            "preview_latency": latency}


def fm_preview_throughput(device, B=16, steps=2000):
    """BASELINE configs[3]: FLUX-Kontext-shaped FMPPOScheduler loop — packed latents [B,4096,64] bf16, 8 steps,
    order_dim=2, resident synthetic velocities as the transformer stand-in, one CUDA graph per preview."""
    import numpy as np

    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview

    torch.manual_seed(0)
    nbytes = (1 + N_STEPS) * B * 4096 * 64 * 2
    pool_n = max(16, int(-(-2 * L2_BYTES // nbytes)))
    pool = []
    for j in range(pool_n):
        s = cb.FMPPOScheduler(shift=3.0, use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15, base_image_seq_len=256,
                              max_image_seq_len=4096, order_dim=2, scaler_dim=0, mu_dim=0,
                              factor_net_kwargs=dict(hidden_dim=256, num_actions=11))
        s.factor_net.to(device)
        x = torch.randn(B, 4096, 64, device=device).bfloat16()
        vs = [torch.randn(B, 4096, 64, device=device).bfloat16() for _ in range(N_STEPS)]
        pool.append(GraphedPreview(s, x, vs, None, N_STEPS, set_timesteps_kwargs=dict(
            sigmas=np.linspace(1.0, 1 / N_STEPS, N_STEPS), mu=1.15)))
    from consolver_b200.denoise import PreviewGroup, PreviewPool

    # same launch form as the headline leg: chains of 2 previews per CUDA graph, replayed round-robin on 4 streams
    g = 2
    groups = [PreviewGroup(pool[i:i + g], rotation=pool_n // g, parallel=False) for i in range(0, pool_n - pool_n % g, g)]
    pp = PreviewPool(groups, streams=4)
    n_sub = max(1, steps // g)
    for k in range(2 * len(groups)):
        pp.submit(k % len(groups))
    pp.join()
    torch.cuda.synchronize(device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(n_sub):
        pp.submit(k % len(groups))
    pp.join()
    b.record()
    torch.cuda.synchronize(device)
    ms = a.elapsed_time(b) / (n_sub * g)
    tensors = 3 + 7 * 4     # n = 1, then 2: (n+2) tensors per step
    return {"value": round(B / (ms / 1e3), 1), "unit": "previews/s", "batch": B, "ms_per_preview_batch": round(ms, 4),
            "shape": "[B,4096,64] bf16, 8 steps, order_dim=2", "algorithmic_gbs": round(
                tensors * B * 4096 * 64 * 2 / (ms / 1e3) / 1e9, 1), "pool_batches": pool_n,
            "previews_timed": n_sub * g,
            "concurrency": f"chains of {g} previews per CUDA graph, {len(groups)} groups round-robin on {len(pp.streams)} streams"}


def run_ours(args, rank, world, device):
    import consolver_b200  # noqa: F401  (fails loudly if libconsolver.so cannot be built/loaded)
    from consolver_b200.denoise import GraphedPreview, preview_from_pairs

    B = args.batch
    sd = policy_state_dict(0)
    bytes_per_batch = (1 + 2 * N_STEPS) * B * SHAPE[0] * SHAPE[1] * SHAPE[2] * 4
    pool_n = int(max(2, -(-2 * L2_BYTES // bytes_per_batch)))
    if not args.eager:
        pool_n = max(pool_n, 2 * max(1, args.streams))     # two resident batches per stream
        if not args.no_groups and args.streams > 1:
            # preview groups (see below) of `group_size` previews each, replayed round-robin over a pool of at least
            # `pool` resident batches
            gsz = args.group_size or args.streams
            pool_n = max(pool_n, args.group_rotation * gsz, args.pool)
            pool_n -= pool_n % gsz
    pool = []
    for j in range(pool_n):
        s = make_scheduler(device, sd)
        x, pairs = synth_batch(B, 1234 + rank * 1000 + j, device)          # shard = distinct seeds per rank
        gp = GraphedPreview(s, x, list(pairs.unbind(0)), GUIDANCE, N_STEPS) if not args.eager else None
        pool.append((s, x, pairs, gp))
    torch.manual_seed(1000 + rank)

    def one_step(k):
        s, x, pairs, gp = pool[k % pool_n]
        if gp is not None:
            return gp.replay()
        s.set_timesteps(N_STEPS, device=device)
        return preview_from_pairs(s, x, pairs.unbind(0), GUIDANCE)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    # Independent preview batches (different prompts/seeds) have no dependency on each other: keep `streams` of them
    # in flight on separate CUDA streams so one batch's launch ramp / tail overlaps another's streaming phase.
    # Pool entry j always runs on stream j % streams, so no two streams touch the same buffers.
    from consolver_b200.denoise import PreviewPool

    ppool = None if args.eager else PreviewPool([p[3] for p in pool], streams=args.streams)
    n_streams = 1 if ppool is None else len(ppool.streams)
    # Groups of `group_size` previews captured as ONE graph each (denoise.PreviewGroup: a serial chain of previews, one
    # shared device-resident generator state): one host launch and one generator update per g previews instead of per
    # preview.  Without it the host needs ~30 us per preview, as much as the GPU does, and eight ranks sharing one host's
    # cores become host-bound.  Previews are still taken round-robin from the pool; a stretch that is not aligned to a
    # group goes through the per-preview graphs.
    groups, g = [], 0
    if ppool is not None and n_streams > 1 and not args.no_groups:
        from consolver_b200.denoise import PreviewGroup

        g = args.group_size or n_streams
        n_groups = (pool_n - pool_n % g) // g
        # each group = g previews chained on ONE branch (parallel=False); the parallelism comes from replaying the
        # n_groups groups on n_groups streams, every stream a gap-free chain of previews
        groups = [PreviewGroup([pool[j][3] for j in range(i, i + g)], rotation=n_groups, parallel=not args.serial_groups)
                  for i in range(0, n_groups * g, g)]
    # the groups are replayed round-robin on `group_rotation` streams: every stream is a gap-free chain of previews
    gpool = PreviewPool(groups, streams=min(len(groups), args.group_rotation), stagger_us=args.stagger_us) \
        if groups else None
    counts = {"group_replays": 0, "single_replays": 0}

    def run_steps(first, count, join=True):
        if ppool is None:
            for k in range(first, first + count):
                one_step(k)
            return
        k, end, pending, gpending = first, first + count, False, False
        while k < end:
            slot = k % pool_n
            if groups and slot % g == 0 and slot + g <= len(groups) * g and k + g <= end:
                if pending:                       # per-preview replays on the side streams touch the same buffers
                    ppool.join()
                    pending = False
                gpool.submit(slot // g)
                counts["group_replays"] += 1
                gpending = True
                k += g
            else:
                if gpending:
                    gpool.join()
                    gpending = False
                ppool.submit(slot)
                counts["single_replays"] += 1
                pending = True
                k += 1
        if join or pending:
            ppool.join()
        if gpool is not None and (join or (gpending and pending)):
            gpool.join()

    def join_all():
        if ppool is not None:
            ppool.join()
        if gpool is not None:
            gpool.join()

    def rank_max(v):
        if world > 1:
            t = torch.tensor([v], device=device, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return t.item()
        return v

    run_steps(0, args.warmup)
    # ---- calibration: ONE isolated block of `steps` previews from an idle GPU (pays the pipeline fill and drain) ------
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run_steps(args.warmup, args.steps)
    b.record()
    barrier()
    block_ms = rank_max(a.elapsed_time(b))

    def timed_blocks(n_blocks, sampler=None):
        """n_blocks x `steps` previews back to back, pipeline kept full, one join at the end -> ms (max over ranks)"""
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with (sampler or contextlib.nullcontext()):
            ea.record()
            h0 = time.perf_counter()
            k0 = 0                      # aligned to the preview groups (a block of `steps` = steps/g group replays)
            for _ in range(n_blocks):
                run_steps(k0, args.steps, join=False)
                k0 += args.steps
            host_us[0] = (time.perf_counter() - h0) * 1e6 / (n_blocks * args.steps)   # host time to ENQUEUE one preview
            join_all()
            eb.record()
            barrier()
        return rank_max(ea.elapsed_time(eb))

    host_us = [0.0]

    # second calibration stage (untimed for the result): ~0.15 s of back-to-back blocks gives the steady-state rate the
    # repetition count is derived from, and doubles as a warm-up that does not depend on how small --warmup is
    pre_blocks = int(max(1, min(-(-150.0 // max(block_ms, 1e-3)), 1_000_000 // max(args.steps, 1))))
    pre_ms = timed_blocks(pre_blocks)
    est_ms_per_block = max(pre_ms / pre_blocks, 1e-4)
    reps = int(max(1, min(-(-TARGET_REGION_S * 1e3 // est_ms_per_block), 4_000_000 // max(args.steps, 1))))
    if args.reps:
        reps = args.reps
    # ---- the timed region: `reps` blocks of `steps` previews back to back, pipeline kept full, one join at the end ----
    clk = ClockSampler(torch.cuda.current_device())
    counts.update(group_replays=0, single_replays=0)
    ms = timed_blocks(reps, clk)
    timed_counts = dict(counts)
    n_timed = reps * args.steps
    value = world * n_timed * B / (ms / 1e3)
    ms_per_step = ms / n_timed
    fill_drain_ms = max(0.0, block_ms - args.steps * ms_per_step)

    # for transparency: the same loop with ONE preview batch in flight (no cross-batch overlap), rank-local
    single = None
    if ppool is not None and n_streams > 1:
        k1 = min(args.steps, 500)
        for k in range(8):
            one_step(k)
        torch.cuda.synchronize(device)
        a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a1.record()
        for k in range(k1):
            one_step(k)
        b1.record()
        torch.cuda.synchronize(device)
        single = {"value_per_gpu": round(k1 * B / (a1.elapsed_time(b1) / 1e3), 1),
                  "us_per_preview_batch": round(a1.elapsed_time(b1) * 1e3 / k1, 2)}

    if args.no_extras:
        return {"metric": "sd15_8step_solver_previews_per_s", "value": round(value, 1), "unit": "previews/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5),
                "reps": reps, "timed_region_s": round(ms / 1e3, 4)}
    # ---- e2e: host buffers -> scheduler API -> host, copies inside the timed region ---------------------------
    # Every step uploads ITS initial latents and ITS 8 CFG pairs from pinned host memory (75.5 MB), runs the 8-step
    # preview through the public API (GraphedPreview over that buffer set) and reads the final latents and the
    # rollout record back to pinned memory.  Two buffer sets: the upload of step k+1 overlaps the compute and the
    # read-back of step k (PCIe is full duplex), so the steady state is bound by the H2D link.
    nset = 2
    hx, hpairs = synth_batch(B, 99 + rank, None, pin=True)
    e2e_sets = []
    for j in range(nset):
        sch = make_scheduler(device, sd)
        dx = torch.empty_like(hx, device=device)
        dp = torch.empty_like(hpairs, device=device)
        dx.copy_(hx)
        dp.copy_(hpairs)
        gp = GraphedPreview(sch, dx, list(dp.unbind(0)), GUIDANCE, N_STEPS) if not args.eager else None
        e2e_sets.append(dict(s=sch, dx=dx, dp=dp, gp=gp, hout=torch.empty_like(hx).pin_memory(),
                             hrec=torch.empty(N_STEPS, B, 3).pin_memory(), up=torch.cuda.Event(), done=torch.cuda.Event(),
                             down=torch.cuda.Event()))
    up_stream, down_stream = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)

    def e2e_step(k):
        st = e2e_sets[k % nset]
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(st["done"])                       # buffer set free again (compute of step k-2)
            st["dx"].copy_(hx, non_blocking=True)
            st["dp"].copy_(hpairs, non_blocking=True)
            st["up"].record(up_stream)
        main.wait_event(st["up"])
        main.wait_event(st["down"])                                # previous read-back of this set finished
        if st["gp"] is not None:
            x = st["gp"].replay()
        else:
            st["s"].set_timesteps(N_STEPS, device=device)
            x = preview_from_pairs(st["s"], st["dx"], st["dp"].unbind(0), GUIDANCE)
        st["done"].record(main)
        with torch.cuda.stream(down_stream):
            down_stream.wait_event(st["done"])
            st["hout"].copy_(x, non_blocking=True)
            st["hrec"].copy_(st["s"]._traj.out["probs"], non_blocking=True)
            st["down"].record(down_stream)

    def e2e_drain():
        up_stream.synchronize()
        main.synchronize()
        down_stream.synchronize()

    e2e_steps = args.steps
    for k in range(4):
        e2e_step(k)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()                                        # calibration block
    for k in range(e2e_steps):
        e2e_step(k)
    e2e_drain()
    barrier()
    e2e_block_s = rank_max(time.perf_counter() - t0)
    e2e_reps = int(max(1, min(-(-TARGET_REGION_S // max(e2e_block_s, 1e-6)) + 1, 200_000 // max(e2e_steps, 1))))
    barrier()
    with ClockSampler(torch.cuda.current_device()) as clk_e2e:     # the second timed region: sampled as well
        t0 = time.perf_counter()
        for k in range(e2e_reps * e2e_steps):
            e2e_step(k)
        e2e_drain()
        barrier()
        e2e_s = time.perf_counter() - t0
    e2e_s = rank_max(e2e_s)
    e2e_val = world * e2e_reps * e2e_steps * B / e2e_s
    h2d = (hx.numel() + hpairs.numel()) * 4
    d2h = (e2e_sets[0]["hout"].numel() + e2e_sets[0]["hrec"].numel()) * 4

    out = {
        "metric": "sd15_8step_solver_previews_per_s", "value": round(value, 1), "unit": "previews/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5),
        "reps": reps, "timed_region_s": round(ms / 1e3, 4), "steps_timed": n_timed,
        "isolated_block": {"ms": round(block_ms, 4), "fill_drain_ms": round(fill_drain_ms, 4),
                           "note": f"one block of {args.steps} previews timed alone from an idle GPU (what a "
                                   f"{args.steps}-step region would have measured): "
                                   f"{round(world * args.steps * B / (block_ms / 1e3), 1)} previews/s"},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, world),
        "run": {"pool": f"rotating pool of {pool_n} resident batches ({pool_n * bytes_per_batch >> 20} MiB)",
                "launch": "eager python launches" if args.eager else
                "CUDA graphs of whole 8-step previews (per preview: table kernel, sample kernels as a parallel branch, "
                "PDL-chained step kernels; one rng-advance node per graph); valid because the stand-in model outputs are "
                "resident",
                "concurrency": f"{n_streams} independent preview batch(es) in flight" + (
                    f": groups of {g} previews captured as one CUDA graph ("
                    + ("a serial chain per group" if args.serial_groups else f"{g} parallel branches") +
                    f"), {len(groups)} groups replayed round-robin on {len(gpool.streams)} streams "
                    f"({timed_counts['group_replays']} group + {timed_counts['single_replays']} single replays timed)"
                    if groups else " on separate CUDA streams"),
                "timing": f"{reps} x {args.steps} previews back to back between two CUDA events, max over ranks",
                "host_enqueue_us_per_step": round(host_us[0], 2),
                "calibration": {"isolated_block_ms": round(block_ms, 4), "pre_blocks": pre_blocks,
                                "pre_ms_per_step": round(pre_ms / (pre_blocks * args.steps), 5)}},
        "e2e": {"value": round(e2e_val, 1), "unit": "previews/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "reps": e2e_reps, "timed_region_s": round(e2e_s, 4)},
        # per preview: table + 8 x (sample + step) kernels; one rng-advance node per graph replay (group or single)
        "gpu_launches": n_timed * (1 + N_STEPS * 2) + timed_counts["group_replays"] + timed_counts["single_replays"],
        "clocks": clk.summary(),
        "clocks_e2e": clk_e2e.summary(),       # the host-link-bound leg, where the SMs idle most of the time
    }
    # whole-loop HBM rate: 58 latent-sized transfers per sample per 8-step preview, per GPU
    loop_gbs = value / world * TENSORS_PER_PREVIEW * SHAPE[0] * SHAPE[1] * SHAPE[2] * 4 / 1e9
    out["solver_loop"] = {"algorithmic_bytes_per_preview": TENSORS_PER_PREVIEW * SHAPE[0] * SHAPE[1] * SHAPE[2] * 4,
                          "achieved_gbs_per_gpu": round(loop_gbs, 1), "one_batch_in_flight": single}
    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak, src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)") if "hbm_gbs" in peaks \
            else (6650.0, "fallback (B200_PROFILING.md)")
        us, nbytes, nsets = time_step_kernel(B, 4, device)
        # DRAM bytes per launch: NOT measured in this run (a profiler cannot run inside a timed bench) — read from the
        # committed `ncu --set full` capture of the same kernel at the same launch shape, and labelled as such
        traffic_db, traffic_src = {}, None
        for cand in ("ncu_traffic_r02.json", "ncu_traffic_r01.json"):
            tpath = os.path.join(ROOT, "profiles", cand)
            if os.path.exists(tpath):
                traffic_db, traffic_src = json.load(open(tpath))["step_kernel_f32_nh4_pair"], f"profiles/{cand}"
                break

        def traffic_of(batch):
            rec = traffic_db.get(str(batch))
            return None if not rec else {"dram_read": rec["dram_read"], "dram_write": rec["dram_write"],
                                         "total": rec["dram_read"] + rec["dram_write"]}

        tr0 = traffic_of(B)
        out["roofline"] = {"bound": "hbm", "kernel": "step_kernel<f32,NH=4,CFG pair>", "batch": B,
                           "achieved": round(nbytes / us / 1e3, 1), "peak": peak, "unit": "GB/s",
                           "frac": round(nbytes / us / 1e3 / peak, 4), "traffic": tr0["total"] if tr0 else None,
                           "traffic_detail": tr0, "traffic_source": traffic_src and
                           f"{traffic_src} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch; at "
                           f"this batch the 2 written tensors can still sit dirty in the 126 MB L2 when the kernel ends, "
                           f"so writes may be under-counted)",
                           "us_per_launch": round(us, 3), "algorithmic_bytes": nbytes, "peak_source": src}
        from consolver_b200._lib import FLAG_CHAIN
        us_c, _, _ = time_step_kernel(B, 4, device, flags=FLAG_CHAIN)
        out["roofline"]["chained"] = {
            "us_per_launch": round(us_c, 3), "achieved": round(nbytes / us_c / 1e3, 1),
            "frac": round(nbytes / us_c / 1e3 / peak, 4),
            "note": "same launches as programmatic dependent launches of one another (CONSOLVER_FLAG_CHAIN), the form "
                    "the replay graphs of the `value` leg use; `frac` above is the plain, fully serialised launch"}
        sweep = []
        for Bs in (256, 1024, 4096):
            us, nbytes, nsets = time_step_kernel(Bs, 4, device, iters=32)
            trb = traffic_of(Bs)
            sweep.append({"batch": Bs, "us_per_launch": round(us, 3), "achieved": round(nbytes / us / 1e3, 1),
                          "frac": round(nbytes / us / 1e3 / peak, 4), "algorithmic_bytes": nbytes,
                          "traffic": trb["total"] if trb else None, "traffic_detail": trb})
        out["roofline_sweep"] = sweep
        out["roofline_sweep_traffic_source"] = traffic_src
        out["solver_loop"]["frac_of_peak"] = round(out["solver_loop"]["achieved_gbs_per_gpu"] / peak, 4)
        # the same kernel INSIDE the timed region of `value`: the region's device time divided by the step launches it
        # contains (8 per preview; the policy kernels run on a parallel graph branch) against the mean algorithmic bytes
        # of those launches (58 tensors per 8 launches: history depths 1,2,3,4,4,4,4,4)
        per_launch_bytes = TENSORS_PER_PREVIEW * B * SHAPE[0] * SHAPE[1] * SHAPE[2] * 4 / N_STEPS
        us_in = out["ms_per_step"] * 1e3 / N_STEPS
        out["roofline"]["in_timed_region"] = {
            "us_per_launch": round(us_in, 3), "algorithmic_bytes": int(per_launch_bytes),
            "achieved": round(per_launch_bytes / us_in / 1e3, 1), "frac": round(per_launch_bytes / us_in / 1e3 / peak, 4),
            "note": "timed region of `value` / step launches in it (several previews in flight, so one launch's ramp and "
                    "tail overlap other launches' streaming phases); `frac` above is ONE launch alone on an idle GPU"}
        fm = []
        for Bs in (1, 8, 64, 512):
            us, nbytes = time_fm_kernel(Bs, device)
            fm.append({"batch": Bs, "us_per_launch": round(us, 3), "achieved": round(nbytes / us / 1e3, 1),
                       "frac": round(nbytes / us / 1e3 / peak, 4)})
        out["roofline_fm_flux_bf16"] = {"kernel": "step_kernel<bf16,NH=2,FM>", "shape": "[B,4096,64] bf16 (FLUX-Kontext "
                                        "packed 1024^2 latents), order_dim=2", "bytes_per_sample": 4 * 4096 * 64 * 2,
                                        "points": fm}
        try:
            out["fm_flux_preview"] = fm_preview_throughput(device)
        except Exception as e:  # noqa: BLE001
            out["fm_flux_preview"] = {"error": repr(e)[:200]}
        # the reference's own classes as stock torch on this GPU (SURVEY §2.2's bar), each in a fresh process with a
        # time limit: dynamo state and a possible compile hang stay out of this process
        torch.cuda.synchronize(device)
        if not args.no_torch_ref:
            for mode in ("eager", "compile"):
                out[f"torch_{mode}_gpu"] = torch_ref_gpu_subprocess(mode, B, local_index=torch.cuda.current_device())
        # the CPU arm last: every GPU is idle now and the other ranks SLEEP in a store wait (no NCCL spin on the host)
        out["cpu_baseline"] = cpu_baseline(B, budget_s=args.cpu_budget)
    rank0_section_done(world, rank, "rank0_legs")
    if not args.no_ppo:
        try:
            out_ppo = ppo_rollout_leg(rank, world, device)
            if rank == 0:
                out["ppo_rollout"] = out_ppo
        except Exception as e:  # noqa: BLE001
            out["ppo_rollout"] = {"error": repr(e)[:300]}
    if not args.no_denoiser:
        try:
            wd = with_denoiser(B, device, sd)
            if world > 1:
                t = torch.tensor([wd["ms_per_preview_batch"]], device=device)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                wd["ms_per_preview_batch"] = round(t.item(), 2)
                wd["value"] = round(world * B / (t.item() / 1e3), 2)
            out["with_denoiser"] = wd
        except Exception as e:  # noqa: BLE001  (the stand-in is optional; the solver numbers above stand on their own)
            out["with_denoiser"] = {"error": repr(e)[:200]}
    return out


# ------------------------------------------------------------------------------------------------------------
# helper: rank-0-only legs while the other ranks sleep
# ------------------------------------------------------------------------------------------------------------
def rank0_section_done(world, rank, tag):
    """Ranks != 0 block in a TCPStore wait (a sleeping socket read — NOT an NCCL collective, which would spin a host
    thread per rank and steal the cores the CPU arm is being timed on) until rank 0 has finished its rank-0-only legs."""
    if world <= 1:
        return
    import datetime

    store = torch.distributed.distributed_c10d._get_default_store()
    if rank == 0:
        store.set(tag, "1")
    else:
        store.wait([tag], datetime.timedelta(seconds=3600))


# ------------------------------------------------------------------------------------------------------------
# reference arms — the only place bench.py touches oracle/ (oracle/_ref = the unmodified reference, staged by
# oracle/stage_ref.py; consolver_oracle = the torch-CPU port, second row and fallback)
# ------------------------------------------------------------------------------------------------------------
def _reference_preview_fn(B, device="cpu", pool=1, compile_step=False):
    """8-step CFG preview through the UNMODIFIED reference PPOScheduler (scheduler_ppo.py:178-332) driven the way
    denoise_ppo.py:62-118 drives it: chunk, u + g*(c-u), scheduler.step(..., return_dict=False)[0], under no_grad, its
    per-step prints sent to /dev/null.  Returns None when the reference files are not available."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim

    if not ref_shim.reference_available():
        return None
    ref = ref_shim.load_reference()
    with ref_shim.quiet():
        sched = ref.PPOScheduler(factor_net_kwargs=dict(FN_KW), **SD_CFG)
    sched.factor_net.load_state_dict(policy_state_dict(0))
    sched.factor_net.to(device)
    sets = [tuple(t.to(device) for t in synth_batch(B, 1234 + j, None)) for j in range(pool)]
    step = sched.step
    if compile_step:
        import torch._dynamo as _dynamo

        _dynamo.config.cache_size_limit = 64
        step = torch.compile(sched.step)
    state = {"k": 0}

    def run():
        x, pairs = sets[state["k"] % pool]
        state["k"] += 1
        with ref_shim.devnull(), torch.no_grad():
            sched.set_timesteps(N_STEPS, device=device)
            lat = x
            for i, t in enumerate(sched.timesteps):
                u, c = pairs[i].chunk(2)                              # denoise_ppo.py:97
                eps = u + GUIDANCE * (c - u)                          # :100
                lat = step(eps, t, lat, return_dict=False)[0]         # :103
        return lat

    run.source = "oracle/_ref (staged copy)" if ref_shim.reference_is_staged_copy() else ref_shim.REFERENCE_ROOT
    return run


def _oracle_preview_fn(B):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import consolver_oracle as orc

    sd = policy_state_dict(0)
    sched = orc.OracleSDScheduler(sd, **{k: v for k, v in SD_CFG.items()})
    x, pairs = synth_batch(B, 1234, None)

    def run():
        sched.set_timesteps(N_STEPS)
        return orc.run_sd_preview(sched, x, list(pairs.unbind(0)), GUIDANCE)[0]

    return run


def _time_cpu(run, budget_s, max_n=2000):
    run()
    t0 = time.perf_counter()
    n, best = 0, float("inf")
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < max_n):
        t1 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t1)
        n += 1
    return n, time.perf_counter() - t0, best


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            return next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "")
    except OSError:
        return ""


def cpu_baseline(B, budget_s=10.0):
    """The reference scheduler itself on the host cores (kind "reference"), the torch-CPU oracle port as a second row;
    kind "port" only when oracle/_ref is absent."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref_run = _reference_preview_fn(B)
    port_run = _oracle_preview_fn(B)
    out = {}
    if ref_run is not None:
        n, dt, best = _time_cpu(ref_run, budget_s * 0.6)
        out = {"value": round(n * B / dt, 1), "unit": "previews/s", "cores": torch.get_num_threads(), "kind": "reference",
               "sample": f"{n} full 8-step previews of batch {B} (same workload unit, ~{dt:.1f} s of CPU work) through the "
                         f"UNMODIFIED reference PPOScheduler.step ({ref_run.source}) incl. the caller's CFG combine, "
                         f"prints to /dev/null",
               "best_ms_per_step": round(best * 1e3, 2),
               "algorithmic_gbs": round(TENSORS_PER_PREVIEW * B * 65536 / best / 1e9, 2)}
        n2, dt2, best2 = _time_cpu(port_run, budget_s * 0.25)
        out["port"] = {"value": round(n2 * B / dt2, 1), "best_ms_per_step": round(best2 * 1e3, 2),
                       "note": "torch-CPU oracle port of the same op sequence (no prints, no .item() syncs)"}
        one = ref_run
    else:
        n, dt, best = _time_cpu(port_run, budget_s * 0.8)
        out = {"value": round(n * B / dt, 1), "unit": "previews/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n} full 8-step previews of batch {B} (~{dt:.1f} s of CPU work), torch-CPU oracle port of the "
                         f"reference scheduler incl. CFG combine (oracle/_ref not staged on this box)",
               "best_ms_per_step": round(best * 1e3, 2),
               "algorithmic_gbs": round(TENSORS_PER_PREVIEW * B * 65536 / best / 1e9, 2)}
        one = port_run
    torch.set_num_threads(1)                       # second row: ONE host thread, a short bounded sample
    n1, dt1, _ = _time_cpu(one, min(2.0, budget_s / 6), max_n=200)
    torch.set_num_threads(cores)
    out.update(value_1_thread=round(n1 * B / dt1, 1), host_cpus=os.cpu_count(), cpu_model=_cpu_model())
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, same config / metric."""
    if rank != 0:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.batch
    run = _reference_preview_fn(B)
    kind = "reference"
    if run is None:
        run, kind = _oracle_preview_fn(B), "port"
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    v = round(args.steps * B / dt, 1)
    what = (f"UNMODIFIED reference PPOScheduler.step ({run.source}), prints to /dev/null" if kind == "reference"
            else "torch-CPU oracle port (oracle/_ref not staged)")
    return {"impl": "reference", "metric": "sd15_8step_solver_previews_per_s", "value": v, "unit": "previews/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
            "timed_region_s": round(dt, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world),
            "run": {"device": "host CPU", "threads": torch.get_num_threads(), "what": what},
            "cpu_baseline": {"value": v, "unit": "previews/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": f"{args.steps} full 8-step previews of batch {B}: {what}"},
            "e2e": {"value": v, "unit": "previews/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------------------------------
# the reference as stock torch ON THE GPU ("the bar to beat on B200")
# ------------------------------------------------------------------------------------------------------------
def torch_ref_gpu_leg(mode, B):
    """Runs in its own process (see torch_ref_gpu_subprocess).  mode: eager | compile.  Same workload as the headline:
    8-step CFG previews of batch B over a rotating pool of resident batches larger than L2, CUDA-event timed."""
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    out = {}
    if mode == "compile":
        out["inductor_host_compiler"] = _inductor_host_compiler()
    # the reference's arithmetic WITHOUT its host overheads: the op sequence of one steady-state step as a plain
    # function (tools/microbench.py::torch_reference_step) — the strongest stock-torch form of this path
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import microbench

        out["op_sequence_only"] = [microbench.bench_torch_reference(bb, mode) for bb in (B, 256)]
    except Exception as e:  # noqa: BLE001
        out["op_sequence_only"] = {"error": repr(e)[-300:]}
    try:
        out.update(_torch_ref_gpu_classes(mode, B, device))
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[-400:]
    return out


def _inductor_host_compiler():
    """torch-inductor compiles the reference's 0-d HOST-scalar arithmetic (alphas_cumprod[t] ** 0.5 ...) for the CPU with
    `$CXX -fopenmp`.  This image exports CXX=/opt/gcc/bin/g++, a trimmed gcc without libgomp.spec, while /usr/bin/g++
    has OpenMP: pick the first candidate that can build an OpenMP program; if none can, fall back to
    tools/inductor_cxx/g++-noomp (g++ minus -fopenmp, with a two-function omp.h stub)."""
    import shutil
    import subprocess
    import tempfile

    def works(cxx):
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "t.cpp")
            with open(src, "w") as f:
                f.write("#include <omp.h>\nint main(){return omp_get_max_threads() > 0 ? 0 : 1;}\n")
            try:
                return subprocess.run([cxx, "-fopenmp", "-shared", "-fPIC", src, "-o", os.path.join(d, "t.so")],
                                      capture_output=True, timeout=60).returncode == 0
            except Exception:  # noqa: BLE001
                return False

    cands = [c for c in (os.environ.get("CXX"), shutil.which("g++"), "/usr/bin/g++") if c]
    chosen = next((c for c in dict.fromkeys(cands) if works(c)), None)
    label = f"{chosen} -fopenmp" if chosen else "tools/inductor_cxx/g++-noomp (no candidate can build with -fopenmp)"
    if chosen is None:
        chosen = os.path.join(ROOT, "tools", "inductor_cxx", "g++-noomp")
    if chosen != os.environ.get("CXX"):
        label += f" (instead of CXX={os.environ.get('CXX')})"
    os.environ["CXX"] = chosen
    try:
        import torch._inductor.config as icfg

        icfg.cpp.cxx = (None, chosen)
    except Exception:  # noqa: BLE001
        pass
    return label


def _torch_ref_gpu_classes(mode, B, device):
    bytes_per_batch = (1 + 2 * N_STEPS) * B * SHAPE[0] * SHAPE[1] * SHAPE[2] * 4
    pool_n = int(max(2, -(-2 * L2_BYTES // bytes_per_batch)))
    t_build = time.perf_counter()
    run = _reference_preview_fn(B, device=device, pool=pool_n, compile_step=(mode == "compile"))
    if run is None:
        return {"unavailable": "oracle/_ref is not staged on this box (python oracle/stage_ref.py)"}
    for _ in range(3 if mode == "eager" else 2 * 4 + 2):      # compile: every history depth has to be traced once
        run()
    torch.cuda.synchronize(device)
    warm_s = time.perf_counter() - t_build
    n = 30
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        run()
    b.record()
    torch.cuda.synchronize(device)
    ms = a.elapsed_time(b) / n
    out = {"value": round(B / (ms / 1e3), 1), "unit": "previews/s", "ms_per_preview_batch": round(ms, 3),
           "us_per_solver_step": round(ms * 1e3 / N_STEPS, 1), "previews_timed": n, "batch": B, "pool_batches": pool_n,
           "warmup_s": round(warm_s, 1),
           "what": f"UNMODIFIED reference PPOScheduler.step + the caller's CFG combine on cuda:0, torch {mode}"
                   + (" (torch.compile(scheduler.step); graph breaks at its prints / .item() calls)" if mode == "compile" else
                      " (incl. its per-step host syncs and tensor prints, to /dev/null)")}
    return out


def torch_ref_gpu_subprocess(mode, B, local_index=0, timeout_s=420):
    import subprocess

    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_index]
               if os.environ.get("CUDA_VISIBLE_DEVICES") else str(local_index))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, os.path.abspath(__file__), "--leg", f"torch_ref_gpu:{mode}", "--batch", str(B)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        for ln in reversed(r.stdout.splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or r.stdout)[-300:]}
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {timeout_s} s"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: PPO rollout + update with the flat gradient all-reduce (train_ppo.py:257,:406-437)
# ------------------------------------------------------------------------------------------------------------
def ppo_rollout_leg(rank, world, device, batch=80, ppo_epochs=2, target_s=0.4):
    """Per iteration and per rank: one rollout of `batch` replicas of one (noise, target) pair with a random step count
    2..15 shared by all ranks (train_ppo.py:345), a synthetic latent-MSE reward, and `ppo_epochs` x (native loss+grad
    kernel, ONE ncclAllReduce(AVG) over the flat 75 041-float gradient buffer, clip, AdamW step).  The denoiser is a
    4x4 channel mix (not the product).  Reference: train_ppo.py:257,:345-437; edit_ppo/train_ppo.py:177,:275-283,:382."""
    import torch.distributed as dist

    import consolver_b200 as cb
    from consolver_b200 import ppo, sharding

    torch.manual_seed(1234 + rank)                            # per-rank seeds (edit_ppo/train_ppo.py:76)
    s = cb.PPOScheduler(factor_net_kwargs=dict(FN_KW), **SD_CFG)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    s.factor_net.to(device)
    flat = ppo.FlatParams(s.factor_net)
    ppo.broadcast_parameters(flat, 0)
    # the reference's optimizer (AdamW, train_ppo.py:223-229) at a tenth of its default rate: the synthetic reward below
    # carries no signal, and at 1e-4 several hundred updates of pure noise collapse the policy (ratios of 1e5 in the loss)
    opt = torch.optim.AdamW(s.factor_net.parameters(), lr=1e-5)
    g = torch.Generator(device=device).manual_seed(rank)
    w = torch.randn(4, 4, device=device, generator=g) * 0.3
    den = lambda x, t, i: torch.einsum("oc,bchw->bohw", w, x)  # noqa: E731  stand-in denoiser (not the product)
    noise = torch.randn(*SHAPE, device=device, generator=g)
    target = torch.randn(*SHAPE, device=device, generator=g)

    exchange, exchange_err = None, None
    if world > 1:
        try:                                                  # fused one-shot exchange over NVLink peer memory
            exchange = ppo.PeerGradExchange(flat)
        except Exception as e:  # noqa: BLE001  (no P2P / symmetric memory on this box: NCCL path)
            exchange_err = repr(e)[:200]
        ok = torch.tensor([1 if exchange is not None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            exchange = None

    # one CUDA graph of the whole sampling loop per step count (2..15), built before anything is timed
    rolls = ppo.GraphedRollouts(s, den, noise, batch, GUIDANCE, step_counts=range(2, 16))
    target_b = target.unsqueeze(0).expand(batch, *target.shape)

    def one(it, ex=None, read_back=False):
        n = ppo.shared_step_count(it, seed=0)
        lat, rec = rolls.rollout(n)
        r = ppo.latent_mse_reward(lat, target_b)
        return ppo.ppo_update(s.factor_net, flat, opt, rec, r, ppo_epochs=ppo_epochs, entropy_coef=0.01,
                              exchange=ex, read_back=read_back), n

    for it in range(3):
        one(it, exchange)
    ar_us = fused_us = plain_us = None
    if world > 1:
        # the update kernels alone: (loss+grad, reduce) + ncclAllReduce  vs  (loss+grad, reduce FUSED with the exchange)
        R_, A_ = 7, s.factor_net.action_dims
        xr = torch.tensor([[999.0 - 125 * r, 874.0 - 125 * r] for r in range(R_)], device=device)
        ix = torch.randint(0, FN_KW["num_actions"], (R_, batch, A_), device=device)
        ol = torch.rand(R_, batch, A_, device=device) * 0.5 + 0.05
        ad = torch.randn(R_, batch, A_, device=device)

        import ctypes

        from consolver_b200 import _lib

        lib = _lib.load()
        fnm = s.factor_net
        wk = fnm.kernel_weights()
        need = int(lib.consolver_ppo_workspace(R_, fnm.hidden_dim, A_, fnm.num_actions))
        wsp = torch.empty((need + 3) // 4, device=device)
        stt = torch.empty(4, device=device)
        strm = torch.cuda.current_stream(device).cuda_stream

        def time_updates(ex, nccl, n=200):
            """DEVICE time per update, timed with CUDA events over `n` back-to-back direct C-ABI calls (~5 us of host time
            each, below the ~55 us the two kernels take, so the GPU — not Python — sets the pace)"""
            def call(peers):
                rc = lib.consolver_ppo_loss_grad_allreduce_f32(
                    *wk[:6], xr.data_ptr(), R_, fnm.x_div, fnm.temperature, fnm.hidden_dim, A_, fnm.num_actions,
                    ix.data_ptr(), ol.data_ptr(), ad.data_ptr(), batch, 0.2, 0.01, wsp.data_ptr(), flat.grad.data_ptr(),
                    stt.data_ptr(), peers, strm)
                assert rc == 0, rc

            for _ in range(5):
                call(ctypes.byref(ex.next_peers()) if ex is not None else None)
                if nccl:
                    ppo.allreduce_gradients(flat)
            dist.barrier()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                call(ctypes.byref(ex.next_peers()) if ex is not None else None)
                if nccl:
                    ppo.allreduce_gradients(flat)
            e1.record()
            torch.cuda.synchronize(device)
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / n], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()

        plain_us = time_updates(None, False)
        nccl_us = time_updates(None, True)
        if exchange is not None:
            fused_us = time_updates(exchange, False)
        flat.grad.zero_()
    if world > 1:                                             # the collective alone, on the flat gradient buffer
        dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            ppo.allreduce_gradients(flat)
        e1.record()
        torch.cuda.synchronize(device)
        ar_us = e0.elapsed_time(e1) * 1e3 / 100
        flat.grad.zero_()
    # calibrate the iteration count for a region of >= target_s
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for it in range(3, 8):
        one(it, exchange)
    torch.cuda.synchronize(device)
    per = (time.perf_counter() - t0) / 5
    iters = int(max(10, min(2000, -(-target_s // max(per, 1e-5)))))
    if world > 1:
        t = torch.tensor([iters], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        iters = int(t.item())
        dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    steps = 0
    for it in range(8, 8 + iters):
        # the loss is read back every 10th iteration (the reference logs it every iteration, train_ppo.py:458; nothing
        # in the update depends on the read-back)
        st, n = one(it, exchange, read_back=(it % 10 == 9) or it == 8 + iters - 1)
        steps += n
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    stats = sharding.gather_job_stats(iters * batch, dt, flat.checksum(), device=device)
    return {"metric": "ppo_rollout_previews_per_s", "value": round(stats["total"] / stats["max_elapsed_s"], 1),
            "unit": "previews/s", "n_gpus": world, "iters": iters, "batch_per_gpu": batch, "ppo_epochs": ppo_epochs,
            "mean_solver_steps": round(steps / iters, 2), "timed_region_s": round(stats["max_elapsed_s"], 3),
            "allreduce_us_300KB": None if ar_us is None else round(ar_us, 2),
            "collective": None if world == 1 else (
                f"one-shot all-reduce(AVG) over NVLink peer memory FUSED into ppo_reduce_allreduce_kernel "
                f"({flat.numel} fp32 = {flat.numel * 4} B per rank, once per PPO epoch)" if exchange is not None else
                f"ncclAllReduce(AVG) over {flat.numel} fp32 ({flat.numel * 4} B), once per PPO epoch — latency-bound"),
            "exchange": None if world == 1 else {
                "used_in_timed_region": "fused peer-memory exchange" if exchange is not None else "ncclAllReduce",
                "timing": "CUDA events over 200 back-to-back direct C-ABI calls (device-bound), max over ranks",
                "update_kernels_us": None if plain_us is None else round(plain_us, 2),
                "update_kernels_plus_nccl_allreduce_us": None if plain_us is None else round(nccl_us, 2),
                "update_kernels_with_fused_exchange_us": None if fused_us is None else round(fused_us, 2),
                "nccl_allreduce_alone_us": None if ar_us is None else round(ar_us, 2),
                "fused_exchange_error": exchange_err},
            "grad_buffer_floats": flat.numel, "grad_buffer_bytes": flat.numel * 4,
            "param_checksum_identical_across_ranks": abs(stats["checksum"] / world - flat.checksum()) < 1e-6,
            "last_loss": st["loss"], "parameters_finite": bool(torch.isfinite(flat.flat).all()),
            "what": "BASELINE configs[4]: rollouts with random step counts 2-15 (one CUDA graph of the whole sampling "
                    "loop per step count, ppo.GraphedRollouts) + native PPO loss/grad kernel + flat gradient all-reduce "
                    "+ torch AdamW; loss read back every 10th iteration; stand-in denoiser = 4x4 channel mix, reward = "
                    "latent MSE"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--eager", action="store_true", help="python-eager launches instead of the CUDA graph")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--streams", type=int, default=4, help="independent preview batches in flight (CUDA streams)")
    ap.add_argument("--no-denoiser", action="store_true", help="skip the with_denoiser (U-Net stand-in) measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / roofline / cpu_baseline (profiling runs)")
    ap.add_argument("--no-torch-ref", action="store_true", help="skip the torch eager / compile reference-on-GPU legs")
    ap.add_argument("--no-ppo", action="store_true", help="skip the ppo_rollout leg (BASELINE configs[4])")
    ap.add_argument("--no-groups", action="store_true", help="replay previews one graph at a time (PreviewPool only)")
    ap.add_argument("--parallel-groups", dest="serial_groups", action="store_false",
                    help="one graph branch per preview inside a group instead of a serial chain")
    ap.add_argument("--pool", type=int, default=32,
                    help="minimum number of resident preview batches in the rotating pool (32 x 68 MiB = 17x the L2: measured "
                         "1.84 / 1.76 / 1.72 M previews/s at 8 / 16 / 32 — smaller pools get partial L2 hits)")
    ap.add_argument("--group-size", type=int, default=2, help="previews per group graph (0 = --streams)")
    ap.add_argument("--stagger-us", type=float, default=0.0,
                    help="phase shift between the group streams when the pipeline opens (a quarter of a preview)")
    ap.add_argument("--group-rotation", type=int, default=4, help="streams the preview groups are replayed on, round-robin")
    ap.add_argument("--reps", type=int, default=0, help="force the number of back-to-back blocks (0 = calibrate)")
    ap.add_argument("--leg", default="", help=argparse.SUPPRESS)       # internal: torch_ref_gpu:<mode> in a subprocess
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.leg.startswith("torch_ref_gpu:"):
        print(json.dumps(torch_ref_gpu_leg(args.leg.split(":", 1)[1], args.batch)), flush=True)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 50:
            args.steps = 50        # bounded sample: ~50 ms per step on a host CPU
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        from consolver_b200.sharding import bind_to_gpu_numa

        bind_to_gpu_numa(local)       # host buffers of the e2e leg stay on the GPU's own socket
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner out of stdout: one JSON line only
        torch.distributed.init_process_group("nccl", device_id=device)
    out = run_ours(args, rank, world, device)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
