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
  cpu_baseline / --impl reference   the CPU oracle port of the reference scheduler on the host cores
"""
from __future__ import annotations

import argparse
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
HIST_DEPTHS = [1, 2, 3, 4, 4, 4, 4, 4]                     # n_hist per step of an 8-step preview
TENSORS_PER_PREVIEW = sum(n + 4 for n in HIST_DEPTHS)       # 58 latent-sized transfers / sample


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


def fm_preview_throughput(device, B=16, steps=200):
    """BASELINE configs[3]: FLUX-Kontext-shaped FMPPOScheduler loop — packed latents [B,4096,64] bf16, 8 steps,
    order_dim=2, resident synthetic velocities as the transformer stand-in, one CUDA graph per preview."""
    import numpy as np

    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview

    torch.manual_seed(0)
    nbytes = (1 + N_STEPS) * B * 4096 * 64 * 2
    pool_n = max(8, int(-(-2 * L2_BYTES // nbytes)))
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
    from consolver_b200.denoise import PreviewPool

    pp = PreviewPool(pool, streams=4)
    for k in range(2 * pool_n):
        pp.submit(k % pool_n)
    pp.join()
    torch.cuda.synchronize(device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(steps):
        pp.submit(k % pool_n)
    pp.join()
    b.record()
    torch.cuda.synchronize(device)
    ms = a.elapsed_time(b) / steps
    tensors = 3 + 7 * 4     # n = 1, then 2: (n+2) tensors per step
    return {"value": round(B / (ms / 1e3), 1), "unit": "previews/s", "batch": B, "ms_per_preview_batch": round(ms, 4),
            "shape": "[B,4096,64] bf16, 8 steps, order_dim=2", "algorithmic_gbs": round(
                tensors * B * 4096 * 64 * 2 / (ms / 1e3) / 1e9, 1), "pool_batches": pool_n,
            "concurrency": f"{len(pp.streams)} preview batches in flight"}


def run_ours(args, rank, world, device):
    import consolver_b200  # noqa: F401  (fails loudly if libconsolver.so cannot be built/loaded)
    from consolver_b200.denoise import GraphedPreview, preview_from_pairs

    B = args.batch
    sd = policy_state_dict(0)
    bytes_per_batch = (1 + 2 * N_STEPS) * B * SHAPE[0] * SHAPE[1] * SHAPE[2] * 4
    pool_n = int(max(2, -(-2 * L2_BYTES // bytes_per_batch)))
    if not args.eager:
        pool_n = max(pool_n, 2 * max(1, args.streams))     # two resident batches per stream
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

    def run_steps(first, count):
        if ppool is None:
            for k in range(first, first + count):
                one_step(k)
            return
        for k in range(first, first + count):
            ppool.submit(k % pool_n)
        ppool.join()

    run_steps(0, args.warmup)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(torch.cuda.current_device()) as clk:
        a.record()
        run_steps(args.warmup, args.steps)
        b.record()
        barrier()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = t.item()
    value = world * args.steps * B / (ms / 1e3)

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
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4)}
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

    e2e_steps = max(10, min(args.steps, 200))
    for k in range(4):
        e2e_step(k)
    e2e_drain()
    barrier()
    with ClockSampler(torch.cuda.current_device()) as clk_e2e:     # the second timed region: sampled as well
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            e2e_step(k)
        e2e_drain()
        barrier()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_val = world * e2e_steps * B / e2e_s
    h2d = (hx.numel() + hpairs.numel()) * 4
    d2h = (e2e_sets[0]["hout"].numel() + e2e_sets[0]["hrec"].numel()) * 4

    out = {
        "metric": "sd15_8step_solver_previews_per_s", "value": round(value, 1), "unit": "previews/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] solver path: SD1.5 latents 4x64x64 fp32, 8-step trailing, CFG=3, "
                               "order_dim=4, scaler_dim=0, policy 2-256-256-33, batch 64/GPU, denoiser = resident "
                               "synthetic CFG pairs", "batch_per_gpu": B, "solver_steps": N_STEPS,
                   "l2_policy": f"inputs larger than L2: rotating pool of {pool_n} resident batches "
                                f"({pool_n * bytes_per_batch >> 20} MiB)",
                   "launch": "eager python launches" if args.eager else
                   "one CUDA graph per 8-step preview (table kernel, sample kernels on a side stream, PDL-chained step "
                   "kernels, rng-advance node); valid because the stand-in model outputs are resident", "sharding": f"dp{world} by prompt/seed, no collective",
                   "concurrency": f"{n_streams} independent preview batch(es) in flight on separate CUDA streams"},
        "e2e": {"value": round(e2e_val, 1), "unit": "previews/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": args.steps * (2 + N_STEPS * 2),   # per preview: table + 8 x (sample + step) + rng-advance kernels
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
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")
        if os.path.exists(tpath):          # DRAM bytes per launch from the committed ncu --set full capture
            rec = json.load(open(tpath))["step_kernel_f32_nh4_pair"].get(str(B))
            if rec:
                traffic = rec["dram_read"] + rec["dram_write"]
        out["roofline"] = {"bound": "hbm", "kernel": "step_kernel<f32,NH=4,CFG pair>", "batch": B,
                           "achieved": round(nbytes / us / 1e3, 1), "peak": peak, "unit": "GB/s",
                           "frac": round(nbytes / us / 1e3 / peak, 4), "traffic": traffic, "us_per_launch": round(us, 3),
                           "algorithmic_bytes": nbytes, "peak_source": src}
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
            sweep.append({"batch": Bs, "us_per_launch": round(us, 3), "achieved": round(nbytes / us / 1e3, 1),
                          "frac": round(nbytes / us / 1e3 / peak, 4)})
        out["roofline_sweep"] = sweep
        out["solver_loop"]["frac_of_peak"] = round(out["solver_loop"]["achieved_gbs_per_gpu"] / peak, 4)
        fm = []
        for Bs in (1, 8, 64, 512):
            us, nbytes = time_fm_kernel(Bs, device)
            fm.append({"batch": Bs, "us_per_launch": round(us, 3), "achieved": round(nbytes / us / 1e3, 1),
                       "frac": round(nbytes / us / 1e3 / peak, 4)})
        out["roofline_fm_flux_bf16"] = {"kernel": "step_kernel<bf16,NH=2,FM>", "shape": "[B,4096,64] bf16 (FLUX-Kontext "
                                        "packed 1024^2 latents), order_dim=2", "bytes_per_sample": 4 * 4096 * 64 * 2,
                                        "points": fm}
        out["cpu_baseline"] = cpu_baseline(B, budget_s=args.cpu_budget)
        try:
            out["fm_flux_preview"] = fm_preview_throughput(device)
        except Exception as e:  # noqa: BLE001
            out["fm_flux_preview"] = {"error": repr(e)[:200]}
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
# CPU arm (oracle port of the reference) — the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------------------
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


def cpu_baseline(B, budget_s=10.0):
    torch.set_num_threads(os.cpu_count() or 1)
    run = _oracle_preview_fn(B)
    run()
    t0 = time.perf_counter()
    n = 0
    best = float("inf")
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < 2000):
        t1 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t1)
        n += 1
    dt = time.perf_counter() - t0
    cores = torch.get_num_threads()
    # second row: the same port on ONE host thread, a short bounded sample
    torch.set_num_threads(1)
    t1 = time.perf_counter()
    n1 = 0
    while n1 < 2 or (time.perf_counter() - t1 < min(2.0, budget_s / 4) and n1 < 200):
        run()
        n1 += 1
    dt1 = time.perf_counter() - t1
    torch.set_num_threads(cores)
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "")
    except OSError:
        pass
    return {"value": round(n * B / dt, 1), "unit": "previews/s", "cores": cores, "kind": "port",
            "sample": f"{n} full 8-step previews of batch {B} (same workload unit, ~{dt:.1f} s of CPU work), torch-CPU "
                      f"oracle port of the reference scheduler incl. CFG combine, no debug prints",
            "best_ms_per_step": round(best * 1e3, 2), "algorithmic_gbs": round(
                TENSORS_PER_PREVIEW * B * 65536 / best / 1e9, 2),
            "value_1_thread": round(n1 * B / dt1, 1), "host_cpus": os.cpu_count(), "cpu_model": model}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.batch
    run = _oracle_preview_fn(B)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    v = round(args.steps * B / dt, 1)
    return {"impl": "reference", "metric": "sd15_8step_solver_previews_per_s", "value": v, "unit": "previews/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] solver path on the host CPU (oracle port of the reference "
                                   "PPOScheduler, torch-CPU ops): SD1.5 latents 4x64x64 fp32, 8-step, CFG=3, batch 64",
                       "batch_per_gpu": B, "solver_steps": N_STEPS},
            "cpu_baseline": {"value": v, "unit": "previews/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.steps} full 8-step previews of batch {B}"},
            "e2e": {"value": v, "unit": "previews/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--eager", action="store_true", help="python-eager launches instead of the CUDA graph")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--streams", type=int, default=4, help="independent preview batches in flight (CUDA streams)")
    ap.add_argument("--no-denoiser", action="store_true", help="skip the with_denoiser (U-Net stand-in) measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / roofline / cpu_baseline (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
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
