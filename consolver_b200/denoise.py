"""Caller side of the hot path: the denoise loop that drives the scheduler (reference: denoise_ppo.py:6-120 for
SD1.5; edit_ppo/denoise_diffusion.py:101-157 for FLUX).  The denoiser itself (U-Net / DiT) is any callable and is
NOT the product; what this module changes relative to the reference loop:

  * CFG is fused into the scheduler step (`step_cfg`), so the three elementwise kernels of
    denoise_ppo.py:97-100 disappear;
  * the rollout record (conds / probs / actions / masks for steps i > 0, denoise_ppo.py:105-118) is a set of
    views into the scheduler's per-trajectory buffers — no per-step unsqueeze, no final cat, and the
    `conds['epsilon']` stack (hundreds of MiB, dead unless use_conv) is never built;
  * `GraphedPreview` captures the whole n-step solver loop (with a graph-safe RNG draw) in one CUDA graph.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch

from .scheduler_ppo import PPOScheduler


def denoise_loop(scheduler: PPOScheduler, denoiser: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor],
                 noise: torch.Tensor, cfg: float = 3.0, num_inference_steps: int = 50,
                 record: bool = True) -> Tuple[torch.Tensor, Optional[Dict[str, torch.Tensor]]]:
    """SD-style sampling loop with classifier-free guidance (denoise_ppo.py:52-120).

    `denoiser(latent_model_input [2B,...], t, i)` returns the noise prediction for the CFG-doubled batch
    (unconditional half first) — or for the plain batch when cfg <= 1.  Returns (latents, record) where record
    has the reference's layout: x [B,n-1,2], probs / actions / masks [B,n-1,A]."""
    scheduler.set_timesteps(num_inference_steps, device=noise.device)
    do_cfg = cfg > 1.0
    B = noise.shape[0]
    if not do_cfg:
        latents = noise.clone()
        for i, t in enumerate(scheduler.timesteps):
            pred = denoiser(scheduler.scale_model_input(latents, t), t, i)
            latents = scheduler.step(pred, t, latents, return_dict=False)[0]
        return latents, (scheduler.trajectory() if record and num_inference_steps > 1 else None)
    # CFG: two ping-pong [2B,...] denoiser inputs.  The step kernel writes x' into BOTH halves of the next input
    # (out / out2), so torch.cat([latents] * 2) (denoise_ppo.py:66: read N, write 2N per step) never runs.
    bufs = [noise.new_empty((2 * B, *noise.shape[1:])) for _ in range(2)]
    bufs[0][:B].copy_(noise)
    bufs[0][B:].copy_(noise)
    for i, t in enumerate(scheduler.timesteps):
        cur = bufs[i % 2]
        pred = denoiser(scheduler.scale_model_input(cur, t), t, i)
        # with a 16-bit denoiser output the latent may turn fp32 (torch promotion in the reference, see
        # PPOScheduler.next_latent_dtype): give the step a destination of the dtype it is going to return
        want = scheduler.next_latent_dtype(pred.dtype, cur.dtype)
        if bufs[(i + 1) % 2].dtype != want:
            bufs[(i + 1) % 2] = cur.new_empty(cur.shape, dtype=want)
        nxt = bufs[(i + 1) % 2]
        scheduler.step_cfg(pred, t, cur[:B], cfg, out=nxt[:B], out2=nxt[B:])
    latents = bufs[len(scheduler.timesteps) % 2][:B]
    return latents, (scheduler.trajectory() if record and num_inference_steps > 1 else None)


def denoise_loop_flow(scheduler, denoiser: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor],
                      initial_latents: torch.Tensor, image_latents: Optional[torch.Tensor] = None,
                      num_inference_steps: int = 8, sigmas=None, mu: Optional[float] = None,
                      record: bool = True) -> Tuple[torch.Tensor, Optional[Dict[str, torch.Tensor]]]:
    """Flow-matching sampling loop of the editing model (edit_ppo/denoise_diffusion.py:84-160): packed latents
    `[B, L, D]`, optionally followed along the sequence axis by the reference-image latents `[B, Li, D]`.

    `denoiser(latent_model_input [B, L(+Li), D], t, i)` returns the velocity for the whole sequence (or for the first L
    tokens); only the first L tokens are used (:140).  The step kernel writes the next latent straight into the head of
    the next `[B, L+Li, D]` transformer input (`out2`, strided by L+Li), whose image-latent tail was filled once, so the
    per-step `torch.cat([latents, image_latents], dim=1)` (:100 — read and write L+Li tokens) never runs.  Works with
    FMPPOScheduler (returns the rollout record: x [B,n-1,2], probs / actions / masks [B,n-1,A]) and with
    FlowMatchGeneralDiscreteScheduler (`use_naive_scheduler` in the reference; record is None)."""
    dev = initial_latents.device
    kw = {}
    if sigmas is not None:
        kw["sigmas"] = sigmas
    if mu is not None:
        kw["mu"] = mu
    scheduler.set_timesteps(num_inference_steps, device=dev, **kw)
    scheduler.set_begin_index(0)                      # edit_ppo/pipeline.py:1072; avoids the timestep search read-back
    B, L = initial_latents.shape[:2]
    learned = hasattr(scheduler, "factor_net")
    latents = initial_latents
    if image_latents is None:
        for i, t in enumerate(scheduler.timesteps):
            pred = denoiser(latents, t, i)
            latents = scheduler.step(pred, t, latents, return_dict=False)[0]
    else:
        Li = image_latents.shape[1]
        wide = [initial_latents.new_empty((B, L + Li, *initial_latents.shape[2:])) for _ in range(2)]
        for w in wide:
            w[:, L:].copy_(image_latents)
        wide[0][:, :L].copy_(initial_latents)
        for i, t in enumerate(scheduler.timesteps):
            pred = denoiser(wide[i % 2], t, i)
            if pred.shape[1] != L:
                pred = pred[:, :L]
            latents = scheduler.step(pred, t, latents, return_dict=False, out2=wide[(i + 1) % 2][:, :L])[0]
    rec = scheduler.trajectory() if (record and learned and num_inference_steps > 1) else None
    return latents, rec


def preview_from_pairs(scheduler: PPOScheduler, x_T: torch.Tensor, pairs: Sequence[torch.Tensor], guidance: float,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Solver-only preview: the denoiser is replaced by given per-step CFG pairs ([2B,...] each) — BASELINE
    config 0/2's 'random eps stand-in for the U-Net'.  `set_timesteps` must have been called."""
    x = x_T
    n = len(pairs)
    for i in range(n):
        x = scheduler.step_cfg(pairs[i], scheduler.timesteps[i], x, guidance, out=out if i == n - 1 else None)[0]
    return x


def preview_from_outputs(scheduler, x_T: torch.Tensor, outputs: Sequence[torch.Tensor]) -> torch.Tensor:
    """Solver-only preview through the plain `step()` contract (no CFG): works for PPOScheduler and FMPPOScheduler.
    `set_timesteps` (and, for FM, `set_begin_index`) must have been called."""
    x = x_T
    for i, mo in enumerate(outputs):
        x = scheduler.step(mo, scheduler.timesteps[i], x, return_dict=False)[0]
    return x


def _quiesce_before_capture(dev):
    """A stream capture in CUDA's default ("global") error mode is invalidated by ANY thread of the process making a
    capture-unsafe call — an NVML/clock sampler thread, or Python's cyclic GC freeing tensors of an earlier trajectory
    (the allocator then queries their stream events).  Collect garbage and drain the device first, and capture in
    "thread_local" mode (see the torch.cuda.graph calls below): other threads can no longer break a capture."""
    import gc

    gc.collect()
    torch.cuda.synchronize(dev)


class GraphedPreview:
    """One CUDA graph for a whole n-step solver-only preview over fixed device buffers.

    The graph holds one probability-table launch, then per step a sample kernel (drawing torch's Exp(1) stream
    itself from a device-resident generator state, see rng.py) and the fused step kernel, and a final one-thread
    node that advances that state: every replay consumes the default generator exactly as eager execution would.
    The sample kernels run on a side stream (a parallel branch of the graph: they depend only on the table and the
    generator state), and consecutive step kernels are PDL-chained (CONSOLVER_FLAG_CHAIN) because the model
    outputs are resident.  Inputs are read from the buffers given at capture time (refill them, or build one
    GraphedPreview per resident batch).  If the fused RNG self-check fails the graph falls back to torch's
    graph-safe exponential_ launch per step."""

    def __init__(self, scheduler, x_T: torch.Tensor, pairs: Sequence[torch.Tensor], guidance: Optional[float],
                 num_inference_steps: int, set_timesteps_kwargs: Optional[dict] = None):
        """`guidance` None: `pairs` are plain model outputs fed to `step()` (FMPPOScheduler, or SD without CFG);
        `set_timesteps_kwargs` (e.g. sigmas=..., mu=... for FM) are passed through."""
        self.scheduler = scheduler
        self.x_T, self.pairs, self.guidance, self.n = x_T, list(pairs), guidance, num_inference_steps
        self.out = None
        dev = x_T.device
        scheduler.set_timesteps(num_inference_steps, device=dev, **(set_timesteps_kwargs or {}))
        if hasattr(scheduler, "set_begin_index"):
            scheduler.set_begin_index(0)

        if getattr(scheduler, "use_fused_rng", False) and getattr(scheduler, "fixed_coefficients", None) is None:
            scheduler.policy_stream = torch.cuda.Stream(device=dev)   # sample chain becomes a parallel graph branch
            scheduler.chain_steps = True       # resident model outputs: consecutive steps overlap via PDL

        def run():
            if guidance is not None:
                return preview_from_pairs(scheduler, x_T, self.pairs, guidance, out=self.out)
            return preview_from_outputs(scheduler, x_T, self.pairs)

        self._run = run
        # warm-up on a side stream (allocates the trajectory buffers and the ring outside the capture)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            warm = run()
        torch.cuda.current_stream(dev).wait_stream(s)
        if guidance is not None:
            # fixed result buffer, in the dtype the trajectory ends in (a 16-bit pipeline's latent is fp32 from its
            # second step on — scheduler.next_latent_dtype — so it is not always x_T's)
            self.out = torch.empty_like(warm)
        del warm
        # graph-safe fused RNG: the sample kernels read {seed, offset} from a device buffer that replay() refreshes
        # from the default generator (and advances it), so every replay draws what eager execution would draw
        from . import _lib, rng as _rng

        tr = scheduler._traj
        self._dev = dev
        self._rng_inc = 0
        if scheduler.use_fused_rng and _rng.fused_rng_available(dev):
            tr.rng_plan = _lib.philox_plan(tr.q.numel())
            tr.graph_rng = torch.zeros(2, dtype=torch.int64, device=dev)
            self._pinned = torch.zeros(64, 2, dtype=torch.int64).pin_memory()
            self._pin_events = [None] * 64
            self._k = 0
        self._rewind()
        _quiesce_before_capture(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            res = run()
            if self.out is None:
                self.out = res          # lives in the graph's private pool: valid after every replay
            if tr.graph_rng is not None and tr.graph_rng_used:
                # last node: the device-resident offset advances by what this graph consumed, so back-to-back
                # replays need no host refresh at all
                self._rng_inc = tr.graph_rng_used * tr.rng_plan[1]
                _lib.check(_lib.load().consolver_rng_state_advance(
                    tr.graph_rng.data_ptr(), self._rng_inc, torch.cuda.current_stream(dev).cuda_stream), "rng advance")
        self._expected = None      # (seed, offset) the device state holds for the next replay
        # the two-stream / chained form is baked into the graph; eager calls on this scheduler go back to the
        # general single-stream form (CHAIN is only valid for resident model outputs)
        self.used_policy_stream = scheduler.policy_stream is not None
        self.used_chain = bool(scheduler.chain_steps)
        scheduler.policy_stream = None
        scheduler.chain_steps = False
        self._rewind()

    def _rewind(self):
        sch = self.scheduler
        sch._hist = []
        if hasattr(sch, "_step_count"):
            sch._step_count = 0
        if hasattr(sch, "_step_index"):
            sch._step_index = None         # FM: restart from begin_index
        if sch._traj is not None:
            sch._traj.rewind()             # the capture (and every replay) re-evaluates the probability tables

    def replay(self) -> torch.Tensor:
        if self._rng_inc:
            from . import rng as _rng

            seed, off = _rng.take(self._dev, self._rng_inc)     # torch's generator advances as eager code would
            if self._expected != (seed, off):
                # first replay, or somebody else used / reseeded the generator since: refresh the device state
                j = self._k % 64
                self._k += 1
                if self._pin_events[j] is not None:
                    self._pin_events[j].synchronize()                 # the copy that last used this slot is done
                self._pinned[j, 0] = seed - (1 << 64) if seed >= (1 << 63) else seed
                self._pinned[j, 1] = off
                self.scheduler._traj.graph_rng.copy_(self._pinned[j], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                self._pin_events[j] = ev
            self._expected = (seed, off + self._rng_inc)
        self.graph.replay()
        return self.out

    def record(self):
        """rollout record of the last replay (views; valid until the next replay)"""
        tr = self.scheduler._traj
        tr.count = self.n
        rec = self.scheduler.trajectory()
        return rec


class PreviewPool:
    """Keeps several independent GraphedPreview objects (different prompts / seeds — no dependency between them) in
    flight on separate CUDA streams, so one batch's launch ramp and tail overlap another's streaming phase.  At the
    SD1.5 shape with batch 64 this lifts the 8-step solver loop from 0.97 M to 1.67 M previews/s on one B200
    (≈0.97 of the HBM copy peak for the whole loop).  Preview j always runs on stream j % n_streams, so two replays
    of the same preview never overlap; the default generator is consumed in submission order."""

    def __init__(self, previews: Sequence[GraphedPreview], streams: int = 4, stagger_us: float = 0.0):
        """`stagger_us`: when the pool opens (first submit after a join) stream j is delayed by j * stagger_us.  Equal
        graphs replayed round-robin by a fast host run in LOCKSTEP — all streams in their latency-bound phases (first
        steps, policy->step hand-offs) at the same moment, then all in their streaming phases — which wastes the overlap
        several streams are there to provide; a one-off phase shift of a fraction of a preview keeps them interleaved."""
        self.stagger_us = float(stagger_us)
        self.previews = list(previews)
        n = max(1, min(streams, len(self.previews)))
        while len(self.previews) % n:
            n -= 1
        dev = self.previews[0].x_T.device
        self.main = torch.cuda.current_stream(dev)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n)] if n > 1 else [self.main]
        self._open = False

    def submit(self, j: int) -> torch.Tensor:
        """enqueue one replay of preview j; returns its output buffer (valid after join() / stream order)"""
        if not self._open and len(self.streams) > 1:
            for i, st in enumerate(self.streams):
                st.wait_stream(self.main)
                if self.stagger_us > 0 and i:
                    with torch.cuda.stream(st):
                        torch.cuda._sleep(int(i * self.stagger_us * 1.9e3))      # ~1.9 cycles per ns at B200 clocks
            self._open = True
        st = self.streams[j % len(self.streams)]
        if st is self.main:
            return self.previews[j].replay()
        with torch.cuda.stream(st):
            return self.previews[j].replay()

    def join(self):
        """make the current stream wait for everything submitted so far"""
        if len(self.streams) > 1:
            for st in self.streams:
                self.main.wait_stream(st)
        self._open = False


class PreviewGroup:
    """SEVERAL independent previews (different prompts / seeds) captured as ONE CUDA graph: the graph forks one branch
    per preview (the branches have no dependency on each other, so the GPU overlaps one preview's launch ramp and tail
    with another's streaming phase exactly as PreviewPool does with streams), joins them, and ends with one node that
    advances a device-resident generator state SHARED by all branches.

    Why: replaying previews one graph at a time costs the host ~25-30 us per preview (stream switch, generator
    bookkeeping, a state refresh whenever replays of different previews interleave, the graph launch itself) — as much
    as the 35 us of GPU time a batch-64 preview takes.  A group of g previews is ONE launch and ONE generator update, so
    the host cost per preview drops by g and the loop is GPU-bound with a wide margin even when eight ranks share one
    host.  The default generator is consumed exactly as g eager previews in order would consume it: branch j draws at
    offset base + j * (what one preview consumes), so results are bit-identical to serial eager execution.

    Built from existing GraphedPreview objects (their buffers, schedulers and eager closures); those stay usable."""

    def __init__(self, previews: Sequence[GraphedPreview], rotation: int = 1, parallel: bool = True):
        """`parallel`: True — one graph branch per preview (the previews of ONE replay overlap; the graph's implicit join
        at its end drains the GPU once per replay).  False — the previews are captured one after the other on a single
        branch: a replay is then a plain chain, and several such groups replayed on separate streams (PreviewPool) keep
        every stream busy back to back with no join anywhere — the form with the highest steady-state throughput.
        `rotation`: how many groups of this size are replayed round-robin (A, B, A, B, ... with rotation 2).  The
        device-resident generator state then advances by the WHOLE rotation's consumption per replay, so that in steady
        round-robin order it already holds the right offset at the next replay and the host never has to refresh it;
        any other order is detected (replay() compares with the default generator) and costs one small refresh."""
        from . import _lib, rng as _rng

        self.previews = list(previews)
        if not self.previews:
            raise ValueError("PreviewGroup needs at least one preview")
        dev = self._dev = self.previews[0].x_T.device
        incs = {p._rng_inc for p in self.previews}
        if len(incs) != 1 or 0 in incs:
            raise ValueError("PreviewGroup needs previews that use the fused RNG and consume the same amount per replay")
        self._inc_one = incs.pop()
        self._inc = self._inc_one * len(self.previews)
        self._advance = self._inc * max(1, int(rotation))
        self.shared = torch.zeros(2, dtype=torch.int64, device=dev)
        self._pinned = torch.zeros(16, 2, dtype=torch.int64).pin_memory()
        self._pin_events = [None] * 16
        self._k = 0
        self._expected = None
        branches = [torch.cuda.Stream(device=dev) for _ in self.previews]
        saved = []
        for j, p in enumerate(self.previews):
            sch, tr = p.scheduler, p.scheduler._traj
            saved.append((tr.graph_rng, sch.policy_stream, sch.chain_steps))
            p._rewind()
            tr.graph_rng = self.shared                       # every branch reads the one shared {seed, offset}
            per_draw = tr.rng_plan[1]
            tr.graph_rng_used = j * (self._inc_one // per_draw)   # ... at its own place in the generator stream
            if p.used_policy_stream:
                sch.policy_stream = torch.cuda.Stream(device=dev)
            sch.chain_steps = p.used_chain
        _quiesce_before_capture(dev)
        self.graph = torch.cuda.CUDAGraph()
        self.outs = []
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            main = torch.cuda.current_stream(dev)
            for p, st in zip(self.previews, branches):
                if parallel:
                    st.wait_stream(main)
                    with torch.cuda.stream(st):
                        res = p._run()
                else:
                    res = p._run()
                # with CFG pairs the result is written into the preview's own `out` buffer; through plain step() it
                # is the last tensor the loop allocated — from THIS graph's pool — so that is what holds the result
                self.outs.append(p.out if p.guidance is not None else res)
            if parallel:
                for st in branches:
                    main.wait_stream(st)
            _lib.check(_lib.load().consolver_rng_state_advance(self.shared.data_ptr(), self._advance, main.cuda_stream),
                       "rng advance")
        for p, (own_rng, ps, chain) in zip(self.previews, saved):
            sch, tr = p.scheduler, p.scheduler._traj
            tr.graph_rng, sch.policy_stream, sch.chain_steps = own_rng, None, False
            p._rewind()
        del _rng

    def __len__(self):
        return len(self.previews)

    @property
    def x_T(self):                      # lets a PreviewPool of groups find the device
        return self.previews[0].x_T

    def replay(self):
        """enqueue one replay of all previews of the group on the current stream; returns their output buffers.
        Replays on ONE stream serialise (each waits for the previous group's join): to keep the GPU full across group
        boundaries put two or more groups into a PreviewPool, one stream each."""
        from . import rng as _rng

        seed, off = _rng.take(self._dev, self._inc)          # torch's generator advances as g eager previews would
        if self._expected != (seed, off):
            # first replay, or somebody else used / reseeded the generator since: refresh the shared device state
            j = self._k % 16
            self._k += 1
            if self._pin_events[j] is not None:
                self._pin_events[j].synchronize()
            self._pinned[j, 0] = seed - (1 << 64) if seed >= (1 << 63) else seed
            self._pinned[j, 1] = off
            self.shared.copy_(self._pinned[j], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pin_events[j] = ev
        self._expected = (seed, off + self._advance)     # what the device state holds after this replay
        self.graph.replay()
        return self.outs


class GraphedDenoiseLoop:
    """The WHOLE n-step CFG sampling loop — denoiser forward passes included — as one CUDA graph over static
    buffers.  For interactive previews (batch 1-4) the denoiser is launch-bound, so replaying one
    graph instead of ~10^4 eager launches is where the latency goes; the scheduler's part is the same kernels as in
    `denoise_loop` (fused CFG + step writing x' into both halves of the next denoiser input, sample kernels on a
    side stream, in-kernel RNG from a device-resident generator state).

    `denoiser(model_in [2B,...], t, i)` must be capturable (static shapes, no host sync); `t` is the python int of
    step i.  Fill `self.noise` (or pass `noise=` to `replay`) and call `replay()`; the result is `self.latents`."""

    def __init__(self, scheduler: PPOScheduler, denoiser: Callable, noise: torch.Tensor, cfg: float,
                 num_inference_steps: int):
        from . import _lib, rng as _rng

        self.scheduler, self.denoiser, self.cfg, self.n = scheduler, denoiser, float(cfg), num_inference_steps
        dev = noise.device
        B = noise.shape[0]
        self.noise = noise.clone()
        self._bufs = [noise.new_empty((2 * B, *noise.shape[1:])) for _ in range(2)]
        scheduler.set_timesteps(num_inference_steps, device=dev)
        ts = [int(t) for t in scheduler._timesteps_host]
        if scheduler.use_fused_rng and scheduler.fixed_coefficients is None:
            scheduler.policy_stream = torch.cuda.Stream(device=dev)

        def run():
            bufs = self._bufs
            bufs[0][:B].copy_(self.noise)
            bufs[0][B:].copy_(self.noise)
            for i, t in enumerate(ts):
                cur, nxt = bufs[i % 2], bufs[(i + 1) % 2]
                pred = self.denoiser(cur, t, i)
                scheduler.step_cfg(pred, t, cur[:B], self.cfg, out=nxt[:B], out2=nxt[B:])
            return bufs[len(ts) % 2][:B]

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            run()
            run()                                     # second pass: library autotuning / lazy allocations settle
        torch.cuda.current_stream(dev).wait_stream(side)
        tr = scheduler._traj
        self._dev, self._rng_inc, self._expected = dev, 0, None
        if scheduler.use_fused_rng and _rng.fused_rng_available(dev):
            tr.rng_plan = _lib.philox_plan(tr.q.numel())
            tr.graph_rng = torch.zeros(2, dtype=torch.int64, device=dev)
        self._rewind()
        _quiesce_before_capture(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"), torch.no_grad():
            self.latents = run()
            if tr.graph_rng is not None and tr.graph_rng_used:
                self._rng_inc = tr.graph_rng_used * tr.rng_plan[1]
                _lib.check(_lib.load().consolver_rng_state_advance(
                    tr.graph_rng.data_ptr(), self._rng_inc, torch.cuda.current_stream(dev).cuda_stream), "rng advance")
        scheduler.policy_stream = None
        self._rewind()

    def _rewind(self):
        sch = self.scheduler
        sch._hist = []
        sch._step_count = 0
        if sch._traj is not None:
            sch._traj.rewind()

    def replay(self, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if noise is not None:
            self.noise.copy_(noise)
        if self._rng_inc:
            from . import rng as _rng

            seed, off = _rng.take(self._dev, self._rng_inc)
            if self._expected != (seed, off):
                state = torch.tensor([seed - (1 << 64) if seed >= (1 << 63) else seed, off], dtype=torch.int64)
                self.scheduler._traj.graph_rng.copy_(state)       # pageable source: staged synchronously, rare
            self._expected = (seed, off + self._rng_inc)
        self.graph.replay()
        return self.latents

    def record(self):
        tr = self.scheduler._traj
        tr.count = self.n
        return self.scheduler.trajectory()
