"""PPOScheduler — drop-in for the reference's `scheduler_ppo.PPOScheduler` (scheduler_ppo.py:48-361): same
constructor kwargs, `set_timesteps()` / `step()` signatures and 5-tuple return, same `factor_net` state_dict —
with the per-step work done by two hand-written sm_100a kernels behind the C ABI (include/consolver.h):

  policy kernel  MLP once per step + softmax + per-sample categorical draw + coefficient/mask assembly
  step kernel    CFG combine + linear-multistep combine over the in-place history ring + DDIM update,
                 one pass over HBM

What changes relative to the reference, none of it numerical:
  * no host synchronisation inside `step()`: per-timestep scalars are tabulated on the host at construction,
    the timestep value comes from the host copy of the grid (see `sync_free`), nothing is printed;
  * the history is a ring of references / scheduler-owned slots, never stacked; `conds['epsilon']` is
    materialised lazily (config_utils.LazyConds);
  * `step_cfg()` (new, optional) takes the raw [2B,...] CFG pair and the guidance scale so the caller's
    `u + g*(c-u)` (denoise_ppo.py:96-100) is fused into the same pass;
  * per-step `actions / probs / masks` are views into per-trajectory buffers (`trajectory()` returns the
    [B, n-1, A] record denoise_ppo.py:105-118 builds with unsqueeze+cat).
There is no CPU path: inputs must be CUDA tensors and libconsolver.so must load.
"""
from __future__ import annotations

import ctypes
import dataclasses
import math
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .config_utils import KARRAS_COMPATIBLES, BaseOutput, ConfigMixin, LazyConds, SchedulerMixin, register_to_config
from .factor_net import FactorNetPPO, alloc_policy_outputs


@dataclasses.dataclass
class PPOSchedulerOutput(BaseOutput):
    """`return_dict=True` result.  (The reference builds a diffusers SchedulerOutput with five fields,
    scheduler_ppo.py:299, which raises with stock diffusers; this carries the same five fields.)"""
    prev_sample: torch.Tensor = None
    actions: Optional[torch.Tensor] = None
    probs: Optional[torch.Tensor] = None
    conds: Optional[Dict] = None
    masks: Optional[torch.Tensor] = None


def _cosine_alpha_bar_betas(n: int, max_beta: float = 0.999) -> torch.Tensor:
    bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
    return torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), max_beta) for i in range(n)], dtype=torch.float32)


def next_rng(sched, tr, device):
    """consolver_rng_t for this step's in-kernel draw, or None to use the torch exponential_ launch."""
    from . import rng as _rng

    if not sched.use_fused_rng:
        return None
    if torch.cuda.is_current_stream_capturing():
        if tr.graph_rng is None or tr.rng_plan is None:
            return None
        nthreads, inc = tr.rng_plan
        r = _lib.Rng(0, tr.graph_rng_used * inc, tr.graph_rng.data_ptr(), nthreads)
        tr.graph_rng_used += 1
    else:
        if not _rng.fused_rng_available(device):
            return None
        if tr.rng_plan is None:
            tr.rng_plan = _lib.philox_plan(tr.q.numel())
        nthreads, inc = tr.rng_plan
        seed, off = _rng.take(device, inc)
        r = _lib.Rng(seed, off, None, nthreads)
    tr._rng_keepalive = r
    return r


class _Trajectory:
    """Per-(set_timesteps, batch shape) device state: policy outputs for every step, the Exp(1) buffer, the
    history ring.  Allocated once; `step()` itself allocates only the returned latent."""

    def __init__(self, sched: "PPOScheduler", B, shape, dtype, device):
        fn = sched.factor_net_module
        self.key = (B, tuple(shape), dtype, device)
        self.n = max(int(sched.num_inference_steps), 1)
        A, K, od = fn.action_dims, fn.num_actions, sched.config.order_dim
        self.out = alloc_policy_outputs(B, A, K, od, device, lead=(self.n,))
        # row pointers by arithmetic: indexing a tensor costs ~2 us of host time, a step needs seven of them
        self._row = {k: (v.data_ptr(), v.stride(0) * v.element_size()) for k, v in self.out.items()}
        self.q = torch.empty((B * A, K), device=device, dtype=torch.float32)
        self.ring = None  # [order_dim, B, *shape], allocated on the first step_cfg()
        self.ring_shape = (od, B, *shape)
        self.dtype, self.device = dtype, device
        # conds['x'] rows for the whole grid: one H2D copy per trajectory instead of one per step
        ts = sched._timesteps_host
        rows = [[float(t), float(t - sched._stride)] for t in ts]
        host = torch.tensor(rows, dtype=dtype)
        self.condx = host.to(device, non_blocking=True)
        self.condx_f32 = host.float().to(device, non_blocking=True)   # policy-table kernel input [n,2]
        self.condx_host = host.float().numpy()
        self.count = 0
        self.table_pass = -1   # trajectory pass (count // n) whose probability tables are in out['probs_table']
        self.rng_plan = None   # (nthreads, offset increment) of torch's exponential_ launch for q's numel
        self.graph_rng = None  # device int64[2] {seed, offset} refreshed before every CUDA-graph replay
        self.graph_rng_used = 0
        self.policy_forked = False   # the policy side stream has been forked off the main stream in this pass

    def p(self, name, i):
        base, stride = self._row[name]
        return base + i * stride

    def conv_buffers(self, fn):
        """use_conv=True scratch: features [B,od-1], reduction workspace, per-sample tables [n,B,A,K]"""
        if getattr(self, "_conv", None) is None:
            from .features import workspace_bytes

            B, od = self.key[0], self.ring_shape[0]
            self._conv = (torch.empty(B, od - 1, device=self.device, dtype=torch.float32),
                          torch.empty(workspace_bytes(B, od) // 8 + 1, device=self.device, dtype=torch.float64),
                          torch.empty(self.n, B, fn.action_dims, fn.num_actions, device=self.device,
                                      dtype=torch.float32))
        return self._conv

    def slot(self, i):
        if self.ring is None:
            self.ring = torch.empty(self.ring_shape, device=self.device, dtype=self.dtype)
        return self.ring[i % self.ring_shape[0]]


class PPOScheduler(SchedulerMixin, ConfigMixin):
    """Learned linear-multistep DDIM-form solver (ConsistencySolver) for eps / v-prediction models."""

    _compatibles = KARRAS_COMPATIBLES
    order = 1

    @register_to_config
    def __init__(
        self,
        num_train_timesteps: int = 1000,
        beta_start: float = 0.0001,
        beta_end: float = 0.02,
        beta_schedule: str = "linear",
        trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
        prediction_type: str = "epsilon",
        timestep_spacing: str = "leading",
        steps_offset: int = 0,
        order_dim: int = 4,
        scaler_dim: int = 2,
        use_conv=False,
        ppo_type="discrete",
        factor_net_kwargs: Optional[Dict] = None,
    ):
        # noise schedule (scheduler_ppo.py:99-114), fp32 on the host like the reference
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = _cosine_alpha_bar_betas(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} schedule not implemented.")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        # sqrt tables the kernel scalars are read from: the same fp32 ops as scheduler_ppo.py:309-330
        # (`** 0.5` of abar_t and of 1 - abar_t), evaluated for every t once instead of per step
        self._sqrt_abar = (self.alphas_cumprod ** 0.5).numpy()
        self._sqrt_1m_abar = ((1 - self.alphas_cumprod) ** 0.5).numpy()

        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self._timesteps_host = np.arange(0, num_train_timesteps)[::-1].copy()
        self.timesteps = torch.from_numpy(self._timesteps_host)

        if order_dim < 2 or order_dim > _lib.MAX_ORDER:
            raise ValueError(f"order_dim must be in [2, {_lib.MAX_ORDER}]")
        if scaler_dim not in (0, 1, 2):
            raise NotImplementedError("more than two scale parameters are not supported")  # scheduler_ppo.py:279
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(f"Unsupported prediction_type: {prediction_type}")
        kw = dict(factor_net_kwargs) if factor_net_kwargs is not None else {}
        kw.update(order_dim=order_dim, scaler_dim=scaler_dim, use_conv=use_conv)
        kw.setdefault("embedding_dim", 32)
        kw.setdefault("hidden_dim", 256)
        if ppo_type != "discrete":
            # scheduler_ppo.py:139 instantiates FactorNetPPOContinous, whose source is not part of the reference
            raise NotImplementedError("ppo_type != 'discrete': the continuous policy does not exist in the reference")
        kw.setdefault("num_actions", 161)
        self.factor_net = FactorNetPPO(**kw)

        self._hist: List[torch.Tensor] = []   # model outputs, NEWEST FIRST (references or ring slots)
        self._traj: Optional[_Trajectory] = None
        self._step_count = 0
        self._stride = 0
        #: True (default): a CUDA `timestep` tensor is NOT read back; the value is taken from the host copy of
        #: the grid at the current step count (pipelines step in grid order).  False: `.item()` it (one sync).
        self.sync_free = True
        #: link the policy and step kernels with programmatic dependent launch
        self.use_pdl = True
        #: generate the Exp(1) draw inside the sample kernel (bit-identical to torch's exponential_, see rng.py)
        self.use_fused_rng = True
        #: optional side stream for the policy kernels (set by GraphedPreview): see the two-stream note in _step
        self.policy_stream: Optional[torch.cuda.Stream] = None
        #: solver-only replays (GraphedPreview): chain consecutive step kernels with programmatic dependent launch
        #: (CONSOLVER_FLAG_CHAIN).  Only valid when the model outputs are NOT produced by the kernel right before.
        self.chain_steps = False
        #: replay instead of sampling: {'idx': seq of [B,A] int64 per step} forces the bins (PPO replay, parity
        #: tests with injected actions); {'q': seq of [B*A,K] fp32 per step} supplies the Exp(1) draw.
        self.replay: Optional[Dict] = None
        #: None, or a callable n_hist -> sequence of n_hist multipliers (newest first, summing to 1): the policy is
        #: bypassed and the fused step runs with these fixed coefficients (see consolver_b200.baselines)
        self.fixed_coefficients = None

    # ------------------------------------------------------------------------------------------------------
    @property
    def factor_net_module(self) -> FactorNetPPO:
        fn = self.factor_net
        return fn.module if hasattr(fn, "module") else fn   # DDP-wrapped during training (scheduler_ppo.py:239)

    @property
    def ets(self) -> List[torch.Tensor]:
        """History oldest-first, the reference's attribute name (scheduler_ppo.py:123)."""
        return self._hist[::-1]

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device] = None):
        """scheduler_ppo.py:142-163."""
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(f"`num_inference_steps` ({num_inference_steps}) cannot be larger than "
                             f"`num_train_timesteps` ({T}).")
        n = self.num_inference_steps = num_inference_steps
        sp = self.config.timestep_spacing
        if sp == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        elif sp == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        elif sp == "trailing":
            ts = np.round(np.arange(T, 0, -(T / n))).astype(np.int64) - 1
        else:
            raise ValueError(f"Unsupported timestep_spacing: {sp}.")
        self._timesteps_host = ts
        self._stride = T // n                       # prev_t = t - T//n (scheduler_ppo.py:203)
        self.timesteps = torch.from_numpy(ts).to(device)
        self._hist = []
        self._traj = None
        self._step_count = 0

    def scale_model_input(self, sample: torch.Tensor, timestep: Optional[int] = None) -> torch.Tensor:
        return sample

    # ------------------------------------------------------------------------------------------------------
    def _host_timestep(self, timestep) -> int:
        if isinstance(timestep, torch.Tensor):
            if timestep.is_cuda:
                if self.sync_free:
                    return int(self._timesteps_host[self._step_count % len(self._timesteps_host)])
                return int(timestep.item())
            return int(timestep)
        return int(timestep)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True):
        """Same contract as scheduler_ppo.py:178-299.  `return_dict=False` ->
        (prev_sample, actions [B,A], probs [B,A], conds {'x','epsilon'}, masks [B,A])."""
        return self._step(model_output, None, 0.0, timestep, sample, return_dict, None)

    def step_cfg(self, noise_pred: torch.Tensor, timestep, sample: torch.Tensor, guidance_scale: float,
                 return_dict: bool = False, out: Optional[torch.Tensor] = None,
                 out2: Optional[torch.Tensor] = None):
        """Fused variant: `noise_pred` is the denoiser output for torch.cat([latents]*2) — unconditional half
        first (denoise_ppo.py:66,:97) — and `u + g*(c-u)` is formed inside the step kernel, which also writes
        it into this step's slot of the scheduler-owned history ring.  `out` / `out2`: optional destinations for the
        next latent — e.g. the two halves of the next CFG-doubled denoiser input, which removes the caller's
        torch.cat([latents] * 2) (denoise_ppo.py:66)."""
        B = sample.shape[0]
        if noise_pred.shape[0] != 2 * B:
            raise ValueError("step_cfg expects the [2B, ...] classifier-free-guidance pair")
        if not noise_pred.is_contiguous():
            noise_pred = noise_pred.contiguous()
        if out2 is not None and (out2.shape != sample.shape or out2.dtype != sample.dtype or
                                 not out2[0].is_contiguous()):
            raise ValueError("out2 must have the sample's shape/dtype with contiguous samples")
        return self._step(noise_pred[:B], noise_pred[B:], float(guidance_scale), timestep, sample, return_dict, out,
                          out2)

    def _step(self, e0, cond, guidance, timestep, sample, return_dict, out, out2=None):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None'. Call 'set_timesteps' first.")
        if not (e0.is_cuda and sample.is_cuda):
            raise RuntimeError("consolver_b200 has no CPU path: model_output and sample must be CUDA tensors")
        if sample.dtype != e0.dtype:
            raise TypeError(f"sample ({sample.dtype}) and model_output ({e0.dtype}) must share a dtype")
        cfg = self.config
        fn = self.factor_net_module
        od = cfg.order_dim
        t = self._host_timestep(timestep)
        prev_t = t - self._stride
        B = sample.shape[0]
        N = sample.numel() // B
        e0 = e0 if e0.is_contiguous() else e0.contiguous()
        sample = sample if sample.is_contiguous() else sample.contiguous()
        tr = self._traj
        if tr is None or tr.key != (B, tuple(sample.shape[1:]), e0.dtype, e0.device):
            tr = self._traj = _Trajectory(self, B, sample.shape[1:], e0.dtype, e0.device)
        i = tr.count % tr.n
        depth = od if self.fixed_coefficients is None else min(od, getattr(self, "fixed_depth", od) or od)
        older = self._hist[: depth - 1]
        n_hist = len(older) + 1

        # policy input row (t, prev_t) rounded through the model dtype (scheduler_ppo.py:207)
        if t == self._timesteps_host[i]:
            x0, x1 = float(tr.condx_host[i, 0]), float(tr.condx_host[i, 1])
            conds_x = tr.condx[i:i + 1].expand(B, 2)
        else:
            row = torch.tensor([[t, prev_t]], dtype=e0.dtype)
            x0, x1 = (float(v) for v in row.float()[0])
            conds_x = row.to(e0.device).expand(B, 2)

        sa_t, sb_t = float(self._sqrt_abar[t]), float(self._sqrt_1m_abar[t])
        pi = prev_t if prev_t >= 0 else 0            # final step uses alphas_cumprod[0] (scheduler_ppo.py:114,:310)
        sa_p, sb_p = float(self._sqrt_abar[pi]), float(self._sqrt_1m_abar[pi])

        o = tr.out
        q_ptr, idx_ptr, rng_arg = tr.q.data_ptr(), None, None
        on_grid = t == self._timesteps_host[i]
        if self.fixed_coefficients is not None:
            pass                                     # baseline solvers draw nothing
        elif self.replay is None:
            # the draw torch.multinomial makes (factor_net_ppo.py:161): generated inside the sample kernel from the
            # default generator's (seed, offset) when possible, else by the torch launch into tr.q
            r = next_rng(self, tr, e0.device) if (on_grid and not fn.use_conv) else None
            if r is None:
                tr.q.exponential_(1)
            else:
                q_ptr, rng_arg = None, ctypes.byref(r)
        elif self.replay.get("idx") is not None:
            forced = self.replay["idx"][tr.count].to(device=e0.device, dtype=torch.int64).contiguous()
            q_ptr, idx_ptr = None, forced.data_ptr()
        else:
            tr.q.copy_(self.replay["q"][tr.count].reshape(tr.q.shape))
        x_out = out if out is not None else torch.empty_like(sample)
        slot = tr.slot(tr.count) if cond is not None else None
        flags = (_lib.FLAG_VPRED if cfg.prediction_type == "v_prediction" else 0) | \
                (_lib.FLAG_PDL if self.use_pdl else 0)
        hist_ptrs = _lib.ptr_array([h.data_ptr() for h in older])
        w = fn.kernel_weights() if self.fixed_coefficients is None else None
        lib = _lib.load()
        stream = torch._C._cuda_getCurrentRawStream(e0.device.index)
        step_args = (_lib.dtype_code(e0.dtype), e0.data_ptr(), cond.data_ptr() if cond is not None else None, guidance,
                     slot.data_ptr() if slot is not None else None, hist_ptrs, n_hist, sample.data_ptr(),
                     x_out.data_ptr(), out2.data_ptr() if out2 is not None else None,
                     out2.stride(0) if out2 is not None else 0)
        if self.fixed_coefficients is not None:
            # baseline solvers: same fused kernel, coefficients from a fixed table instead of the
            # policy — DDIM is depth 1, Adams-Bashforth / iPNDM style multistep are depths 2..4
            coef_t = self._fixed_coef_rows(tr, n_hist, B, od)
            rc = lib.consolver_step_sd(*step_args, coef_t.data_ptr(), od + 2, od, sa_t, sb_t, sa_p, sb_p,
                                       flags & ~_lib.FLAG_PDL, B, N, stream)
            _lib.check(rc, "consolver_step_sd")
        elif not fn.use_conv:
            # The policy input row depends only on the timestep grid: evaluate the MLP + softmax for ALL n rows in
            # one launch at the first step of a pass; every step then only samples from its row of the table.
            if on_grid and tr.table_pass != tr.count // tr.n:
                fn.policy_tables(tr.condx_f32, o["probs_table"])
                tr.table_pass = tr.count // tr.n
                tr.policy_forked = False                          # the side stream must see the new tables
            probs_in = tr.p("probs_table", i) if on_grid else None
            ps = self.policy_stream
            if ps is not None and on_grid and rng_arg is not None:
                # Two-stream form: the sample kernel needs nothing from the step kernels (only the table and the
                # generator state), so it runs on its own stream and the step kernel just waits for its event.
                # Captured in a CUDA graph this makes the sample chain a parallel branch: every step's coefficients
                # are ready before its step kernel starts, and the critical path is the step kernels alone.
                main = torch.cuda.current_stream(e0.device)
                if not tr.policy_forked:
                    ps.wait_stream(main)                          # fork (also orders after the table launch above)
                    tr.policy_forked = True
                rc = lib.consolver_policy_sample_f32(
                    probs_in, w[6], None, None, rng_arg, None, B, fn.action_dims, fn.num_actions, od, cfg.scaler_dim,
                    n_hist, tr.p("idx", i), tr.p("actions", i), tr.p("probs", i),
                    tr.p("logp", i), tr.p("masks", i), tr.p("coef", i), ps.cuda_stream)
                _lib.check(rc, "consolver_policy_sample_f32")
                ev = torch.cuda.Event()
                ev.record(ps)
                main.wait_event(ev)
                sflags = (flags & ~_lib.FLAG_PDL) | (_lib.FLAG_EFF_SCALE if cfg.scaler_dim >= 1 else 0) | \
                    (_lib.FLAG_X_SCALE if cfg.scaler_dim >= 2 else 0) | (_lib.FLAG_CHAIN if self.chain_steps else 0)
                rc = lib.consolver_step_sd(*step_args, tr.p("coef", i), od + 2, od, sa_t, sb_t, sa_p, sb_p,
                                           sflags, B, N, stream)
                _lib.check(rc, "consolver_step_sd")
            else:
                rc = lib.consolver_sd_policy_and_step(
                    *w, probs_in, x0, x1, fn.x_div, fn.temperature, q_ptr, idx_ptr, rng_arg,
                    fn.hidden_dim, fn.action_dims, fn.num_actions, cfg.scaler_dim,
                    tr.p("probs_table", i), tr.p("idx", i), tr.p("actions", i),
                    tr.p("probs", i), tr.p("logp", i), tr.p("masks", i),
                    tr.p("coef", i), *step_args, od, sa_t, sb_t, sa_p, sb_p, flags, B, N, stream)
                _lib.check(rc, "consolver_sd_policy_and_step")
        else:
            # use_conv=True (factor_net_ppo.py:146-149): two passes.  Pass 1 reduces the cosine features of the
            # history against the newest output (formed from the CFG pair on the fly); the policy MLP then runs per
            # sample on [t, t_prev, cos_1..]; pass 2 is the ordinary fused step.
            from .features import cosine_features_cuda

            feat, ws, full = tr.conv_buffers(fn)
            cosine_features_cuda(e0, cond, guidance, older, od, feat, ws, stream)
            rc = lib.consolver_policy_f32(
                *w, x0, x1, fn.x_div, fn.temperature, feat.data_ptr(), od - 1, q_ptr, idx_ptr,
                B, fn.hidden_dim, fn.action_dims, fn.num_actions, od, cfg.scaler_dim, n_hist,
                full[i].data_ptr(), tr.p("idx", i), tr.p("actions", i), tr.p("probs", i),
                tr.p("logp", i), tr.p("masks", i), tr.p("coef", i), stream)
            _lib.check(rc, "consolver_policy_f32")
            sflags = flags | (_lib.FLAG_EFF_SCALE if cfg.scaler_dim >= 1 else 0) | \
                (_lib.FLAG_X_SCALE if cfg.scaler_dim >= 2 else 0)
            rc = lib.consolver_step_sd(*step_args, tr.p("coef", i), od + 2, od, sa_t, sb_t, sa_p, sb_p,
                                       sflags, B, N, stream)
            _lib.check(rc, "consolver_step_sd")

        newest = slot if cond is not None else e0     # plain step keeps the caller's tensor by reference, as
        self._hist = [newest] + older                 # the reference does (scheduler_ppo.py:214-218)
        tr.count += 1
        self._step_count += 1

        hist_now = list(self._hist)
        shape = tuple(sample.shape[1:])

        def _stack():
            s = torch.stack(hist_now, dim=1)
            if len(hist_now) < od:
                s = torch.cat([s, s.new_zeros(B, od - len(hist_now), *shape)], dim=1)
            return s

        actions, probs, masks = o["actions"][i], o["probs"][i], o["masks"][i]
        if self.fixed_coefficients is not None:
            actions = probs = masks = None
        conds = LazyConds(conds_x, _stack)
        if not return_dict:
            return (x_out, actions, probs, conds, masks)
        return PPOSchedulerOutput(prev_sample=x_out, actions=actions, probs=probs, conds=conds, masks=masks)

    def _fixed_coef_rows(self, tr, n_hist, B, od):
        cache = tr.__dict__.setdefault("_fixed", {})
        if n_hist not in cache:
            c = [float(v) for v in self.fixed_coefficients(n_hist)]
            if len(c) != n_hist:
                raise ValueError("fixed_coefficients(n_hist) must return n_hist values")
            row = torch.tensor(c + [0.0] * (od - n_hist) + [1.0, 1.0], dtype=torch.float32)
            cache[n_hist] = row.to(tr.device).expand(B, od + 2).contiguous()
        return cache[n_hist]

    # ------------------------------------------------------------------------------------------------------
    def trajectory(self, skip_first: bool = True):
        """The rollout record denoise_ppo.py:105-118 assembles with unsqueeze+cat, as views of the
        per-trajectory buffers: dict(x [B,n',2], probs/actions/masks/idx [B,n',A]) for steps 1..count-1."""
        tr = self._traj
        if tr is None:
            raise ValueError("no trajectory recorded; call step() first")
        lo, hi = (1 if skip_first else 0), min(tr.count, tr.n)
        B = tr.key[0]
        pick = lambda k: tr.out[k][lo:hi].transpose(0, 1)  # noqa: E731
        return dict(x=tr.condx[lo:hi].unsqueeze(0).expand(B, hi - lo, 2), probs=pick("probs"),
                    actions=pick("actions"), masks=pick("masks"), idx=pick("idx"), logp=pick("logp"))

    def last_policy(self):
        """Full softmax table [A,K], sampled indices [B,A] and coefficient records [B,order_dim+2] of the most
        recent step (views)."""
        tr = self._traj
        i = (tr.count - 1) % tr.n
        table = tr._conv[2][i] if getattr(tr, "_conv", None) is not None else tr.out["probs_table"][i]
        return dict(probs_table=table, idx=tr.out["idx"][i], coef=tr.out["coef"][i], logp=tr.out["logp"][i])

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """DDPM forward noising (scheduler_ppo.py:336-358); not on the hot path, plain torch."""
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = (ac[timesteps] ** 0.5).flatten()
        b = ((1 - ac[timesteps]) ** 0.5).flatten()
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise

    def __len__(self):
        return self.config.num_train_timesteps
