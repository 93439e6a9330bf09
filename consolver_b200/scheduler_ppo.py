"""PPOScheduler — drop-in for the reference's `scheduler_ppo.PPOScheduler` (scheduler_ppo.py:48-361): same
constructor kwargs, `set_timesteps()` / `step()` signatures and 5-tuple return, same `factor_net` state_dict —
with the per-step work done by hand-written sm_100a kernels behind the C ABI (include/consolver.h):

  table kernel   policy MLP + softmax for every row of the timestep grid, once per trajectory
  sample kernel  per-sample categorical draw (torch's own RNG stream, regenerated in the kernel) + coefficient/mask
                 assembly, once per step
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

import dataclasses
import math
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from . import _lib
from ._sched_common import SolverOptions, Trajectory, draw_source, lazy_conds, next_rng  # noqa: F401 (next_rng re-export)
from .config_utils import KARRAS_COMPATIBLES, BaseOutput, ConfigMixin, SchedulerMixin, register_to_config
from .factor_net import FactorNetPPO, FactorNetPPOContinous


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


class PPOScheduler(SolverOptions, SchedulerMixin, ConfigMixin):
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
            # scheduler_ppo.py:139 instantiates FactorNetPPOContinous, whose source is not part of the reference: this
            # is an EXTENSION with self-defined semantics (factor_net.FactorNetPPOContinous), parity unpinned
            self.factor_net = FactorNetPPOContinous(**kw)
        else:
            kw.setdefault("num_actions", 161)
            self.factor_net = FactorNetPPO(**kw)

        self._init_solver_options()
        self._step_count = 0
        self._grid_offset = None
        self._stride = 0

    # ------------------------------------------------------------------------------------------------------
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
        self._grid_offset = None

    def scale_model_input(self, sample: torch.Tensor, timestep: Optional[int] = None) -> torch.Tensor:
        return sample

    # ------------------------------------------------------------------------------------------------------
    def _host_timestep(self, timestep) -> int:
        """Value of `timestep`.  Host ints / CPU tensors are read directly.  A 0-d VIEW of `scheduler.timesteps` — what
        `for t in scheduler.timesteps` and `timesteps[t_start:]` hand out — is located by its storage offset: no device
        read-back at all.  Every other CUDA tensor is read back (`.item()`), as the reference does.
        `sync_free = False` disables the view shortcut too."""
        if not isinstance(timestep, torch.Tensor):
            return int(timestep)
        if not timestep.is_cuda:
            return int(timestep)
        if not self.sync_free:
            return int(timestep.item())
        grid = self._timesteps_host
        # `for t in scheduler.timesteps` (and `timesteps[t_start:]`) hands out 0-d VIEWS of our own grid tensor: the
        # position is the storage offset, no read-back at all
        ts = self.timesteps
        if (timestep.dim() == 0 and ts.is_cuda and timestep.dtype == ts.dtype and
                timestep.untyped_storage().data_ptr() == ts.untyped_storage().data_ptr()):
            j = timestep.storage_offset() - ts.storage_offset()
            if 0 <= j < len(grid):
                return int(grid[j])
        # Any other CUDA tensor (a clone, a value computed by the caller, a custom subset) is READ BACK, every step, exactly
        # like the reference does (scheduler_ppo.py:205): the value that was passed is honoured, whatever order the caller
        # steps in.  Only under stream capture — where a read-back is impossible — the grid order is assumed.
        if torch.cuda.is_current_stream_capturing():
            j = self._step_count
            if 0 <= j < len(grid):
                return int(grid[j])
            raise RuntimeError("cannot read a timestep tensor back during CUDA-graph capture; pass scheduler.timesteps[i] "
                               "(a view of the scheduler's own grid) or a python int")
        return int(timestep.item())

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
        if out2 is not None and (out2.shape != sample.shape or not out2[0].is_contiguous()):
            raise ValueError("out2 must have the sample's shape with contiguous samples")
        if out is not None and (out.shape != sample.shape or not out.is_contiguous()):
            raise ValueError("out must have the sample's shape and be contiguous (it is written through its data pointer)")
        if noise_pred.device != sample.device:
            raise ValueError("noise_pred and sample live on different devices")
        return self._step(noise_pred[:B], noise_pred[B:], float(guidance_scale), timestep, sample, return_dict, out,
                          out2)

    def next_latent_dtype(self, model_dtype: torch.dtype, sample_dtype: torch.dtype) -> torch.dtype:
        """dtype of the latent the next `step` returns — what torch promotion makes of the reference's arithmetic
        (see the note in `_step`): fp32, except for an all-16-bit step whose estimate is still the raw model output."""
        if model_dtype in (torch.float16, torch.bfloat16) and sample_dtype == model_dtype:
            n_hist = len(self._hist[: self._history_depth(self.config.order_dim) - 1]) + 1
            if self._estimate_stays_lowp(model_dtype, n_hist):
                return model_dtype
        return torch.float32

    def _estimate_stays_lowp(self, model_dtype: torch.dtype, n_hist: int) -> bool:
        """Does the combined estimate keep the 16-bit model dtype at this step?  In the reference the per-sample
        coefficients and scalers are gathered from the policy's `action_values` buffer (factor_net_ppo.py:163) and
        carry ITS dtype: an fp32 policy promotes the estimate at every step that multiplies by one (all but a
        scaler-free first step).  A policy cast to the pipeline's own 16-bit dtype (gen_ppo.py:193-195) keeps a depth-1
        estimate 16-bit through its 16-bit scalers, but from depth 2 on the closing coefficient `1 - torch.sum(...)`
        (scheduler_ppo.py:172) is fp32 — torch.sum returns fp32 under autocast (gen_ppo.py:309) — and promotes it
        (tests/golden/cuda_genppo_*: fp16 latents only after the first step)."""
        if self.factor_net_module.action_values.dtype == model_dtype:
            return n_hist == 1
        return n_hist == 1 and self.config.scaler_dim == 0

    def _new_trajectory(self, B, shape, dtype, device) -> Trajectory:
        rows = [[float(t), float(t - self._stride)] for t in self._timesteps_host]      # (t, prev_t): scheduler_ppo.py:203-207
        return Trajectory(self.factor_net_module, self.num_inference_steps, self.config.order_dim, B, shape, dtype,
                          device, rows)

    def _step(self, e0, cond, guidance, timestep, sample, return_dict, out, out2=None):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None'. Call 'set_timesteps' first.")
        if not (e0.is_cuda and sample.is_cuda):
            raise RuntimeError("consolver_b200 has no CPU path: model_output and sample must be CUDA tensors")
        if e0.device.index != torch.cuda.current_device():
            # one process driving several GPUs: launch in the tensors' own device context (kernel attributes, SM count
            # and the stream handle all belong to that device)
            with torch.cuda.device(e0.device):
                return self._step(e0, cond, guidance, timestep, sample, return_dict, out, out2)
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
            tr = self._traj = self._new_trajectory(B, sample.shape[1:], e0.dtype, e0.device)
        i = tr.count % tr.n
        older = self._hist[: self._history_depth(od) - 1]
        n_hist = len(older) + 1
        fixed = self.fixed_coefficients is not None

        # Latent dtype, resolved the way torch promotion resolves it in the reference (scheduler_ppo.py:263-280,
        # :316-330).  With a 16-bit model output the combined estimate turns fp32 as soon as an fp32 per-sample
        # coefficient or scaler multiplies it (every step but a scaler-free first one), and the latent with it: an
        # fp16 pipeline (gen_ppo.py) gets fp32 latents back from its second step on, an autocast rollout
        # (train_ppo.py:353: fp32 latents, 16-bit U-Net output) keeps fp32 throughout.  The kernel then reads / writes x
        # as fp32 (CONSOLVER_FLAG_X_F32) with the model outputs and history still 16-bit.
        mixed = 0
        lowp = e0.dtype in (torch.float16, torch.bfloat16)
        if sample.dtype != e0.dtype and not (lowp and sample.dtype == torch.float32) and not (
                e0.dtype == torch.float32 and sample.dtype in (torch.float16, torch.bfloat16)):
            raise TypeError(f"sample ({sample.dtype}) and model_output ({e0.dtype}) cannot be combined")
        if lowp:
            raw_estimate = self._estimate_stays_lowp(e0.dtype, n_hist)
            if sample.dtype != torch.float32 and not raw_estimate:
                sample = sample.float()                   # exact; the reference's sample is still 16-bit at this step,
                mixed = _lib.FLAG_X_WAS_LOWP              # which only v-prediction's scalar*sample product can see
            if sample.dtype == torch.float32:
                mixed |= _lib.FLAG_X_F32
        elif sample.dtype != torch.float32:
            sample = sample.float()
        for name, dst in (("out", out), ("out2", out2)):
            if dst is not None and dst.dtype != sample.dtype:
                raise ValueError(f"{name} is {dst.dtype} but this step returns a {sample.dtype} latent "
                                 "(16-bit model outputs promote the latent to fp32 from the second step on)")

        # policy input row (t, prev_t) rounded through the model dtype (scheduler_ppo.py:207).  gi = position of t in
        # the grid (normally the step count; elsewhere when the caller starts mid-grid), None when t is off the grid
        grid = self._timesteps_host
        gi = i if grid[i] == t else next(iter(np.nonzero(grid == t)[0].tolist()), None)
        on_grid = gi is not None
        if on_grid:
            x0, x1 = float(tr.condx_host[gi, 0]), float(tr.condx_host[gi, 1])
            conds_x = tr.condx[gi:gi + 1].expand(B, 2)
        else:
            row = torch.tensor([[t, prev_t]], dtype=e0.dtype)
            x0, x1 = (float(v) for v in row.float()[0])
            conds_x = row.to(e0.device).expand(B, 2)

        sa_t, sb_t = float(self._sqrt_abar[t]), float(self._sqrt_1m_abar[t])
        pi = prev_t if prev_t >= 0 else 0            # final step uses alphas_cumprod[0] (scheduler_ppo.py:114,:310)
        sa_p, sb_p = float(self._sqrt_abar[pi]), float(self._sqrt_1m_abar[pi])

        if getattr(fn, "continuous", False):
            q_ptr = idx_ptr = rng_arg = None              # the Gaussian policy draws torch.randn, see _gauss_draw
        else:
            q_ptr, idx_ptr, rng_arg = draw_source(self, tr, e0.device, fused_ok=on_grid and not fn.use_conv)
        x_out = out if out is not None else torch.empty_like(sample)
        slot = tr.slot(tr.count) if cond is not None else None
        sem_flags, pflags, act_dt = self._semantics(e0.dtype)
        vflag = (_lib.FLAG_VPRED if cfg.prediction_type == "v_prediction" else 0) | mixed | sem_flags
        sflag = (_lib.FLAG_EFF_SCALE if cfg.scaler_dim >= 1 else 0) | (_lib.FLAG_X_SCALE if cfg.scaler_dim >= 2 else 0)
        pdl = _lib.FLAG_PDL if self.use_pdl else 0
        lib = _lib.load()
        stream = torch._C._cuda_getCurrentRawStream(e0.device.index)
        step_args = (_lib.dtype_code(e0.dtype), e0.data_ptr(), cond.data_ptr() if cond is not None else None, guidance,
                     slot.data_ptr() if slot is not None else None, _lib.ptr_array([h.data_ptr() for h in older]),
                     n_hist, sample.data_ptr(), x_out.data_ptr(), out2.data_ptr() if out2 is not None else None,
                     out2.stride(0) if out2 is not None else 0)
        tail = (od, sa_t, sb_t, sa_p, sb_p)
        outs = tuple(tr.p(k, i) for k in ("idx", "actions", "probs", "logp", "masks", "coef"))

        if fixed:
            # baseline solvers: same fused kernel, coefficients from a fixed table instead of the
            # policy — DDIM is depth 1, Adams-Bashforth / iPNDM style multistep are depths 2..4
            rc = lib.consolver_step_sd(*step_args, tr.fixed_rows(self.fixed_coefficients, n_hist).data_ptr(), od + 2,
                                       *tail, vflag, B, N, stream)
            _lib.check(rc, "consolver_step_sd")
        elif getattr(fn, "continuous", False):
            # Gaussian policy (extension, parity unpinned): one fused kernel — MLP head, the torch.randn draw regenerated
            # in the kernel, log-prob, masks, coefficient assembly — then the ordinary fused step as its PDL dependent
            z_ptr, act_ptr, g_rng = self._gauss_draw(tr, e0.device, B, fn.action_dims)
            rc = lib.consolver_policy_gauss_f32(
                *fn.kernel_weights(), x0, x1, fn.x_div, z_ptr, act_ptr, g_rng, B, fn.hidden_dim, fn.action_dims, od,
                cfg.scaler_dim, n_hist, pflags & _lib.POLICY_HOST_DIV, tr.p("probs_table", gi if on_grid else tr.n),
                None, *outs[1:], stream)
            _lib.check(rc, "consolver_policy_gauss_f32")
            rc = lib.consolver_step_sd(*step_args, outs[5], od + 2, *tail, vflag | sflag | pdl, B, N, stream)
            _lib.check(rc, "consolver_step_sd")
        elif fn.use_conv:
            # use_conv=True (factor_net_ppo.py:146-149): two passes.  Pass 1 reduces the cosine features of the
            # history against the newest output (formed from the CFG pair on the fly); the policy MLP then runs per
            # sample on [t, t_prev, cos_1..]; pass 2 is the ordinary fused step.
            from .features import cosine_features_cuda

            feat, ws, full = tr.conv_buffers(fn)
            cosine_features_cuda(e0, cond, guidance, older, od, feat, ws, stream)
            rc = lib.consolver_policy_f32(
                *fn.kernel_weights(act_dt), x0, x1, fn.x_div, fn.temperature, feat.data_ptr(), od - 1, q_ptr, idx_ptr,
                B, fn.hidden_dim, fn.action_dims, fn.num_actions, od, cfg.scaler_dim, n_hist, pflags,
                full[i].data_ptr(), *outs, stream)
            _lib.check(rc, "consolver_policy_f32")
            rc = lib.consolver_step_sd(*step_args, outs[5], od + 2, *tail, vflag | sflag | pdl, B, N, stream)
            _lib.check(rc, "consolver_step_sd")
        else:
            w = fn.kernel_weights(act_dt)
            # The policy input row depends only on the timestep grid: evaluate the MLP + softmax for ALL n rows in
            # one launch at the first step of a pass; every step then only samples from its row of the table.
            if on_grid and tr.table_pass != tr.count // tr.n:
                fn.policy_tables(tr.condx_f32, tr.out["probs_table"], policy_flags=pflags, act_dtype=act_dt)
                tr.table_pass = tr.count // tr.n
                tr.policy_forked = False                          # the side stream must see the new tables
            probs_in = tr.p("probs_table", gi) if on_grid else None
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
                rc = lib.consolver_policy_sample_f32(probs_in, w[6], None, None, rng_arg, None, B, fn.action_dims,
                                                     fn.num_actions, od, cfg.scaler_dim, n_hist, pflags, *outs,
                                                     ps.cuda_stream)
                _lib.check(rc, "consolver_policy_sample_f32")
                ev = torch.cuda.Event()
                ev.record(ps)
                main.wait_event(ev)
                # no policy-PDL here: the previous node on this stream is a step kernel (see CONSOLVER_FLAG_CHAIN)
                chain = _lib.FLAG_CHAIN if self.chain_steps else 0
                rc = lib.consolver_step_sd(*step_args, outs[5], od + 2, *tail, vflag | sflag | chain, B, N, stream)
                _lib.check(rc, "consolver_step_sd")
            else:
                rc = lib.consolver_sd_policy_and_step(
                    *w, probs_in, x0, x1, fn.x_div, fn.temperature, q_ptr, idx_ptr, rng_arg,
                    fn.hidden_dim, fn.action_dims, fn.num_actions, cfg.scaler_dim, pflags,
                    tr.p("probs_table", gi if on_grid else tr.n), *outs, *step_args, *tail, vflag | pdl, B, N, stream)
                _lib.check(rc, "consolver_sd_policy_and_step")

        tr.last_table_row = gi if on_grid else tr.n
        newest = slot if cond is not None else e0     # plain step keeps the caller's tensor by reference, as
        self._hist = [newest] + older                 # the reference does (scheduler_ppo.py:214-218)
        tr.grid_rows.append(gi)
        tr.count += 1
        self._step_count += 1

        o = tr.out
        actions, probs, masks = (None, None, None) if fixed else (o["actions"][i], o["probs"][i], o["masks"][i])
        if actions is not None and fn.action_values.dtype != torch.float32:
            actions = actions.to(fn.action_values.dtype)      # exact: a policy cast to 16 bit returns 16-bit bin values
        conds = lazy_conds(conds_x, list(self._hist), od, tr, ring=cond is not None)
        if not return_dict:
            return (x_out, actions, probs, conds, masks)
        return PPOSchedulerOutput(prev_sample=x_out, actions=actions, probs=probs, conds=conds, masks=masks)

    def _gauss_draw(self, tr, device, B, A):
        """draw source of the continuous policy -> (z_ptr, actions_ptr, rng_arg), exactly one non-null: replay['z'] /
        replay['actions'] per step, else the in-kernel regeneration of torch.randn([B,A]) (or that launch itself)"""
        import ctypes

        from . import rng as _rng

        rp = self.replay
        if rp is not None and rp.get("actions") is not None:
            forced = rp["actions"][tr.count].to(device=device, dtype=torch.float32).contiguous()
            tr._forced_keepalive = forced
            return None, forced.data_ptr(), None
        if rp is not None and rp.get("z") is not None:
            z = rp["z"][tr.count].to(device=device, dtype=torch.float32).contiguous()
            tr._forced_keepalive = z
            return z.data_ptr(), None, None
        if self.use_fused_rng and not torch.cuda.is_current_stream_capturing() and _rng.fused_rng_available(device):
            nthreads, inc = _lib.philox_plan(B * A)
            seed, off = _rng.take(device, inc)
            r = _lib.Rng(seed, off, None, nthreads)
            tr._rng_keepalive = r
            return None, None, ctypes.byref(r)
        z = torch.randn(B, A, device=device)
        tr._forced_keepalive = z
        return z.data_ptr(), None, None

    # ------------------------------------------------------------------------------------------------------
    def last_policy(self):
        """Full softmax table [A,K] ([B,A,K] with use_conv), sampled indices [B,A], coefficient records
        [B,order_dim+2] and log-probs of the most recent step (views)."""
        lp = self._traj.last()
        fn = self.factor_net_module
        if getattr(fn, "continuous", False):       # the table row holds {mean[A], std[A]} for the Gaussian policy
            ms = lp.pop("probs_table").reshape(2, fn.action_dims)
            lp.update(mean=ms[0], std=ms[1])
            lp.pop("idx", None)
        return lp

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
