"""FMPPOScheduler — drop-in for the reference's flow-matching solver `edit_ppo/scheduler_fmppo.FMPPOScheduler`
(edit_ppo/scheduler_fmppo.py:56-553): same constructor kwargs, `set_timesteps(num_inference_steps, device,
sigmas, mu, timesteps)`, `set_begin_index`, `step(...)` signature (incl. the ignored s_churn/... arguments),
`scale_noise`, and `factor_net` state_dict.  The step runs the policy kernel and the fused Euler-form step kernel
(bf16/fp16/fp32 I/O, fp32 math) through the C ABI; see scheduler_ppo.py in this package for the design notes.
`per_token_timesteps` (edit_ppo/scheduler_fmppo.py:363-371) is exercised by no caller and is not supported."""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from . import _lib
from ._sched_common import SolverOptions, Trajectory, draw_source, lazy_conds
from .config_utils import BaseOutput, ConfigMixin, SchedulerMixin, register_to_config
from .factor_net import FactorNetPPOFM


@dataclasses.dataclass
class FMPPOSchedulerOutput(BaseOutput):
    """edit_ppo/scheduler_fmppo.py:33-54."""
    prev_sample: torch.Tensor = None
    actions: Optional[torch.Tensor] = None
    probs: Optional[torch.Tensor] = None
    conds: Optional[Dict] = None
    masks: Optional[torch.Tensor] = None


class FMPPOScheduler(SolverOptions, SchedulerMixin, ConfigMixin):
    """Learned linear-multistep Euler-form solver (ConsistencySolver) for flow-matching models."""

    _compatibles = []
    order = 1

    @register_to_config
    def __init__(
        self,
        num_train_timesteps: int = 1000,
        shift: float = 1.0,
        use_dynamic_shifting: bool = False,
        base_shift: Optional[float] = 0.5,
        max_shift: Optional[float] = 1.15,
        base_image_seq_len: Optional[int] = 256,
        max_image_seq_len: Optional[int] = 4096,
        invert_sigmas: bool = False,
        shift_terminal: Optional[float] = None,
        use_karras_sigmas: Optional[bool] = False,
        use_exponential_sigmas: Optional[bool] = False,
        use_beta_sigmas: Optional[bool] = False,
        time_shift_type: str = "exponential",
        stochastic_sampling: bool = False,
        order_dim: int = 4,
        scaler_dim: int = 2,
        mu_dim: int = 1,
        use_conv: bool = False,
        ppo_type: str = "discrete",
        factor_net_kwargs: Optional[Dict] = None,
    ):
        if use_beta_sigmas:
            try:
                import scipy.stats  # noqa: F401
            except ImportError as e:  # edit_ppo/scheduler_fmppo.py:132-133
                raise ImportError("Make sure to install scipy if you want to use beta sigmas.") from e
        if sum([bool(use_beta_sigmas), bool(use_exponential_sigmas), bool(use_karras_sigmas)]) > 1:
            raise ValueError("Only one of `use_beta_sigmas`, `use_exponential_sigmas`, `use_karras_sigmas` can be used.")
        if time_shift_type not in {"exponential", "linear"}:
            raise ValueError("`time_shift_type` must either be 'exponential' or 'linear'.")
        if order_dim < 2 or order_dim > _lib.MAX_ORDER:
            raise ValueError(f"order_dim must be in [2, {_lib.MAX_ORDER}]")
        if scaler_dim not in (0, 1, 2):
            raise NotImplementedError("More than two scale parameters not supported.")

        # default 1000-point grid (edit_ppo/scheduler_fmppo.py:142-151)
        ts = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        sig = torch.from_numpy(ts).to(torch.float32) / num_train_timesteps
        if not use_dynamic_shifting:
            sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = sig * num_train_timesteps
        self.sigmas = sig.to("cpu")
        self._sigmas_host = self.sigmas.numpy()
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self._shift = shift
        self._step_index = None
        self._begin_index = None
        self.num_inference_steps = None

        kw = dict(factor_net_kwargs) if factor_net_kwargs is not None else {}
        kw.update(order_dim=order_dim, scaler_dim=scaler_dim, mu_dim=mu_dim, use_conv=use_conv)
        kw.setdefault("embedding_dim", 32)
        kw.setdefault("hidden_dim", 256)
        if ppo_type != "discrete":
            raise NotImplementedError("ppo_type != 'discrete' is `assert 0` in the reference "
                                      "(edit_ppo/scheduler_fmppo.py:169-170)")
        kw.setdefault("num_actions", 161)
        self.factor_net = FactorNetPPOFM(**kw)
        self._init_solver_options()
        self._curr_sigma = None

    # ---- small properties / helpers of the reference surface ---------------------------------------------------
    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    @property
    def shift(self):
        return self._shift

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def set_shift(self, shift: float):
        self._shift = shift

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    def time_shift(self, mu: float, sigma: float, t):
        """edit_ppo/scheduler_fmppo.py:489-493,:546-550."""
        if self.config.time_shift_type == "exponential":
            return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)
        return mu / (mu + (1 / t - 1) ** sigma)

    def stretch_shift_to_terminal(self, t):
        one_minus = 1 - t
        return 1 - one_minus / (one_minus[-1] / (1 - self.config.shift_terminal))

    def _resample(self, sig, n, kind):
        """karras / exponential / beta re-spacings between the first and last sigma (:516-544)."""
        lo = self.config.sigma_min if hasattr(self.config, "sigma_min") else sig[-1].item()
        hi = self.config.sigma_max if hasattr(self.config, "sigma_max") else sig[0].item()
        if kind == "karras":
            rho, ramp = 7.0, np.linspace(0, 1, n)
            return (hi ** (1 / rho) + ramp * (lo ** (1 / rho) - hi ** (1 / rho))) ** rho
        if kind == "exponential":
            return np.exp(np.linspace(math.log(hi), math.log(lo), n))
        import scipy.stats

        return np.array([lo + scipy.stats.beta.ppf(u, 0.6, 0.6) * (hi - lo) for u in 1 - np.linspace(0, 1, n)])

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device: Union[str, torch.device] = None,
                      sigmas: Optional[List[float]] = None, mu: Optional[float] = None,
                      timesteps: Optional[List[float]] = None):
        """edit_ppo/scheduler_fmppo.py:171-245."""
        cfg = self.config
        if cfg.use_dynamic_shifting and mu is None:
            raise ValueError("`mu` must be passed when `use_dynamic_shifting` is set to be `True`")
        if sigmas is not None and timesteps is not None and len(sigmas) != len(timesteps):
            raise ValueError("`sigmas` and `timesteps` should have the same length")
        if num_inference_steps is not None:
            if (sigmas is not None and len(sigmas) != num_inference_steps) or (
                    timesteps is not None and len(timesteps) != num_inference_steps):
                raise ValueError("`sigmas` and `timesteps` should have the same length as num_inference_steps, "
                                 "if `num_inference_steps` is provided")
        else:
            num_inference_steps = len(sigmas) if sigmas is not None else len(timesteps)
        self.num_inference_steps = num_inference_steps
        given_ts = timesteps is not None
        if given_ts:
            timesteps = np.array(timesteps).astype(np.float32)
        if sigmas is None:
            if timesteps is None:
                timesteps = np.linspace(self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min),
                                        num_inference_steps)
            sig = timesteps / cfg.num_train_timesteps
        else:
            sig = np.array(sigmas).astype(np.float32)
            num_inference_steps = len(sig)
        if cfg.use_dynamic_shifting:
            sig = self.time_shift(mu, 1.0, sig)
        else:
            sig = self.shift * sig / (1 + (self.shift - 1) * sig)
        if cfg.shift_terminal:
            sig = self.stretch_shift_to_terminal(sig)
        if cfg.use_karras_sigmas:
            sig = self._resample(sig, num_inference_steps, "karras")
        elif cfg.use_exponential_sigmas:
            sig = self._resample(sig, num_inference_steps, "exponential")
        elif cfg.use_beta_sigmas:
            sig = self._resample(sig, num_inference_steps, "beta")
        sig_t = torch.from_numpy(np.asarray(sig)).to(dtype=torch.float32)      # host; moved to `device` below
        ts_t = torch.from_numpy(timesteps).to(dtype=torch.float32) if given_ts else sig_t * cfg.num_train_timesteps
        if cfg.invert_sigmas:
            sig_t = 1.0 - sig_t
            ts_t = sig_t * cfg.num_train_timesteps
            sig_t = torch.cat([sig_t, torch.ones(1)])
        else:
            sig_t = torch.cat([sig_t, torch.zeros(1)])
        self._sigmas_host = sig_t.numpy().copy()
        self._timesteps_host = ts_t.numpy().copy()
        self.timesteps = ts_t.to(device=device)
        self.sigmas = sig_t.to(device=device)
        self._step_index = None
        self._begin_index = None
        self._hist = []
        self._traj = None
        self._curr_sigma = None

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        """edit_ppo/scheduler_fmppo.py:501-506 on the host copy of the grid (one read-back if `timestep` lives on
        the GPU; pipelines avoid it with set_begin_index)."""
        if schedule_timesteps is None:
            grid = self._timesteps_host
        else:
            grid = schedule_timesteps.detach().float().cpu().numpy()
        tv = np.float32(timestep.item() if isinstance(timestep, torch.Tensor) else timestep)
        hits = np.nonzero(grid == tv)[0]
        return int(hits[1 if len(hits) > 1 else 0])

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            self._step_index = self.index_for_timestep(timestep)
        else:
            self._step_index = self._begin_index

    # ------------------------------------------------------------------------------------------------------
    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, s_churn: float = 0.0,
             s_tmin: float = 0.0, s_tmax: float = float("inf"), s_noise: float = 1.0,
             generator: Optional[torch.Generator] = None, per_token_timesteps: Optional[torch.Tensor] = None,
             return_dict: bool = True, out2: Optional[torch.Tensor] = None):
        """Same contract as edit_ppo/scheduler_fmppo.py:306-455 (s_churn/s_tmin/s_tmax/s_noise/generator are
        accepted and unused there as well).  `out2` (new, optional): a second destination for the next latent with
        contiguous samples and any sample stride, e.g. `latent_model_input[:, :L]` of the next transformer call,
        which removes the caller's torch.cat([latents, image_latents], dim=1) copy of the latents
        (edit_ppo/denoise_diffusion.py:102)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None'. Call 'set_timesteps' first.")
        if isinstance(timestep, int) or (isinstance(timestep, torch.Tensor) and
                                         timestep.dtype in (torch.int32, torch.int64)):
            raise ValueError("Passing integer indices as timesteps to `step()` is not supported. "
                             "Pass one of `scheduler.timesteps`.")
        if per_token_timesteps is not None:
            raise NotImplementedError("per_token_timesteps is not supported (no caller of the reference uses it)")
        if not (model_output.is_cuda and sample.is_cuda):
            raise RuntimeError("consolver_b200 has no CPU path: model_output and sample must be CUDA tensors")
        if self._step_index is None:
            self._init_step_index(timestep)
        cfg = self.config
        fn = self.factor_net_module
        od = cfg.order_dim
        e0 = model_output if model_output.is_contiguous() else model_output.contiguous()
        sample = sample if sample.is_contiguous() else sample.contiguous()
        if sample.dtype != e0.dtype:
            sample = sample.float()                     # the reference upcasts the sample anyway (:354)
        B = sample.shape[0]
        N = sample.numel() // B
        tr = self._traj
        if tr is None or tr.key != (B, tuple(sample.shape[1:]), e0.dtype, e0.device):
            sg = self._sigmas_host
            rows = [[float(sg[j]), float(sg[j + 1])] for j in range(len(sg) - 1)]      # (sigma, sigma_next): :383
            tr = self._traj = Trajectory(fn, self.num_inference_steps, od, B, sample.shape[1:], e0.dtype, e0.device, rows)
        si = self._step_index
        if si + 1 >= len(self._sigmas_host):
            raise IndexError("FMPPOScheduler.step called past the end of the sigma schedule")
        i = tr.count % tr.n
        older = self._hist[: self._history_depth(od) - 1]
        n_hist = len(older) + 1
        fixed = self.fixed_coefficients is not None
        dt = float(np.float32(self._sigmas_host[si + 1]) - np.float32(self._sigmas_host[si]))   # :373-376
        x0, x1 = float(tr.condx_host[si, 0]), float(tr.condx_host[si, 1])
        conds_x = tr.condx[si:si + 1].expand(B, 2)

        q_ptr, idx_ptr, rng_arg = draw_source(self, tr, e0.device, fused_ok=not fn.use_conv)
        x_out = torch.empty(sample.shape, device=e0.device, dtype=e0.dtype)
        lib = _lib.load()
        stream = torch._C._cuda_getCurrentRawStream(e0.device.index)
        outs = tuple(tr.p(k, i) for k in ("idx", "actions", "probs", "logp", "masks", "coef"))
        coef_ptr = outs[5]
        flags = (_lib.FLAG_EFF_SCALE if cfg.scaler_dim >= 1 else 0) | (_lib.FLAG_X_SCALE if cfg.scaler_dim >= 2 else 0)
        if fixed:
            coef_ptr = tr.fixed_rows(self.fixed_coefficients, n_hist).data_ptr()      # baseline solvers: no policy
        elif fn.use_conv:
            # use_conv=True: cosine features of the history (pass 1), per-sample MLP, then the fused step (pass 2)
            from .features import cosine_features_cuda

            feat, ws, full = tr.conv_buffers(fn)
            cosine_features_cuda(e0, None, 0.0, older, od, feat, ws, stream)
            rc = lib.consolver_policy_f32(
                *fn.kernel_weights(), x0, x1, fn.x_div, fn.temperature, feat.data_ptr(), od - 1, q_ptr, idx_ptr,
                B, fn.hidden_dim, fn.action_dims, fn.num_actions, od, cfg.scaler_dim, n_hist,
                full[i].data_ptr(), *outs, stream)
            _lib.check(rc, "consolver_policy_f32")
            flags |= _lib.FLAG_PDL if self.use_pdl else 0
        else:
            # all n (sigma, sigma_next) rows of the schedule go through the MLP in one launch per pass
            if tr.table_pass != tr.count // tr.n:
                fn.policy_tables(tr.condx_f32, tr.out["probs_table"], stream)
                tr.table_pass = tr.count // tr.n
                tr.policy_forked = False
            ps = self.policy_stream if rng_arg is not None else None      # two-stream form: see PPOScheduler._step
            if ps is not None:
                main = torch.cuda.current_stream(e0.device)
                if not tr.policy_forked:
                    ps.wait_stream(main)
                    tr.policy_forked = True
            rc = lib.consolver_policy_sample_f32(
                tr.p("probs_table", si), fn.kernel_weights()[6], q_ptr, idx_ptr, rng_arg, None, B, fn.action_dims,
                fn.num_actions, od, cfg.scaler_dim, n_hist, *outs, ps.cuda_stream if ps is not None else stream)
            _lib.check(rc, "consolver_policy_sample_f32")
            if ps is not None:
                ev = torch.cuda.Event()
                ev.record(ps)
                main.wait_event(ev)
                flags |= _lib.FLAG_CHAIN if self.chain_steps else 0   # previous node on this stream is a step kernel
            elif self.use_pdl:
                flags |= _lib.FLAG_PDL
        rc = lib.consolver_step_fm(
            _lib.dtype_code(e0.dtype), _lib.dtype_code(sample.dtype), e0.data_ptr(), None,
            _lib.ptr_array([h.data_ptr() for h in older]), n_hist, sample.data_ptr(), x_out.data_ptr(),
            out2.data_ptr() if out2 is not None else None, out2.stride(0) if out2 is not None else 0,
            coef_ptr, od + 2, od, dt, flags, B, N, stream)
        _lib.check(rc, "consolver_step_fm")

        self._hist = [e0] + older
        self._step_index += 1
        tr.count += 1
        self._curr_sigma = self.sigmas[si + 1]

        o = tr.out
        actions, probs, masks = (None, None, None) if fixed else (o["actions"][i], o["probs"][i], o["masks"][i])
        conds = lazy_conds(conds_x, list(self._hist), od)
        if not return_dict:
            return (x_out, actions, probs, conds, masks)
        return FMPPOSchedulerOutput(prev_sample=x_out, actions=actions, probs=probs, conds=conds, masks=masks)

    def last_policy(self):
        """see PPOScheduler.last_policy; the table row is indexed by the sigma index of the step"""
        return self._traj.last(table_row=(self._step_index - 1) % self._traj.n)

    def scale_noise(self, sample: torch.Tensor, timestep, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Forward process of flow matching (edit_ppo/scheduler_fmppo.py:457-484); not on the hot path."""
        sigmas = self.sigmas.to(device=sample.device, dtype=sample.dtype)
        if self._begin_index is None:
            idx = [self.index_for_timestep(t) for t in timestep]
        elif self._step_index is not None:
            idx = [self._step_index] * timestep.shape[0]
        else:
            idx = [self._begin_index] * timestep.shape[0]
        sigma = sigmas[idx].flatten()
        while sigma.dim() < sample.dim():
            sigma = sigma.unsqueeze(-1)
        return sigma * noise + (1.0 - sigma) * sample

    def __len__(self):
        return self.config.num_train_timesteps
