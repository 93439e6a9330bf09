"""FMPPOScheduler — drop-in for the reference's flow-matching solver `edit_ppo/scheduler_fmppo.FMPPOScheduler`
(edit_ppo/scheduler_fmppo.py:56-553): same constructor kwargs, `set_timesteps(num_inference_steps, device,
sigmas, mu, timesteps)`, `set_begin_index`, `step(...)` signature (incl. the ignored s_churn/... arguments),
`scale_noise`, and `factor_net` state_dict.  The step runs the policy kernel and the fused Euler-form step kernel
(bf16/fp16/fp32 I/O, fp32 math) through the C ABI; see scheduler_ppo.py in this package for the design notes.
`per_token_timesteps` (edit_ppo/scheduler_fmppo.py:363-371) is exercised by no caller and is not supported."""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from . import _lib
from ._fm_schedule import FlowSigmaSchedule
from ._sched_common import SolverOptions, Trajectory, draw_source, lazy_conds, strided_model_outputs
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


class FMPPOScheduler(FlowSigmaSchedule, SolverOptions, SchedulerMixin, ConfigMixin):
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
        self._check_sigma_options(use_beta_sigmas, use_exponential_sigmas, use_karras_sigmas, time_shift_type)
        if order_dim < 2 or order_dim > _lib.MAX_ORDER:
            raise ValueError(f"order_dim must be in [2, {_lib.MAX_ORDER}]")
        if scaler_dim not in (0, 1, 2):
            raise NotImplementedError("More than two scale parameters not supported.")

        self._init_sigma_grid(num_train_timesteps, shift, use_dynamic_shifting)

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

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device: Union[str, torch.device] = None,
                      sigmas: Optional[List[float]] = None, mu: Optional[float] = None,
                      timesteps: Optional[List[float]] = None):
        """edit_ppo/scheduler_fmppo.py:171-245."""
        self._set_sigma_schedule(num_inference_steps, device, sigmas, mu, timesteps)
        self._hist = []
        self._traj = None
        self._curr_sigma = None

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
        e0 = model_output                               # may be a sample-strided view: see strided_model_outputs below
        sample = sample if sample.is_contiguous() else sample.contiguous()
        if sample.dtype != e0.dtype:
            sample = sample.float()                     # the reference upcasts the sample anyway (:354)
        if out2 is not None and (out2.dtype != e0.dtype or out2.shape != e0.shape or not out2[0].is_contiguous()):
            raise ValueError("out2 must have the model output's dtype and shape, with contiguous samples "
                             "(the next latent is returned in the model dtype, :436)")
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
        if fn.use_conv and self.fixed_coefficients is None:      # the feature reduction reads contiguous tensors
            e0, e_stride = (e0 if e0.is_contiguous() else e0.contiguous()), 0
        else:
            e0, older, e_stride = strided_model_outputs(e0, older)
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
        # the FM update itself has no host-resident scalars (sigmas live on the device; dt is a CUDA 0-d tensor): only
        # the policy's `logits / 0.01` and autocast depend on where / how the reference runs
        _, pflags, act_dt = self._semantics(e0.dtype)
        pflags &= ~(_lib.POLICY_COEF_F16 | _lib.POLICY_COEF_BF16)
        if fixed:
            coef_ptr = tr.fixed_rows(self.fixed_coefficients, n_hist).data_ptr()      # baseline solvers: no policy
        elif fn.use_conv:
            # use_conv=True: cosine features of the history (pass 1), per-sample MLP, then the fused step (pass 2)
            from .features import cosine_features_cuda

            feat, ws, full = tr.conv_buffers(fn)
            cosine_features_cuda(e0, None, 0.0, older, od, feat, ws, stream)
            rc = lib.consolver_policy_f32(
                *fn.kernel_weights(act_dt), x0, x1, fn.x_div, fn.temperature, feat.data_ptr(), od - 1, q_ptr, idx_ptr,
                B, fn.hidden_dim, fn.action_dims, fn.num_actions, od, cfg.scaler_dim, n_hist, pflags,
                full[i].data_ptr(), *outs, stream)
            _lib.check(rc, "consolver_policy_f32")
            flags |= _lib.FLAG_PDL if self.use_pdl else 0
        else:
            # all n (sigma, sigma_next) rows of the schedule go through the MLP in one launch per pass
            if tr.table_pass != tr.count // tr.n:
                fn.policy_tables(tr.condx_f32, tr.out["probs_table"], stream, policy_flags=pflags, act_dtype=act_dt)
                tr.table_pass = tr.count // tr.n
                tr.policy_forked = False
            ps = self.policy_stream if rng_arg is not None else None      # two-stream form: see PPOScheduler._step
            if ps is not None:
                main = torch.cuda.current_stream(e0.device)
                if not tr.policy_forked:
                    ps.wait_stream(main)
                    tr.policy_forked = True
            rc = lib.consolver_policy_sample_f32(
                tr.p("probs_table", si), fn.kernel_weights(act_dt)[6], q_ptr, idx_ptr, rng_arg, None, B, fn.action_dims,
                fn.num_actions, od, cfg.scaler_dim, n_hist, pflags, *outs,
                ps.cuda_stream if ps is not None else stream)
            _lib.check(rc, "consolver_policy_sample_f32")
            if ps is not None:
                ev = torch.cuda.Event()
                ev.record(ps)
                main.wait_event(ev)
                flags |= _lib.FLAG_CHAIN if self.chain_steps else 0   # previous node on this stream is a step kernel
            elif self.use_pdl:
                flags |= _lib.FLAG_PDL
        rc = lib.consolver_step_fm_strided(
            _lib.dtype_code(e0.dtype), _lib.dtype_code(sample.dtype), e0.data_ptr(), e_stride, None,
            _lib.ptr_array([h.data_ptr() for h in older]), n_hist, sample.data_ptr(), x_out.data_ptr(),
            out2.data_ptr() if out2 is not None else None, out2.stride(0) if out2 is not None else 0,
            coef_ptr, od + 2, od, dt, flags, B, N, stream)
        _lib.check(rc, "consolver_step_fm_strided")

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
