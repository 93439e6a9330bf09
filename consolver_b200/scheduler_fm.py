"""FlowMatchGeneralDiscreteScheduler — the reference's training-free flow-matching baselines
(`edit_ppo/scheduler_fm.py:46-488`: `type` = euler / heun / dpm-solver / dpm-solver-multistep) on the SAME fused
step kernel as the learned solver, so that speed tables compare solvers and not implementations.

Every step of every kind is one `consolver_step_fm` launch; what changes is which latent is the base of the update,
the step size and the (fixed) multipliers of the model-output history:

  kind                   stage                base latent        step size                     model outputs
  euler                  every step           sample             sigma[i+1] - sigma[i]         v
  heun                   even index           sample             sigma[i+2] - sigma[i]         v           (kept)
                         odd index            kept sample        0.5 * kept step               kept v + v
  dpm-solver             even index           sample             sigma[i+1] - sigma[i]         v           (kept)
                         odd index            kept sample        kept step + (s[i+1] - s[i])   v
  dpm-solver-multistep   index 0              sample             sigma[1] - sigma[0]           v           (kept)
                         index > 0            kept sample        kept step + (s[i+1] - s[i])   v   (then keep this step)

The reference upcasts the sample to fp32 and casts the result back (:399,:485); the kernel reads the latent in its
own dtype and does the same arithmetic in registers.  Kept tensors are held BY REFERENCE (the reference keeps the
model output by reference too, and the upcast sample when it is fp32 already): do not overwrite them in place
between the two stages.  As in the reference, `set_timesteps` does not clear the kept stage (its one-line override
at :141-145 is shadowed by the full definition at :259)."""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Union

import numpy as np
import torch

from . import _lib
from ._fm_schedule import FlowSigmaSchedule
from ._sched_common import strided_model_outputs
from .config_utils import BaseOutput, ConfigMixin, SchedulerMixin, register_to_config

SOLVER_TYPES = ("euler", "heun", "dpm-solver", "dpm-solver-multistep")
_ORDER_DIM = 2              # deepest history any of the kinds combines (heun: kept v + v)


@dataclasses.dataclass
class FlowMatchHeunDiscreteSchedulerOutput(BaseOutput):
    """edit_ppo/scheduler_fm.py:32-43."""
    prev_sample: torch.Tensor = None


class FlowMatchGeneralDiscreteScheduler(FlowSigmaSchedule, SchedulerMixin, ConfigMixin):
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
        type="euler",
    ):
        self._check_sigma_options(use_beta_sigmas, use_exponential_sigmas, use_karras_sigmas, time_shift_type)
        self._init_sigma_grid(num_train_timesteps, shift, use_dynamic_shifting)
        self.type = type
        self.prev_model_output = None
        self.prev_sample = None
        self.prev_dt = None
        self._ones = None

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device: Union[str, torch.device] = None,
                      sigmas: Optional[List[float]] = None, mu: Optional[float] = None,
                      timesteps: Optional[List[float]] = None):
        """edit_ppo/scheduler_fm.py:259-353."""
        self._set_sigma_schedule(num_inference_steps, device, sigmas, mu, timesteps)

    def _unit_coefficients(self, B: int, device) -> torch.Tensor:
        """[B, order_dim + 2] records of ones: multipliers (1, 1) of heun's second stage; unused by depth-1 steps"""
        o = self._ones
        if o is None or o.shape[0] < B or o.device != device:
            o = self._ones = torch.ones(B, _ORDER_DIM + 2, device=device, dtype=torch.float32)
        return o

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, s_churn: float = 0.0,
             s_tmin: float = 0.0, s_tmax: float = float("inf"), s_noise: float = 1.0,
             generator: Optional[torch.Generator] = None, per_token_timesteps: Optional[torch.Tensor] = None,
             return_dict: bool = True, out2: Optional[torch.Tensor] = None):
        """edit_ppo/scheduler_fm.py:384-488 (s_churn ... per_token_timesteps are accepted and unused there too).
        `out2`: optional second destination for the next latent, see FMPPOScheduler.step."""
        if not (model_output.is_cuda and sample.is_cuda):
            raise RuntimeError("consolver_b200 has no CPU path: model_output and sample must be CUDA tensors")
        if self.type not in SOLVER_TYPES:
            # the reference leaves `prev_sample` unbound here (UnboundLocalError at :485)
            raise ValueError(f"unknown solver type {self.type!r}; expected one of {SOLVER_TYPES}")
        if self._step_index is None:
            self._init_step_index(timestep)
        e0 = model_output             # may be a sample-strided view (noise_pred[:, :L]); resolved before the launch
        sample = sample if sample.is_contiguous() else sample.contiguous()
        if sample.dtype not in (e0.dtype, torch.float32):
            sample = sample.float()                                       # :399
        if out2 is not None and (out2.dtype != e0.dtype or out2.shape != e0.shape or not out2[0].is_contiguous()):
            raise ValueError("out2 must have the model output's dtype and shape, with contiguous samples")
        i, sg, f32 = self._step_index, self._sigmas_host, np.float32

        def sigma(j, clamp):
            return f32(sg[j]) if (not clamp or j < len(sg)) else f32(sg[-1])

        base, older, flags = sample, [], 0
        kind = self.type
        if kind == "euler":                                               # :405-410
            dt = sigma(i + 1, True) - sigma(i, False)
        elif kind == "heun":                                              # :412-430
            if i % 2 == 0:
                dt = sigma(i + 2, True) - sigma(i, False)
                self.prev_dt, self.prev_sample, self.prev_model_output = dt, sample, e0
            else:
                self._need_first_stage()
                dt = f32(0.5) * self.prev_dt
                base, older, flags = self.prev_sample, [self.prev_model_output], _lib.FLAG_LOWP_COMBINE
        elif kind == "dpm-solver":                                        # :431-452
            if i % 2 == 0:
                dt = sigma(i + 1, False) - sigma(i, False)
                self.prev_dt, self.prev_sample, self.prev_model_output = dt, sample, e0
            else:
                self._need_first_stage()
                dt = self.prev_dt + (sigma(i + 1, False) - sigma(i, False))
                base = self.prev_sample
        else:                                                             # dpm-solver-multistep, :454-483
            if i == 0:
                dt = sigma(1, False) - sigma(0, False)
                self.prev_dt, self.prev_sample, self.prev_model_output = dt, sample, e0
            else:
                self._need_first_stage()
                h = sigma(i + 1, False) - sigma(i, False)
                dt = self.prev_dt + h
                base = self.prev_sample
                self.prev_dt, self.prev_sample = h, sample
        for t in [base] + older:
            if t.shape != e0.shape or t.device != e0.device:
                raise ValueError("the kept first-stage tensors do not match this step's model output "
                                 f"({tuple(t.shape)} on {t.device} vs {tuple(e0.shape)} on {e0.device})")
        if older and older[0].dtype != e0.dtype:
            raise ValueError("model outputs of the two stages must have one dtype")

        e0, older, e_stride = strided_model_outputs(e0, older)
        B = e0.shape[0]
        N = e0.numel() // B
        x_out = torch.empty(e0.shape, device=e0.device, dtype=e0.dtype)    # :485: result in the model dtype
        stream = torch._C._cuda_getCurrentRawStream(e0.device.index)
        rc = _lib.load().consolver_step_fm_strided(
            _lib.dtype_code(e0.dtype), _lib.dtype_code(base.dtype), e0.data_ptr(), e_stride, None,
            _lib.ptr_array([h.data_ptr() for h in older]), len(older) + 1, base.data_ptr(), x_out.data_ptr(),
            out2.data_ptr() if out2 is not None else None, out2.stride(0) if out2 is not None else 0,
            self._unit_coefficients(B, e0.device).data_ptr(), _ORDER_DIM + 2, _ORDER_DIM, float(dt), flags, B, N, stream)
        _lib.check(rc, "consolver_step_fm_strided")
        self._step_index += 1
        if not return_dict:
            return (x_out,)
        return FlowMatchHeunDiscreteSchedulerOutput(prev_sample=x_out)

    def _need_first_stage(self):
        if self.prev_dt is None:
            raise RuntimeError(f"{self.type}: this step index is a second stage but no first stage was taken "
                               "(the reference fails on `None` arithmetic here)")
