"""DPMSolverMultistepScheduler with AMED scaling — the distillation-based baseline solver the reference ships as
`diffusers_amed_plugin_dpmpp.py` (class of the same name; used by gen_ppo.py:157-163,:288-307), on one fused CUDA
kernel per step.

What the plugin adds to diffusers' multistep DPM-Solver(++): a caller-supplied timestep list whose odd entries are
moved to where sigma equals `sigma * scale_times[i]` (:47-58, the time the denoiser is evaluated at), and a
per-step `scale_dirs[i]` that multiplies every model-output term of the update (:121,:205-208,:417).  Set both as
attributes before `set_timesteps(..., timesteps=[...])`, exactly like gen_ppo.py:292-296.  Without a timestep list
the stock diffusers grid is used and `scale_dirs` defaults to ones (the plugin itself requires it to be set).

The update of step i is   x' = cx*x - a0*m0 [- a1*D1 [- a2*D2]]   with m the converted model outputs, D1 / D2 their
first / second divided differences
(data prediction for dpmsolver++).  All scalars are evaluated on the host with the same 0-d fp32 torch expressions
the plugin uses (so they carry its roundings) once per `set_timesteps`; `consolver_step_dpm` then does CFG combine,
conversion, update and the write of m0 into a three-slot ring in one pass over HBM (5-7 latent-sized tensors per
step instead of ~20 for the op-by-op version).

Scope: the ODE variants (`dpmsolver++`, `dpmsolver`), solver_order 1-3, midpoint / heun, epsilon / sample /
v_prediction; no dynamic thresholding, no SDE variants, no Karras / Lu spacings (the reference's AMED run uses none
of them).  diffusers itself (0.26.3, env.yaml:52) is not part of the reference tree: the inherited pieces
(`convert_model_output`, sigma tables, stock grid) follow the published library algorithm; see DESIGN.md §7 for
what the golden vectors pin."""
from __future__ import annotations

import ctypes
import dataclasses
from typing import List, Optional, Union

import numpy as np
import torch

from . import _lib
from .config_utils import BaseOutput, ConfigMixin, SchedulerMixin, register_to_config


@dataclasses.dataclass
class SchedulerOutput(BaseOutput):
    prev_sample: torch.Tensor = None


@dataclasses.dataclass
class _StepPlan:
    """host scalars of one step (python floats holding fp32 values)"""
    convert: int
    ck0: float
    ck1: float
    cx: float
    a0: float
    a1: Optional[float]       # second-order terms; None at step 0 (no previous model output)
    rinv: Optional[float]
    third: Optional[tuple] = None     # (a1, rinv, a2, rinv1, w, rs) of the third-order form; None at steps 0 and 1


class DPMSolverMultistepScheduler(SchedulerMixin, ConfigMixin):
    _compatibles = []
    order = 1

    @register_to_config
    def __init__(
        self,
        num_train_timesteps: int = 1000,
        beta_start: float = 0.0001,
        beta_end: float = 0.02,
        beta_schedule: str = "linear",
        trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
        solver_order: int = 2,
        prediction_type: str = "epsilon",
        thresholding: bool = False,
        dynamic_thresholding_ratio: float = 0.995,
        sample_max_value: float = 1.0,
        algorithm_type: str = "dpmsolver++",
        solver_type: str = "midpoint",
        lower_order_final: bool = True,
        euler_at_final: bool = False,
        use_karras_sigmas: Optional[bool] = False,
        use_lu_lambdas: Optional[bool] = False,
        final_sigmas_type: Optional[str] = "zero",
        lambda_min_clipped: float = -float("inf"),
        variance_type: Optional[str] = None,
        timestep_spacing: str = "linspace",
        steps_offset: int = 0,
    ):
        if trained_betas is not None:
            betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if algorithm_type not in ("dpmsolver", "dpmsolver++"):
            raise NotImplementedError(f"algorithm_type {algorithm_type!r}: only the ODE variants 'dpmsolver' and "
                                      "'dpmsolver++' are built on the fused kernel")
        if solver_type not in ("midpoint", "heun"):
            raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        if solver_order not in (1, 2, 3):
            raise NotImplementedError(f"solver_order {solver_order}: the multistep solver has orders 1 to 3")
        if thresholding or use_karras_sigmas or use_lu_lambdas or variance_type in ("learned", "learned_range"):
            raise NotImplementedError("thresholding / Karras / Lu spacings / learned variance are not built")
        if prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample`, or "
                             "`v_prediction` for the DPMSolverMultistepScheduler.")
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).to("cpu")
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(
            np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy())
        self.lower_order_nums = 0
        self._step_index = None
        self._begin_index = None
        self._plans = None
        self._ring = None
        self._n_prev = 0
        #: "cuda" (default): `(sample - sigma*e) / alpha` of convert_model_output as ATen evaluates it on CUDA tensors (a
        #: multiplication by the fp32 reciprocal of the host-resident scalar); "cpu": the true division of a run on CPU
        #: tensors (what the amed_* fixtures were made under).  See _sched_common.SolverOptions.reference_device.
        from . import _sched_common

        self.reference_device = _sched_common.DEFAULT_REFERENCE_DEVICE

    # ---- reference / diffusers surface -------------------------------------------------------------------------
    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def scale_model_input(self, sample: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return sample

    def __len__(self):
        return self.config.num_train_timesteps

    def _all_sigmas(self) -> np.ndarray:
        return (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()

    def set_timesteps(self, num_inference_steps: int = None, device: Union[str, torch.device] = None,
                      timesteps: List[int] = None):
        """diffusers_amed_plugin_dpmpp.py:29-68; without `timesteps` the stock diffusers grid."""
        cfg = self.config
        all_sigmas = self._all_sigmas()
        if timesteps is None:
            n, T = num_inference_steps, cfg.num_train_timesteps
            if cfg.timestep_spacing == "linspace":
                ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
            elif cfg.timestep_spacing == "leading":
                ts = (np.arange(0, n + 1) * (T // (n + 1))).round()[::-1][:-1].copy().astype(np.int64)
                ts += cfg.steps_offset
            elif cfg.timestep_spacing == "trailing":
                ts = np.arange(T, 0, -(T / n)).round().copy().astype(np.int64) - 1
            else:
                raise ValueError(f"{cfg.timestep_spacing} is not supported. Please make sure to choose one of "
                                 "'linspace', 'leading' or 'trailing'.")
            sig = np.interp(ts, np.arange(0, len(all_sigmas)), all_sigmas)
            if cfg.final_sigmas_type == "sigma_min":
                last = all_sigmas[0]
            elif cfg.final_sigmas_type == "zero":
                last = 0
            else:
                raise ValueError("`final_sigmas_type` must be one of 'zero', or 'sigma_min', but got "
                                 f"{cfg.final_sigmas_type}")
            self.sigmas = torch.from_numpy(np.concatenate([sig, [last]]).astype(np.float32))
            host_ts = ts
            self.num_inference_steps = len(ts)
        else:
            if not hasattr(self, "scale_dirs") or not hasattr(self, "scale_times"):      # asserts at :48-49
                raise AssertionError("scale_dirs and scale_times must be set before calling set_timesteps")
            timesteps = [int(t) for t in timesteps]
            self.sigmas = torch.from_numpy(all_sigmas[timesteps])
            host_ts = np.array(timesteps[:-1], dtype=np.int64)                           # the trailing 0 is dropped
            for i in range(len(self.scale_times)):                                        # :55-58
                if i % 2 == 1:
                    target = self.sigmas[i] * self.scale_times[i]
                    src = torch.from_numpy(all_sigmas[timesteps[i + 1] + 1:timesteps[i - 1]])
                    host_ts[i] = timesteps[i + 1] + 1 + int(torch.argmin(torch.abs(src - target)))
            self.num_inference_steps = len(timesteps)                                     # :60
        self._timesteps_host = host_ts
        self.timesteps = torch.from_numpy(host_ts.copy()).to(device=device, dtype=torch.int64)
        self.lower_order_nums = 0
        self._step_index = None
        self._begin_index = None
        self._plans = None
        self._n_prev = 0

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        grid = self._timesteps_host if schedule_timesteps is None else schedule_timesteps.detach().cpu().numpy()
        tv = None
        ts = self.timesteps
        if (schedule_timesteps is None and isinstance(timestep, torch.Tensor) and timestep.dim() == 0 and ts.is_cuda
                and timestep.dtype == ts.dtype
                and timestep.untyped_storage().data_ptr() == ts.untyped_storage().data_ptr()):
            j = timestep.storage_offset() - ts.storage_offset()       # a view of our own grid tensor: no read-back
            if 0 <= j < len(grid):
                tv = int(grid[j])
        if tv is None:
            tv = int(timestep.item() if isinstance(timestep, torch.Tensor) else timestep)
        hits = np.nonzero(grid == tv)[0]
        if len(hits) == 0:
            return len(self._timesteps_host) - 1
        return int(hits[1 if len(hits) > 1 else 0])

    def _init_step_index(self, timestep):
        self._step_index = self.index_for_timestep(timestep) if self._begin_index is None else self._begin_index

    # ---- host scalars ------------------------------------------------------------------------------------------
    @staticmethod
    def _alpha_sigma(sigma):
        alpha = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha, sigma * alpha

    def _build_plans(self):
        """One _StepPlan per step index, from the plugin's own 0-d fp32 tensor expressions."""
        cfg = self.config
        pp = cfg.algorithm_type == "dpmsolver++"
        sg = self.sigmas
        dirs = getattr(self, "scale_dirs", None)
        plans = []
        for i in range(len(sg) - 1):
            sd = 1.0 if dirs is None else dirs[i]
            alpha_t, sigma_t = self._alpha_sigma(sg[i + 1])
            alpha_s, sigma_s = self._alpha_sigma(sg[i])
            # conversion of the model output (diffusers convert_model_output)
            if cfg.prediction_type == "epsilon":
                conv = (_lib.DPM_CONVERT_DIV, sigma_s, alpha_s) if pp else (_lib.DPM_CONVERT_NONE, sigma_s * 0, alpha_s)
            elif cfg.prediction_type == "sample":
                conv = (_lib.DPM_CONVERT_NONE, sigma_s * 0, alpha_s) if pp else (_lib.DPM_CONVERT_DIV, alpha_s, sigma_s)
            else:
                conv = (_lib.DPM_CONVERT_LIN, -sigma_s, alpha_s) if pp else (_lib.DPM_CONVERT_LIN, alpha_s, sigma_s)
            lam_t = torch.log(alpha_t) - torch.log(sigma_t)
            lam_s = torch.log(alpha_s) - torch.log(sigma_s)
            h = lam_t - lam_s
            if pp:
                cx = sigma_t / sigma_s
                a0 = sd * (alpha_t * (torch.exp(-h) - 1.0))                               # :121, :207
            else:
                cx = alpha_t / alpha_s
                a0 = sd * (sigma_t * (torch.exp(h) - 1.0))                                # :123, :221
            a1 = rinv = None
            if i > 0:
                alpha_p, sigma_p = self._alpha_sigma(sg[i - 1])
                h_0 = lam_s - (torch.log(alpha_p) - torch.log(sigma_p))
                rinv = 1.0 / (h_0 / h)                                                    # :200-201
                if pp and cfg.solver_type == "midpoint":
                    a1 = sd * 0.5 * (alpha_t * (torch.exp(-h) - 1.0))                     # :208
                elif pp:
                    a1 = -(sd * (alpha_t * ((torch.exp(-h) - 1.0) / h + 1.0)))            # :214 (added there)
                elif cfg.solver_type == "midpoint":
                    a1 = sd * 0.5 * (sigma_t * (torch.exp(h) - 1.0))                      # :222
                else:
                    a1 = sd * (sigma_t * ((torch.exp(h) - 1.0) / h - 1.0))                # :228
            third = None
            if i > 1 and cfg.solver_order == 3:                                           # :307-346
                alpha_q, sigma_q = self._alpha_sigma(sg[i - 2])
                h_1 = (torch.log(alpha_p) - torch.log(sigma_p)) - (torch.log(alpha_q) - torch.log(sigma_q))
                r0, r1 = h_0 / h, h_1 / h
                if pp:
                    b1 = -(sd * (alpha_t * ((torch.exp(-h) - 1.0) / h + 1.0)))            # :337 (added there)
                    b2 = sd * (alpha_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5))      # :338
                else:
                    b1 = sd * (sigma_t * ((torch.exp(h) - 1.0) / h - 1.0))                # :345
                    b2 = sd * (sigma_t * ((torch.exp(h) - 1.0 - h) / h ** 2 - 0.5))       # :346
                third = tuple(float(v) for v in (b1, 1.0 / r0, b2, 1.0 / r1, r0 / (r0 + r1), 1.0 / (r0 + r1)))
            f = lambda v: None if v is None else float(v)  # noqa: E731
            plans.append(_StepPlan(conv[0], float(conv[1]), float(conv[2]), float(cx), float(a0), f(a1), f(rinv),
                                   third))
        self._plans = plans
        self._plans_dirs = None if dirs is None else tuple(dirs)

    # ---- the step ----------------------------------------------------------------------------------------------
    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None,
             variance_noise: Optional[torch.Tensor] = None, return_dict: bool = True,
             out2: Optional[torch.Tensor] = None):
        """diffusers_amed_plugin_dpmpp.py:350-436."""
        x = self._step(model_output, None, 0.0, timestep, sample, out2)
        return SchedulerOutput(prev_sample=x) if return_dict else (x,)

    def step_cfg(self, noise_pred_pair: torch.Tensor, timestep, sample: torch.Tensor, guidance_scale: float,
                 return_dict: bool = False, out2: Optional[torch.Tensor] = None):
        """Classifier-free-guidance combine fused into the step: `noise_pred_pair` is the denoiser's [2B, ...]
        output, unconditional half first (gen_pretrain/pipeline.py:1069-1071)."""
        B = sample.shape[0]
        if noise_pred_pair.shape[0] != 2 * B or not noise_pred_pair.is_contiguous():
            raise ValueError("noise_pred_pair must be a contiguous [2B, ...] tensor (unconditional half first)")
        x = self._step(noise_pred_pair[:B], noise_pred_pair[B:], float(guidance_scale), timestep, sample, out2)
        return SchedulerOutput(prev_sample=x) if return_dict else (x,)

    def _plan_for_step(self, timestep):
        """(step index, host scalars, order of the update) of the step about to be taken; no device work"""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if self._step_index is None:
            self._init_step_index(timestep)
        dirs = getattr(self, "scale_dirs", None)
        if self._plans is None or self._plans_dirs != (None if dirs is None else tuple(dirs)):
            self._build_plans()
        cfg = self.config
        i, nts = self._step_index, len(self._timesteps_host)
        if i >= len(self._plans):
            raise IndexError("DPMSolverMultistepScheduler.step called past the end of the sigma schedule")
        final_first = (i == nts - 1) and (cfg.euler_at_final or (cfg.lower_order_final and nts < 15)
                                          or cfg.final_sigmas_type == "zero")               # :394-398
        second_only = (i == nts - 2) and cfg.lower_order_final and nts < 15                # :399-401
        if cfg.solver_order == 1 or self.lower_order_nums < 1 or final_first:              # :418-423
            order = 1
        elif cfg.solver_order == 2 or self.lower_order_nums < 2 or second_only:
            order = 2
        else:
            order = 3
        return i, self._plans[i], order

    def _advance(self):
        if self.lower_order_nums < self.config.solver_order:                               # :425-426
            self.lower_order_nums += 1
        self._step_index += 1                                                              # :432

    def _step(self, e0, cond, guidance, timestep, sample, out2):
        if not (e0.is_cuda and sample.is_cuda):
            raise RuntimeError("consolver_b200 has no CPU path: model_output and sample must be CUDA tensors")
        i, p, order = self._plan_for_step(timestep)
        e0 = e0 if e0.is_contiguous() else e0.contiguous()
        sample = sample if sample.is_contiguous() else sample.contiguous()
        if sample.dtype not in (e0.dtype, torch.float32):
            sample = sample.float()                                                        # :409
        B = e0.shape[0]
        N = e0.numel() // B
        ring = self._ring
        if ring is None or ring.shape[1:] != e0.shape or ring.dtype != e0.dtype or ring.device != e0.device:
            ring = self._ring = torch.empty((3,) + tuple(e0.shape), device=e0.device, dtype=e0.dtype)
            self._n_prev = 0
        if order - 1 > self._n_prev:
            raise RuntimeError(f"order-{order} step with only {self._n_prev} earlier model outputs in the ring")
        slot = ring[i % 3]                              # converted outputs of steps i, i-1, i-2 live in i, i-1, i-2 mod 3
        m1 = ring[(i - 1) % 3].data_ptr() if order >= 2 else None
        m2 = ring[(i - 2) % 3].data_ptr() if order == 3 else None
        upd = _lib.DpmUpdate(cx=p.cx, a0=p.a0)
        if order == 2:
            upd.a1, upd.rinv = p.a1, p.rinv
        elif order == 3:
            upd.a1, upd.rinv, upd.a2, upd.rinv1, upd.w, upd.rs = p.third
        x_out = torch.empty(e0.shape, device=e0.device, dtype=e0.dtype)                    # :429: model dtype
        stream = torch._C._cuda_getCurrentRawStream(e0.device.index)
        rc = _lib.load().consolver_step_dpm(
            _lib.dtype_code(e0.dtype), _lib.dtype_code(sample.dtype), e0.data_ptr(),
            cond.data_ptr() if cond is not None else None, guidance, slot.data_ptr(), m1, m2,
            sample.data_ptr(), x_out.data_ptr(),
            out2.data_ptr() if out2 is not None else None, out2.stride(0) if out2 is not None else 0,
            (_lib.DPM_CONVERT_DIV_RECIP if p.convert == _lib.DPM_CONVERT_DIV and self.reference_device == "cuda"
             else p.convert), p.ck0, p.ck1, ctypes.byref(upd), B, N, stream)
        _lib.check(rc, "consolver_step_dpm")
        self._n_prev = min(self._n_prev + 1, 2)
        self._advance()
        return x_out

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """diffusers' forward noising on the VP sigmas of the current grid; not on the hot path."""
        sigmas = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
        if self._begin_index is None:
            idx = [self.index_for_timestep(t) for t in timesteps]
        elif self._step_index is not None:
            idx = [self._step_index] * timesteps.shape[0]
        else:
            idx = [self._begin_index] * timesteps.shape[0]
        sigma = sigmas[idx].flatten()
        while sigma.dim() < original_samples.dim():
            sigma = sigma.unsqueeze(-1)
        alpha_t, sigma_t = self._alpha_sigma(sigma)
        return alpha_t * original_samples + sigma_t * noise
