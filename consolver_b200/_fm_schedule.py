"""Sigma schedule shared by the flow-matching schedulers: the learned solver (edit_ppo/scheduler_fmppo.py:142-245,
:457-550) and the training-free baselines (edit_ppo/scheduler_fm.py:117-353) carry the same schedule code in the
reference.  Host-side only; the grid is kept on the host as well as on the device so that stepping never reads a
sigma back from the GPU."""
from __future__ import annotations

import math
from typing import List, Optional, Union

import numpy as np
import torch


class FlowSigmaSchedule:
    """Mixin: expects `self.config` with the reference's schedule kwargs."""

    @staticmethod
    def _check_sigma_options(use_beta_sigmas, use_exponential_sigmas, use_karras_sigmas, time_shift_type):
        if use_beta_sigmas:
            try:
                import scipy.stats  # noqa: F401
            except ImportError as e:  # edit_ppo/scheduler_fmppo.py:132-133
                raise ImportError("Make sure to install scipy if you want to use beta sigmas.") from e
        if sum([bool(use_beta_sigmas), bool(use_exponential_sigmas), bool(use_karras_sigmas)]) > 1:
            raise ValueError("Only one of `use_beta_sigmas`, `use_exponential_sigmas`, `use_karras_sigmas` can be used.")
        if time_shift_type not in {"exponential", "linear"}:
            raise ValueError("`time_shift_type` must either be 'exponential' or 'linear'.")

    def _init_sigma_grid(self, num_train_timesteps, shift, use_dynamic_shifting):
        """default 1000-point grid (edit_ppo/scheduler_fmppo.py:142-151)"""
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

    def _set_sigma_schedule(self, num_inference_steps: Optional[int] = None,
                            device: Union[str, torch.device] = None, sigmas: Optional[List[float]] = None,
                            mu: Optional[float] = None, timesteps: Optional[List[float]] = None):
        """edit_ppo/scheduler_fmppo.py:171-245 (== edit_ppo/scheduler_fm.py:259-353)."""
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

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        """edit_ppo/scheduler_fmppo.py:501-506 on the host copy of the grid (one read-back if `timestep` lives on
        the GPU; pipelines avoid it with set_begin_index)."""
        if schedule_timesteps is None:
            grid = self._timesteps_host
        else:
            grid = schedule_timesteps.detach().float().cpu().numpy()
        tv = None
        ts = self.timesteps
        if (schedule_timesteps is None and isinstance(timestep, torch.Tensor) and timestep.dim() == 0 and ts.is_cuda
                and timestep.dtype == ts.dtype
                and timestep.untyped_storage().data_ptr() == ts.untyped_storage().data_ptr()):
            # `for t in scheduler.timesteps` hands out views of the grid tensor: read the value from the host copy
            j = timestep.storage_offset() - ts.storage_offset()
            if 0 <= j < len(grid):
                tv = np.float32(grid[j])
        if tv is None:
            tv = np.float32(timestep.item() if isinstance(timestep, torch.Tensor) else timestep)
        hits = np.nonzero(grid == tv)[0]
        return int(hits[1 if len(hits) > 1 else 0])

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            self._step_index = self.index_for_timestep(timestep)
        else:
            self._step_index = self._begin_index

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
