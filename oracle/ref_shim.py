"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference from /root/reference.

The reference (G-U-N/consolver) is pure Python but does not import as shipped in this image:
  * `diffusers` is not installed (scheduler_ppo.py:19-20, edit_ppo/scheduler_fmppo.py:22-24),
  * `factor_net_ppo_continous` is imported (scheduler_ppo.py:23) but absent from the reference tree,
  * the SD and FM policies are two different modules that are both called `factor_net_ppo`.

This file installs minimal stand-ins for the diffusers base classes (they contain no arithmetic — only
`self.config` plumbing), a one-class stub for the missing module, and loads the SD / FM variants under
distinct module names.  Nothing from the reference is copied; the reference source files are executed
where they lie.  /root/reference exists only in the build container, so this module is used solely by
`oracle/make_golden.py` (to write tests/golden/*.npz), by the optional `-m "not gpu"` cross-check tests (which skip
when the tree is missing) and by bench.py's reference legs.  On the GPU box the same unmodified files are found under
the git-ignored oracle/_ref/ (oracle/stage_ref.py).  It is never imported by the product package.
"""
from __future__ import annotations

import contextlib
import enum
import functools
import importlib.util
import inspect
import io
import os
import sys
import types
from collections import OrderedDict

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _resolve_root() -> str:
    """CONSOLVER_REFERENCE_ROOT if set; else the live tree (build container); else the byte-identical copies that
    oracle/stage_ref.py put under the git-ignored oracle/_ref/ (the GPU box, where /root/reference does not exist)."""
    env = os.environ.get("CONSOLVER_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/scheduler_ppo.py"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _resolve_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "scheduler_ppo.py"))


def reference_is_staged_copy() -> bool:
    return os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_STAGED)


class _FrozenDict(OrderedDict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(name) from e


def _register_to_config(init):
    """Stand-in for diffusers.configuration_utils.register_to_config: record every ctor arg
    (including defaults) on `self.config` BEFORE running the body (FMPPOScheduler.__init__ reads
    self.config inside its own body, edit_ppo/scheduler_fmppo.py:132)."""
    sig = inspect.signature(init)

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = _FrozenDict((k, v) for k, v in bound.arguments.items() if k != "self")
        object.__setattr__(self, "config", cfg)
        init(self, *args, **kwargs)

    return wrapper


def _install_diffusers_standins():
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_consolver_standin", False):
        return  # a real diffusers is importable: use it
    import dataclasses

    diffusers = types.ModuleType("diffusers")
    diffusers._consolver_standin = True
    cu = types.ModuleType("diffusers.configuration_utils")
    su_pkg = types.ModuleType("diffusers.schedulers")
    su = types.ModuleType("diffusers.schedulers.scheduling_utils")
    ut = types.ModuleType("diffusers.utils")

    class ConfigMixin:  # no arithmetic; config plumbing only
        pass

    class SchedulerMixin:
        pass

    class BaseOutput(OrderedDict):
        def __post_init__(self):
            for f in dataclasses.fields(self):
                self[f.name] = getattr(self, f.name)

    @dataclasses.dataclass
    class SchedulerOutput(BaseOutput):
        prev_sample: object = None

    class KarrasDiffusionSchedulers(enum.Enum):
        DDIMScheduler = 1
        DDPMScheduler = 2
        PNDMScheduler = 3

    class _Logging:
        @staticmethod
        def get_logger(name):
            import logging

            return logging.getLogger(name)

    tu = types.ModuleType("diffusers.utils.torch_utils")

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        import torch

        return torch.randn(shape, generator=generator, device=device, dtype=dtype)

    tu.randn_tensor = randn_tensor
    ut.torch_utils = tu
    ut.deprecate = lambda *a, **k: None
    diffusers.DPMSolverMultistepScheduler = _dpm_base(ConfigMixin, SchedulerMixin)
    sys.modules["diffusers.utils.torch_utils"] = tu

    cu.ConfigMixin = ConfigMixin
    cu.register_to_config = _register_to_config
    su.SchedulerMixin = SchedulerMixin
    su.SchedulerOutput = SchedulerOutput
    su.KarrasDiffusionSchedulers = KarrasDiffusionSchedulers
    ut.BaseOutput = BaseOutput
    ut.is_scipy_available = lambda: importlib.util.find_spec("scipy") is not None
    ut.logging = _Logging
    diffusers.configuration_utils = cu
    diffusers.schedulers = su_pkg
    su_pkg.scheduling_utils = su
    diffusers.utils = ut
    sys.modules.update({
        "diffusers": diffusers,
        "diffusers.configuration_utils": cu,
        "diffusers.schedulers": su_pkg,
        "diffusers.schedulers.scheduling_utils": su,
        "diffusers.utils": ut,
    })


def _dpm_base(ConfigMixin, SchedulerMixin):
    """Stand-in for diffusers 0.26.3 `DPMSolverMultistepScheduler`, the base class of the reference's AMED plugin
    (diffusers_amed_plugin_dpmpp.py:22,:27).  diffusers is NOT in this image and not in the reference tree, so this
    is a restatement of the published library code — only the pieces the plugin inherits: the beta/sigma tables,
    `_sigma_to_alpha_sigma_t`, `convert_model_output`, step-index bookkeeping and the stock `set_timesteps`
    (basic sigma interpolation; Karras / Lu spacings and dynamic thresholding are not restated).  Everything
    the plugin overrides (custom-timestep `set_timesteps`, both update formulas, `step`) runs from the unmodified
    reference file on top of this.  PARITY NOTE: goldens made this way pin the plugin's own arithmetic; the
    inherited pieces are pinned only to this restatement."""
    import numpy as np
    import torch

    class DPMSolverMultistepScheduler(SchedulerMixin, ConfigMixin):
        order = 1

        @_register_to_config
        def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                     trained_betas=None, solver_order=2, prediction_type="epsilon", thresholding=False,
                     dynamic_thresholding_ratio=0.995, sample_max_value=1.0, algorithm_type="dpmsolver++",
                     solver_type="midpoint", lower_order_final=True, euler_at_final=False,
                     use_karras_sigmas=False, use_lu_lambdas=False, final_sigmas_type="zero",
                     lambda_min_clipped=-float("inf"), variance_type=None, timestep_spacing="linspace",
                     steps_offset=0):
            if trained_betas is not None:
                self.betas = torch.tensor(trained_betas, dtype=torch.float32)
            elif beta_schedule == "linear":
                self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
            elif beta_schedule == "scaled_linear":
                self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                            dtype=torch.float32) ** 2
            else:
                raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
            self.alphas = 1.0 - self.betas
            self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
            self.alpha_t = torch.sqrt(self.alphas_cumprod)
            self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
            self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
            self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
            self.init_noise_sigma = 1.0
            if algorithm_type not in ("dpmsolver", "dpmsolver++", "sde-dpmsolver", "sde-dpmsolver++"):
                raise NotImplementedError(f"{algorithm_type} is not implemented for {self.__class__}")
            if solver_type not in ("midpoint", "heun"):
                raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
            self.num_inference_steps = None
            timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy()
            self.timesteps = torch.from_numpy(timesteps)
            self.model_outputs = [None] * solver_order
            self.lower_order_nums = 0
            self._step_index = None
            self._begin_index = None
            self.sigmas = self.sigmas.to("cpu")

        @property
        def step_index(self):
            return self._step_index

        @property
        def begin_index(self):
            return self._begin_index

        def set_begin_index(self, begin_index=0):
            self._begin_index = begin_index

        def set_timesteps(self, num_inference_steps=None, device=None):
            cfg = self.config
            clipped_idx = torch.searchsorted(torch.flip(self.lambda_t, [0]), cfg.lambda_min_clipped)
            last_timestep = ((cfg.num_train_timesteps - clipped_idx).numpy()).item()
            if cfg.timestep_spacing == "linspace":
                timesteps = (np.linspace(0, last_timestep - 1, num_inference_steps + 1).round()[::-1][:-1]
                             .copy().astype(np.int64))
            elif cfg.timestep_spacing == "leading":
                step_ratio = last_timestep // (num_inference_steps + 1)
                timesteps = (np.arange(0, num_inference_steps + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
                timesteps += cfg.steps_offset
            elif cfg.timestep_spacing == "trailing":
                step_ratio = cfg.num_train_timesteps / num_inference_steps
                timesteps = np.arange(last_timestep, 0, -step_ratio).round().copy().astype(np.int64)
                timesteps -= 1
            else:
                raise ValueError(f"{cfg.timestep_spacing} is not supported.")
            sigmas = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
            if cfg.use_karras_sigmas or cfg.use_lu_lambdas:
                raise NotImplementedError("Karras / Lu spacings are not restated in this stand-in")
            sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
            if cfg.final_sigmas_type == "sigma_min":
                sigma_last = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5
            elif cfg.final_sigmas_type == "zero":
                sigma_last = 0
            else:
                raise ValueError(f"`final_sigmas_type` must be one of 'zero', or 'sigma_min', got {cfg.final_sigmas_type}")
            sigmas = np.concatenate([sigmas, [sigma_last]]).astype(np.float32)
            self.sigmas = torch.from_numpy(sigmas)
            self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
            self.num_inference_steps = len(timesteps)
            self.model_outputs = [None] * cfg.solver_order
            self.lower_order_nums = 0
            self._step_index = None
            self._begin_index = None
            self.sigmas = self.sigmas.to("cpu")

        def _sigma_to_alpha_sigma_t(self, sigma):
            alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
            sigma_t = sigma * alpha_t
            return alpha_t, sigma_t

        def convert_model_output(self, model_output, *args, sample=None, **kwargs):
            cfg = self.config
            if cfg.thresholding:
                raise NotImplementedError("dynamic thresholding is not restated in this stand-in")
            sigma = self.sigmas[self.step_index]
            alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma)
            if cfg.algorithm_type in ("dpmsolver++", "sde-dpmsolver++"):
                if cfg.prediction_type == "epsilon":
                    return (sample - sigma_t * model_output) / alpha_t
                if cfg.prediction_type == "sample":
                    return model_output
                if cfg.prediction_type == "v_prediction":
                    return alpha_t * sample - sigma_t * model_output
            else:
                if cfg.prediction_type == "epsilon":
                    return model_output
                if cfg.prediction_type == "sample":
                    return (sample - alpha_t * model_output) / sigma_t
                if cfg.prediction_type == "v_prediction":
                    return alpha_t * model_output + sigma_t * sample
            raise ValueError(f"prediction_type given as {cfg.prediction_type} must be one of `epsilon`, `sample`, or"
                             " `v_prediction` for the DPMSolverMultistepScheduler.")

        def index_for_timestep(self, timestep, schedule_timesteps=None):
            if schedule_timesteps is None:
                schedule_timesteps = self.timesteps
            cand = (schedule_timesteps == timestep).nonzero()
            if len(cand) == 0:
                return len(self.timesteps) - 1
            if len(cand) > 1:
                return cand[1].item()
            return cand[0].item()

        def _init_step_index(self, timestep):
            if self.begin_index is None:
                if isinstance(timestep, torch.Tensor):
                    timestep = timestep.to(self.timesteps.device)
                self._step_index = self.index_for_timestep(timestep)
            else:
                self._step_index = self._begin_index

        def scale_model_input(self, sample, *args, **kwargs):
            return sample

        def __len__(self):
            return self.config.num_train_timesteps

    return DPMSolverMultistepScheduler


def _load(name: str, path: str, aliases: dict):
    """Execute reference file `path` as module `name`, with `aliases` (import-name -> module)
    temporarily placed in sys.modules so the file's own `import factor_net_ppo` resolves to
    the right variant."""
    saved = {k: sys.modules.get(k) for k in aliases}
    sys.modules.update(aliases)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@functools.lru_cache(maxsize=None)
def load_reference():
    """Returns a namespace with the reference's PPOScheduler, FMPPOScheduler, the baseline
    FlowMatchGeneralDiscreteScheduler and both FactorNetPPO classes (sd / fm), loaded from REFERENCE_ROOT
    without modification."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_diffusers_standins()
    r = REFERENCE_ROOT
    conv = _load("_ref_conv_net", os.path.join(r, "conv_net.py"), {})
    fn_sd = _load("_ref_factor_net_sd", os.path.join(r, "factor_net_ppo.py"), {"conv_net": conv})
    cont = types.ModuleType("factor_net_ppo_continous")  # absent from the reference (scheduler_ppo.py:23)

    class FactorNetPPOContinous:  # noqa: N801 - name dictated by the reference import
        def __init__(self, *a, **k):
            raise NotImplementedError("factor_net_ppo_continous is not part of the reference tree")

    cont.FactorNetPPOContinous = FactorNetPPOContinous
    sched_sd = _load("_ref_scheduler_ppo", os.path.join(r, "scheduler_ppo.py"),
                     {"factor_net_ppo": fn_sd, "factor_net_ppo_continous": cont, "conv_net": conv})
    conv_fm = _load("_ref_conv_net_fm", os.path.join(r, "edit_ppo", "conv_net.py"), {})
    fn_fm = _load("_ref_factor_net_fm", os.path.join(r, "edit_ppo", "factor_net_ppo.py"), {"conv_net": conv_fm})
    sched_fm = _load("_ref_scheduler_fmppo", os.path.join(r, "edit_ppo", "scheduler_fmppo.py"),
                     {"factor_net_ppo": fn_fm, "conv_net": conv_fm})
    amed = _load("_ref_amed_plugin", os.path.join(r, "diffusers_amed_plugin_dpmpp.py"), {})
    sched_fm_base = _load("_ref_scheduler_fm", os.path.join(r, "edit_ppo", "scheduler_fm.py"), {})
    return types.SimpleNamespace(
        FlowMatchGeneralDiscreteScheduler=sched_fm_base.FlowMatchGeneralDiscreteScheduler,
        AMEDDPMSolverMultistepScheduler=amed.DPMSolverMultistepScheduler,
        PPOScheduler=sched_sd.PPOScheduler,
        FMPPOScheduler=sched_fm.FMPPOScheduler,
        FactorNetPPO_SD=fn_sd.FactorNetPPO,
        FactorNetPPO_FM=fn_fm.FactorNetPPO,
    )


@contextlib.contextmanager
def quiet():
    """The reference prints on every step (scheduler_ppo.py:243,:289); silence it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


@contextlib.contextmanager
def devnull():
    """stdout -> /dev/null for timed runs of the reference (its per-step prints still format their tensors:
    that cost is the reference's own)."""
    with open(os.devnull, "w") as f, contextlib.redirect_stdout(f):
        yield
