"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference from /root/reference.

The reference (G-U-N/consolver) is pure Python but does not import as shipped in this image:
  * `diffusers` is not installed (scheduler_ppo.py:19-20, edit_ppo/scheduler_fmppo.py:22-24),
  * `factor_net_ppo_continous` is imported (scheduler_ppo.py:23) but absent from the reference tree,
  * the SD and FM policies are two different modules that are both called `factor_net_ppo`.

This file installs minimal stand-ins for the diffusers base classes (they contain no arithmetic — only
`self.config` plumbing), a one-class stub for the missing module, and loads the SD / FM variants under
distinct module names.  Nothing from the reference is copied; the reference source files are executed
where they lie.  /root/reference exists only in the build container, so this module is used solely by
`oracle/make_golden.py` (to write tests/golden/*.npz) and by the optional `-m "not gpu"` cross-check
tests, which skip when the tree is missing.  It is never imported by the product package.
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

REFERENCE_ROOT = os.environ.get("CONSOLVER_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "scheduler_ppo.py"))


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
    sched_fm_base = _load("_ref_scheduler_fm", os.path.join(r, "edit_ppo", "scheduler_fm.py"), {})
    return types.SimpleNamespace(
        FlowMatchGeneralDiscreteScheduler=sched_fm_base.FlowMatchGeneralDiscreteScheduler,
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
