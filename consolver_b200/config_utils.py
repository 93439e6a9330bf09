"""Scheduler base classes: diffusers' mixins when diffusers is importable, otherwise local stand-ins with the
same surface the reference relies on (`self.config` populated with every constructor argument incl.
defaults — diffusers.configuration_utils.register_to_config — attribute and `.get` access)."""
from __future__ import annotations

import functools
import inspect
from collections import OrderedDict

try:  # pragma: no cover - diffusers is not in the build image
    from diffusers.configuration_utils import ConfigMixin, register_to_config  # type: ignore
    from diffusers.schedulers.scheduling_utils import KarrasDiffusionSchedulers, SchedulerMixin  # type: ignore
    from diffusers.utils import BaseOutput  # type: ignore

    HAVE_DIFFUSERS = True
    KARRAS_COMPATIBLES = [e.name for e in KarrasDiffusionSchedulers]
except Exception:  # noqa: BLE001
    HAVE_DIFFUSERS = False
    KARRAS_COMPATIBLES = []

    class FrozenConfig(OrderedDict):
        def __getattr__(self, name):
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name) from None

        def __setattr__(self, name, value):
            raise AttributeError("scheduler config is read-only")

    class ConfigMixin:
        config_name = "scheduler_config.json"

        @classmethod
        def from_config(cls, config, **kwargs):
            cfg = dict(config)
            cfg.update(kwargs)
            params = inspect.signature(cls.__init__).parameters
            return cls(**{k: v for k, v in cfg.items() if k in params})

        @classmethod
        def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, **kwargs):
            """Local-directory form of diffusers' loader: read `<path>[/<subfolder>]/scheduler_config.json`, let
            kwargs override (edit_ppo/train_ppo.py:87 passes order_dim=... this way).  No hub access."""
            import json
            import os

            d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
            with open(os.path.join(d, cls.config_name)) as f:
                cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
            return cls.from_config(cfg, **kwargs)

        def save_config(self, save_directory):
            import json
            import os

            os.makedirs(save_directory, exist_ok=True)
            cfg = {"_class_name": type(self).__name__}
            cfg.update({k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in self.config.items()})
            with open(os.path.join(save_directory, self.config_name), "w") as f:
                json.dump(cfg, f, indent=2, sort_keys=True)

        save_pretrained = save_config

    class SchedulerMixin:
        pass

    def register_to_config(init):
        sig = inspect.signature(init)

        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = FrozenConfig((k, v) for k, v in bound.arguments.items() if k != "self")
            object.__setattr__(self, "config", cfg)
            init(self, *args, **kwargs)

        return wrapper

    class BaseOutput(OrderedDict):
        """dataclass-style output that is also a mapping and converts with to_tuple()"""

        def __post_init__(self):
            import dataclasses

            for f in dataclasses.fields(self):
                self[f.name] = getattr(self, f.name)

        def to_tuple(self):
            return tuple(self[k] for k in self.keys())


class LazyConds(dict):
    """The `conds` dict of the reference's step() return (`{'x': [B,2], 'epsilon': [B,order_dim,...]}`,
    scheduler_ppo.py:234-237).  'epsilon' — the newest-first, zero-padded stack of the history — is numerically
    dead unless use_conv=True, so it is materialised only when somebody reads it (torch.stack of the history
    tensors at that moment).  With `step_cfg` the history lives in the scheduler's ring, whose slots are overwritten
    order_dim steps later: `still_valid` (a callable) says whether the referenced slots still hold this step's history,
    and a late read raises instead of silently returning another step's model outputs."""

    def __init__(self, x, eps_thunk, still_valid=None):
        super().__init__(x=x)
        self._eps_thunk = eps_thunk
        self._still_valid = still_valid

    def _materialise(self):
        if self._eps_thunk is not None:
            if self._still_valid is not None and not self._still_valid():
                raise RuntimeError(
                    "conds['epsilon'] of this step was requested after the scheduler's history ring had been overwritten "
                    "by later step_cfg() calls (the ring holds order_dim model outputs). Read conds['epsilon'] before "
                    "stepping order_dim more times, or drive the scheduler through step(), which keeps the caller's "
                    "tensors by reference like the reference does.")
            thunk, self._eps_thunk = self._eps_thunk, None
            super().__setitem__("epsilon", thunk())

    def __iter__(self):
        self._materialise()
        return super().__iter__()

    def __len__(self):
        return 2

    def copy(self):
        self._materialise()
        return dict(self)

    def __getitem__(self, k):
        if k == "epsilon":
            self._materialise()
        return super().__getitem__(k)

    def get(self, k, default=None):
        if k == "epsilon":
            self._materialise()
        return super().get(k, default)

    def __contains__(self, k):
        return k == "epsilon" or super().__contains__(k)

    def keys(self):
        self._materialise()
        return super().keys()

    def items(self):
        self._materialise()
        return super().items()

    def values(self):
        self._materialise()
        return super().values()
