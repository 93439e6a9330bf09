"""Baseline solvers as coefficient sources of the SAME fused step kernels, for apples-to-apples
speed tables against the learned solver: the only thing that changes is where the per-step multipliers of the
model-output history come from.

  DDIM (eta = 0)                 depth-1 history: the DDIM update of scheduler_ppo.py:306-332 on the newest estimate
  linear multistep (iPNDM / AB)  Adams-Bashforth weights on the eps history, warm-up with lower orders
  flow-matching Euler            x' = x + (sigma_next - sigma) v   (edit_ppo/scheduler_fm.py `type == "euler"`, :405-410)
  flow-matching AB multistep     Adams-Bashforth weights on the velocity history
"""
from __future__ import annotations

from typing import Sequence

from .scheduler_fmppo import FMPPOScheduler
from .scheduler_ppo import PPOScheduler

# Adams-Bashforth weights, newest first; each row sums to 1 (the constraint ConsistencySolver's learned
# coefficients obey by construction, scheduler_ppo.py:172)
ADAMS_BASHFORTH = {
    1: (1.0,),
    2: (3 / 2, -1 / 2),
    3: (23 / 12, -16 / 12, 5 / 12),
    4: (55 / 24, -59 / 24, 37 / 24, -9 / 24),
}


def _ab(order: int):
    def coef(n_hist: int) -> Sequence[float]:
        return ADAMS_BASHFORTH[min(n_hist, order)] + (0.0,) * max(0, n_hist - order)
    return coef


def _tiny_policy(kw):
    fk = dict(kw.pop("factor_net_kwargs", None) or {})
    fk.setdefault("hidden_dim", 8)
    fk.setdefault("num_actions", 3)
    kw["factor_net_kwargs"] = fk
    return kw


def ddim_solver(**scheduler_kwargs) -> PPOScheduler:
    """DDIM with eta = 0 through the fused kernel (CFG fusion via step_cfg included)."""
    kw = _tiny_policy(dict(scheduler_kwargs, order_dim=2, scaler_dim=0))
    s = PPOScheduler(**kw)
    s.fixed_coefficients = lambda n_hist: (1.0,) + (0.0,) * (n_hist - 1)
    s.fixed_depth = 1          # keep no history: every step is the depth-1 (bypass) form of the kernel
    return s


def multistep_solver(order: int = 4, **scheduler_kwargs) -> PPOScheduler:
    """Adams-Bashforth linear multistep in eps space with DDIM transfer (PLMS / iPNDM family)."""
    if order not in ADAMS_BASHFORTH:
        raise ValueError("order must be 1..4")
    kw = _tiny_policy(dict(scheduler_kwargs, order_dim=max(order, 2), scaler_dim=0))
    s = PPOScheduler(**kw)
    s.fixed_coefficients = _ab(order)
    s.fixed_depth = order
    return s


def flow_euler_solver(**scheduler_kwargs) -> FMPPOScheduler:
    kw = _tiny_policy(dict(scheduler_kwargs, order_dim=2, scaler_dim=0, mu_dim=0))
    s = FMPPOScheduler(**kw)
    s.fixed_coefficients = lambda n_hist: (1.0,) + (0.0,) * (n_hist - 1)
    s.fixed_depth = 1
    return s


def flow_multistep_solver(order: int = 2, **scheduler_kwargs) -> FMPPOScheduler:
    if order not in ADAMS_BASHFORTH:
        raise ValueError("order must be 1..4")
    kw = _tiny_policy(dict(scheduler_kwargs, order_dim=max(order, 2), scaler_dim=0, mu_dim=0))
    s = FMPPOScheduler(**kw)
    s.fixed_coefficients = _ab(order)
    s.fixed_depth = order
    return s
