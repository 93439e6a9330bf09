"""FactorNetPPO — the coefficient policy of ConsistencySolver, as an nn.Module whose state_dict is
interchangeable with the reference's (`action_values`, `mlp.{0,2,4}.{weight,bias}`; factor_net_ppo.py:57-102 for
the SD variant, edit_ppo/factor_net_ppo.py:57-110 for the flow-matching variant).

Sampling (`sample_action`) runs the hand-written policy kernel (csrc/policy.cu) through the C ABI: the MLP +
softmax is evaluated once per step (the input row is identical for every sample, scheduler_ppo.py:207-210),
then every sample draws its bins as argmax(p/q) with q ~ Exp(1) taken from torch's default CUDA generator in
exactly the shape torch.multinomial would consume (factor_net_ppo.py:161), so seeds reproduce the reference's
actions.  The PPO-update side (`get_action_probs`) needs autograd and stays on torch ops.
There is no CPU path: tensors must live on a CUDA device."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


def _action_values(variant: str, K: int, order_dim: int, scaler_dim: int, mu_dim: int) -> torch.Tensor:
    first = torch.linspace(0, 2 if variant == "sd" else 1, K)
    second = torch.linspace(-2, 0, K)
    middle = torch.linspace(-1, 1, K)
    scaler = torch.linspace(-0.05, 0.05, K)
    A = order_dim + scaler_dim - 1 + (mu_dim if variant == "fm" else 0)
    rows = []
    for i in range(A):
        if i == 0:
            rows.append(first)
        elif i == 1 and (variant == "sd" or i < order_dim - 1):
            rows.append(second)
        elif i < order_dim - 1:
            rows.append(middle)
        elif variant == "sd" or i < order_dim + scaler_dim - 1:
            rows.append(scaler)
        else:
            rows.append(torch.cat((torch.tensor([0.0]), torch.linspace(0.5, 0.99, K - 1))))
    return torch.stack(rows)


class FactorNetPPO(nn.Module):
    """SD variant (factor_net_ppo.py:57-184).  Constructor kwargs as in the reference; `embedding_dim`,
    `input_channels`, `conv_out_channels` are accepted and ignored there too (:58-60 vs :70-81)."""

    variant = "sd"
    x_div = 999.0      # normalize_input: x.float() / 999.0   (factor_net_ppo.py:104-106)
    temperature = 1.0  # softmax(logits)                     (factor_net_ppo.py:156)

    def __init__(self, embedding_dim=1024, hidden_dim=256, num_actions=161, order_dim=4, scaler_dim=2,
                 use_conv=False, input_channels=4, conv_out_channels=8, mu_dim=0):
        super().__init__()
        if order_dim < 2 or order_dim > _lib.MAX_ORDER:
            raise ValueError(f"order_dim must be in [2, {_lib.MAX_ORDER}]")
        self.num_actions = num_actions
        self.order_dim = order_dim
        self.scaler_dim = scaler_dim
        self.mu_dim = mu_dim if self.variant == "fm" else 0
        self.action_dims = order_dim + scaler_dim + self.mu_dim - 1
        self.use_conv = use_conv
        self.hidden_dim = hidden_dim
        in_dim = 2 + ((order_dim - 1) if use_conv else 0)
        self.mlp = nn.Sequential(
            nn.Linear(in_dim, hidden_dim), nn.ReLU(),
            nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
            nn.Linear(hidden_dim, num_actions * self.action_dims),
        )
        if self.variant == "sd":  # uniform policy at start (factor_net_ppo.py:82-83); the FM variant keeps
            nn.init.zeros_(self.mlp[-1].bias)  # the default init (edit_ppo/factor_net_ppo.py:87-88)
            nn.init.zeros_(self.mlp[-1].weight)
        self.register_buffer("action_values",
                             _action_values(self.variant, num_actions, order_dim, scaler_dim, self.mu_dim))
        if hidden_dim > _lib.MAX_HIDDEN or num_actions * self.action_dims > _lib.MAX_LOGITS:
            raise ValueError("policy too large for the kernel limits in include/consolver.h")
        self._w32_cache = None
        self._kparams = None

    # ---- fp32 weight pointers for the kernel (the reference may cast the module to fp16, gen_ppo.py:193-195;
    #      the kernel always computes in fp32) -----------------------------------------------------------------
    def _params(self):
        # nn.Module attribute lookups cost ~1 us each: hold the six Parameter objects (stable across .to() /
        # load_state_dict, which rewrite .data in place) and re-fetch only the buffer (a new tensor after .to())
        mlp = self._modules["mlp"]
        if self._kparams is None or self._kparams[0] is not mlp:
            self._kparams = (mlp, [mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias, mlp[4].weight, mlp[4].bias])
        return self._kparams[1]

    def _kparams_dtype(self):
        return self._params()[0].dtype

    def kernel_weights(self, act_dtype: Optional[torch.dtype] = None):
        """Seven device pointers (w1 b1 w2 b2 w3 b3 action_values) to fp32 copies of the parameters.  `act_dtype`
        (fp16 / bf16): the MLP runs under autocast — the six Linear parameters are rounded to that dtype first, as
        autocast's cast of the operands does; the kernel then rounds the activations (CONSOLVER_POLICY_ACT_*)."""
        ps = self._params() + [self._buffers["action_values"]]
        key = (act_dtype,) + tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        c = self._w32_cache
        if c is None or c[0] != key:
            if not ps[0].is_cuda:
                raise RuntimeError("consolver_b200 has no CPU path: move factor_net to a CUDA device")

            def f32(p, lowp):
                t = p.detach()
                if lowp is not None and t.dtype != lowp:
                    t = t.to(lowp)
                return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()

            ws = [f32(p, act_dtype) for p in ps[:6]] + [f32(ps[6], None)]
            self._w32_cache = c = (key, ws, [w.data_ptr() for w in ws])
        return c[2]

    def normalize_input(self, x):
        return x.float() / 999.0 if self.variant == "sd" else x.float()

    def forward(self, x_dict, actions=None):
        if actions is None:
            return self.sample_action(x_dict)
        return self.get_action_probs(x_dict, actions)

    # ---- sampling side: CUDA kernel --------------------------------------------------------------------------
    def sample_action(self, x_dict: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """(actions [B,A], probs [B,A]) as factor_net_ppo.py:159-168.  `x_dict['x']` is [B,2] with identical
        rows (the scheduler's contract); use_conv additionally reads x_dict['epsilon']."""
        x = x_dict["x"]
        if not x.is_cuda:
            raise RuntimeError("consolver_b200 has no CPU path: x_dict['x'] must be a CUDA tensor")
        if self.use_conv:
            raise NotImplementedError("use_conv=True sampling goes through the scheduler (cosine features)")
        B = x.shape[0]
        A, K = self.action_dims, self.num_actions
        lib = _lib.load()
        dev = x.device
        # no host read-back of the row: the table kernel takes it from device memory, the sample kernel draws from the
        # table (two launches; the schedulers use the same pair with the tables of the whole grid evaluated once)
        act = None
        pflags = 0
        wdt = self._kparams_dtype()
        if wdt in (torch.float16, torch.bfloat16):
            act = wdt
        elif torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") in (torch.float16, torch.bfloat16):
            act = torch.get_autocast_dtype("cuda")
        if act is not None:
            pflags |= _lib.POLICY_ACT_F16 if act == torch.float16 else _lib.POLICY_ACT_BF16
        bdt = self.action_values.dtype
        if bdt in (torch.float16, torch.bfloat16):
            pflags |= _lib.POLICY_COEF_F16 if bdt == torch.float16 else _lib.POLICY_COEF_BF16
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            out = alloc_policy_outputs(B, A, K, self.order_dim, dev)
            table = out["probs_table"].view(1, A, K)
            self.policy_tables(x[:1].float().contiguous(), table, stream, policy_flags=pflags, act_dtype=act)
            q = torch.empty((B * A, K), device=dev, dtype=torch.float32).exponential_(1)   # torch.multinomial's own draw
            w = self.kernel_weights(act)
            rc = lib.consolver_policy_sample_f32(
                table.data_ptr(), w[6], q.data_ptr(), None, None, None, B, A, K, self.order_dim, self.scaler_dim,
                self.order_dim, pflags, out["idx"].data_ptr(), out["actions"].data_ptr(), out["probs"].data_ptr(),
                out["logp"].data_ptr(), out["masks"].data_ptr(), out["coef"].data_ptr(), stream)
            _lib.check(rc, "consolver_policy_sample_f32")
        self._last_sample = out
        actions = out["actions"] if bdt == torch.float32 else out["actions"].to(bdt)
        return actions, out["probs"]

    def policy_launch(self, x0: float, x1: float, B: int, n_hist: int, *, q: Optional[torch.Tensor] = None,
                      idx_in: Optional[torch.Tensor] = None, out: Optional[dict] = None, stream=None,
                      feat: Optional[torch.Tensor] = None, policy_flags: int = 0,
                      act_dtype: Optional[torch.dtype] = None):
        """Launch the policy kernel.  `q` None => drawn here from the default CUDA generator with the shape
        torch.multinomial consumes ([B*A, K] fp32).  `policy_flags`: CONSOLVER_POLICY_* (see include/consolver.h)."""
        lib = _lib.load()
        w = self.kernel_weights(act_dtype)
        dev = self.action_values.device
        A, K = self.action_dims, self.num_actions
        if q is None and idx_in is None:
            q = torch.empty((B * A, K), device=dev, dtype=torch.float32).exponential_(1)
        if out is None:
            out = alloc_policy_outputs(B, A, K, self.order_dim, dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.consolver_policy_f32(
            *w, x0, x1, self.x_div, self.temperature,
            feat.data_ptr() if feat is not None else None, feat.shape[1] if feat is not None else 0,
            q.data_ptr() if q is not None else None, idx_in.data_ptr() if idx_in is not None else None,
            B, self.hidden_dim, A, K, self.order_dim, self.scaler_dim, n_hist, policy_flags,
            out["probs_table"].data_ptr(), out["idx"].data_ptr(), out["actions"].data_ptr(),
            out["probs"].data_ptr(), out["logp"].data_ptr(), out["masks"].data_ptr(), out["coef"].data_ptr(),
            stream)
        _lib.check(rc, "consolver_policy_f32")
        return out

    def policy_tables(self, x_rows: torch.Tensor, out: torch.Tensor, stream=None, policy_flags: int = 0,
                      act_dtype: Optional[torch.dtype] = None):
        """Probability tables for many input rows in one launch: x_rows [R,2] fp32 (device) -> out [R,A,K]."""
        lib = _lib.load()
        w = self.kernel_weights(act_dtype)
        if stream is None:
            stream = torch.cuda.current_stream(x_rows.device).cuda_stream
        rc = lib.consolver_policy_table_f32(*w[:6], x_rows.data_ptr(), x_rows.shape[0], self.x_div, self.temperature,
                                            self.hidden_dim, self.action_dims, self.num_actions, policy_flags,
                                            out.data_ptr(), stream)
        _lib.check(rc, "consolver_policy_table_f32")
        return out

    # ---- PPO-update side: torch autograd (factor_net_ppo.py:137-157, :170-184) ----------------------------------
    def forward_(self, x_dict):
        x = self.normalize_input(x_dict["x"])
        if self.use_conv:
            from .features import cosine_features

            x = torch.cat([x, cosine_features(x_dict["epsilon"], self.order_dim)], dim=-1)
        logits = self.mlp(x).view(-1, self.action_dims, self.num_actions)
        if self.variant == "fm":
            logits = logits / 0.01
        return torch.softmax(logits, dim=-1)

    def get_action_probs(self, x_dict, actions):
        probs = self.forward_(x_dict)
        actions = actions.to(probs.device)
        idx = (actions.unsqueeze(-1) - self.action_values.unsqueeze(0)).abs().argmin(dim=-1)
        ent = torch.distributions.Categorical(probs=probs).entropy() / math.log(self.num_actions)
        return probs.gather(2, idx.unsqueeze(-1)).squeeze(-1), ent


class FactorNetPPOFM(FactorNetPPO):
    """Flow-matching variant (edit_ppo/factor_net_ppo.py:57-196): identity input normalisation (:112-114),
    softmax temperature 0.01 (:168), first bin row linspace(0,1) (:92), optional (unused) mu dims (:96,:109)."""

    variant = "fm"
    x_div = 1.0
    temperature = 0.01

    def __init__(self, embedding_dim=1024, hidden_dim=256, num_actions=161, order_dim=4, scaler_dim=2, mu_dim=1,
                 use_conv=False, input_channels=4, conv_out_channels=8):
        super().__init__(embedding_dim, hidden_dim, num_actions, order_dim, scaler_dim, use_conv,
                         input_channels, conv_out_channels, mu_dim=mu_dim)


class FactorNetPPOContinous(nn.Module):
    """CONTINUOUS (Gaussian) policy — `ppo_type != "discrete"`.  EXTENSION, PARITY UNPINNED.

    The reference imports and instantiates a class of this name (scheduler_ppo.py:23,:139 — the spelling is the
    reference's) but ships no source for it, so there is nothing to restate or to pin against: the semantics are this
    repo's own, documented in csrc/policy_gauss.cu and tested against closed forms (torch.distributions.Normal).  What IS
    the reference's — the (actions, probs) return convention of `sample_action`, the masks and the coefficient assembly
    of the scheduler — is shared with the discrete policy.

      trunk   Linear(2,H) ReLU Linear(H,H) ReLU Linear(H, 2A): raw mean / raw log-std per action dim
      mean_a  = mid_a + half_a * tanh(raw_mean_a);   std_a = half_a * exp(clamp(raw_logstd_a, -7, 1))
              [lo_a, hi_a] (buffer `action_range`) = the value range of the discrete policy's bins for that dim
      draw    actions = mean + std * z, z = ONE torch.randn([B, A]) per step from the default CUDA generator
      probs   exp(log N(action; mean, std)) — a density; the PPO loss takes log(probs + 1e-9) like train_ppo.py:410-411
    The last layer starts at zero with raw log-std `log_std_init` (mean = centre of the range)."""

    variant = "sd"
    x_div = 999.0
    temperature = 1.0
    num_actions = 2                  # the head has two outputs per action dim (sizes the shared trajectory buffers)
    continuous = True

    def __init__(self, embedding_dim=1024, hidden_dim=256, num_actions=None, order_dim=4, scaler_dim=2, use_conv=False,
                 input_channels=4, conv_out_channels=8, log_std_init=-1.0):
        super().__init__()
        if use_conv:
            raise NotImplementedError("use_conv is not defined for the continuous policy extension")
        if order_dim < 2 or order_dim > _lib.MAX_ORDER:
            raise ValueError(f"order_dim must be in [2, {_lib.MAX_ORDER}]")
        self.order_dim, self.scaler_dim, self.hidden_dim, self.use_conv = order_dim, scaler_dim, hidden_dim, False
        self.action_dims = A = order_dim + scaler_dim - 1
        self.mlp = nn.Sequential(nn.Linear(2, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
                                 nn.Linear(hidden_dim, 2 * A))
        nn.init.zeros_(self.mlp[-1].weight)
        with torch.no_grad():
            self.mlp[-1].bias[:A].zero_()
            self.mlp[-1].bias[A:].fill_(float(log_std_init))
        bins = _action_values("sd", 3, order_dim, scaler_dim, 0)                      # rows: [lo, mid, hi] of each dim
        self.register_buffer("action_range", torch.stack([bins[:, 0], bins[:, -1]], dim=1).contiguous())   # [A,2]
        self._w32_cache = None

    # the schedulers read these two like the discrete policy's
    @property
    def action_values(self):
        return self.action_range

    def _kparams_dtype(self):
        return self.mlp[0].weight.dtype

    def kernel_weights(self, act_dtype=None):
        ps = [self.mlp[0].weight, self.mlp[0].bias, self.mlp[2].weight, self.mlp[2].bias, self.mlp[4].weight,
              self.mlp[4].bias, self.action_range]
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        c = self._w32_cache
        if c is None or c[0] != key:
            if not ps[0].is_cuda:
                raise RuntimeError("consolver_b200 has no CPU path: move factor_net to a CUDA device")
            ws = [p.detach().float().contiguous() for p in ps]
            self._w32_cache = c = (key, ws, [w.data_ptr() for w in ws])
        return c[2]

    def mean_std(self, x):
        """torch (autograd) evaluation of the head: x [R,2] -> (mean [R,A], std [R,A])"""
        A = self.action_dims
        raw = self.mlp(x.float() / 999.0)
        lo, hi = self.action_range[:, 0], self.action_range[:, 1]
        mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
        return mid + half * torch.tanh(raw[:, :A]), half * torch.exp(raw[:, A:].clamp(-7.0, 1.0))

    def forward(self, x_dict, actions=None):
        if actions is None:
            return self.sample_action(x_dict)
        return self.get_action_probs(x_dict, actions)

    def policy_launch(self, x0, x1, B, n_hist, *, z=None, actions_in=None, rng=None, out=None, stream=None):
        import ctypes

        lib = _lib.load()
        w = self.kernel_weights()
        dev = self.action_range.device
        A = self.action_dims
        if z is None and actions_in is None and rng is None:
            z = torch.randn(B, A, device=dev)
        if out is None:
            out = alloc_policy_outputs(B, A, 2, self.order_dim, dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.consolver_policy_gauss_f32(
            *w, float(x0), float(x1), self.x_div, z.data_ptr() if z is not None else None,
            actions_in.data_ptr() if actions_in is not None else None, ctypes.byref(rng) if rng is not None else None,
            B, self.hidden_dim, A, self.order_dim, self.scaler_dim, n_hist, 0, out["probs_table"].data_ptr(), None,
            out["actions"].data_ptr(), out["probs"].data_ptr(), out["logp"].data_ptr(), out["masks"].data_ptr(),
            out["coef"].data_ptr(), stream)
        _lib.check(rc, "consolver_policy_gauss_f32")
        return out

    def sample_action(self, x_dict):
        """(actions [B,A], probs [B,A]) — the reference's return convention (factor_net_ppo.py:159-168)."""
        x = x_dict["x"]
        if not x.is_cuda:
            raise RuntimeError("consolver_b200 has no CPU path: x_dict['x'] must be a CUDA tensor")
        row = x[0].float().tolist()
        out = self.policy_launch(row[0], row[1], x.shape[0], n_hist=self.order_dim)
        return out["actions"], out["probs"]

    def get_action_probs(self, x_dict, actions):
        """PPO-update side (autograd): (density of `actions` under the current policy [R,A], normalised entropy [R,A])."""
        mean, std = self.mean_std(x_dict["x"])
        dist = torch.distributions.Normal(mean, std)
        return dist.log_prob(actions.to(mean.device)).exp(), dist.entropy()


def alloc_policy_outputs(B, A, K, order_dim, device, lead=()):
    """Buffers the policy kernel writes; `lead` prepends dims (the schedulers allocate [n_steps, ...] once per
    trajectory so the per-step results ARE the trajectory record — no unsqueeze/cat at the end)."""
    f = dict(device=device, dtype=torch.float32)
    return dict(
        probs_table=torch.empty(*lead, A, K, **f),
        idx=torch.empty(*lead, B, A, device=device, dtype=torch.int64),
        actions=torch.empty(*lead, B, A, **f),
        probs=torch.empty(*lead, B, A, **f),
        logp=torch.empty(*lead, B, A, **f),
        masks=torch.empty(*lead, B, A, **f),
        coef=torch.empty(*lead, B, order_dim + 2, **f),
    )
