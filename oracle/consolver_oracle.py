"""TEST INFRASTRUCTURE ONLY — CPU oracle for the ConsistencySolver sampling step.

A CPU restatement (torch-CPU fp32 tensor ops, op-for-op in the reference's rounding order) of the
hot path of G-U-N/consolver.  The reference is itself torch code, so torch-CPU is the faithful
arithmetic; every function cites the reference file:line it follows (paths relative to the
reference root).  This module is the CHECKER for the CUDA kernels:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
    legs may import it.  The product package `consolver_b200` never does and has no CPU fallback.
  * PARITY PIN: the reference ships no tests / golden vectors for this path, so the
    oracle is pinned against outputs of the unmodified reference itself, produced in the build
    container by `oracle/make_golden.py` (through `oracle/ref_shim.py`) and committed as
    `tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks this module against every fixture.
  * UNPINNED: the continuous/Gaussian policy (`ppo_type != "discrete"`): its source
    (`factor_net_ppo_continous`) is absent from the reference, nothing exists to pin against, so it
    is not restated here.

Layout conventions: a "latent" is one sample's tensor (SD 4x64x64, FM 4096x64); history is a list of
model outputs NEWEST FIRST; `coef[b, i]` multiplies history entry i (i=0 newest).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

F32 = torch.float32


# --------------------------------------------------------------------------------------------
# where the reference runs: torch evaluates `0-d HOST scalar (op) tensor` differently for CPU and CUDA tensors
# --------------------------------------------------------------------------------------------
class TorchSemantics:
    """The reference keeps its schedule scalars as 0-d fp32 tensors ON THE HOST (scheduler_ppo.py:110-114,:309-312)
    while latents and model outputs live on the GPU.  ATen treats such a scalar differently per device:

      device="cpu"   (the reference executed on CPU tensors — what the CPU-made fixtures tests/golden/{sd,sd16,fm}_* hold)
          scalar * t16   -> the scalar is first rounded to t16's dtype, the product rounded again
          t / scalar     -> true division
      device="cuda"  (the reference executed the way its drivers run it — fixtures tests/golden/cuda_*, made on a B200)
          scalar * t     -> the scalar enters the kernel as an fp32 opmath value: round_T(float(t) * k), one rounding
                            (ATen opmath_symmetric_gpu_kernel_with_scalars)
          t / scalar     -> t * (1 / scalar), the reciprocal taken once in fp32 (ATen div_true_kernel_cuda's
                            CPU-scalar branch) — this changes fp32 results too, by an ulp here and there

    `autocast` (None / torch.float16 / torch.bfloat16): the step runs inside torch.autocast("cuda", dtype)
    (gen_ppo.py:309, train_ppo.py:353): nn.Linear runs in that dtype, softmax and torch.sum return fp32.
    Everything here is torch-CPU arithmetic that restates those rules; tests/test_oracle_golden.py pins the
    restatement against the cuda_* fixtures bit for bit."""

    def __init__(self, device: str = "cpu", autocast: Optional[torch.dtype] = None):
        if device not in ("cpu", "cuda"):
            raise ValueError("device must be 'cpu' or 'cuda'")
        self.device, self.autocast = device, autocast

    @property
    def cuda(self) -> bool:
        return self.device == "cuda"

    def smul(self, k, t: torch.Tensor) -> torch.Tensor:
        """0-d host scalar * tensor"""
        if not self.cuda:
            return k * t
        return (t.float() * torch.as_tensor(k, dtype=F32)).to(t.dtype)

    def sdiv(self, t: torch.Tensor, k) -> torch.Tensor:
        """tensor / 0-d host scalar (or python number)"""
        if not self.cuda:
            return t / k
        inv = torch.tensor(1.0, dtype=F32) / torch.as_tensor(k, dtype=F32)
        return (t.float() * inv).to(t.dtype)

    def sum0(self, stacked: torch.Tensor) -> torch.Tensor:
        """torch.sum(stacked, dim=0) of set_default_coefficients (scheduler_ppo.py:172) — fp32 in, fp32 out under
        autocast.  fp32 addition is not associative and ATen's CUDA reduce kernel (ATen/native/cuda/Reduce.cuh) does not
        add left to right:
          * B >= 2 per-sample values (the batch is the fastest dimension): each thread reduces its own sample with
            vt0 = 4 accumulators, acc[i % 4] += s_i, then ((acc0 + acc1) + acc2) + acc3 — left to right up to 4 terms;
          * B == 1 (the reduced dimension is the fastest one): block.x = last_pow2(m) threads share the sum: thread x
            takes s_x, s_{x+W}, ... into its accumulators, then a shuffle tree with decreasing offsets —
            three terms give (s0 + s2) + s1.
        Pinned by the B = 1 / order_dim 6 and 8 fixtures made on the B200 and by the live differential tests."""
        if self.autocast is not None and stacked.dtype in (torch.float16, torch.bfloat16):
            stacked = stacked.float()
        if not self.cuda or stacked.dtype != F32:
            return torch.sum(stacked, dim=0)
        m = stacked.shape[0]
        zero = torch.zeros_like(stacked[0])

        def four(vals):
            acc = [zero, zero, zero, zero]
            for k, v in enumerate(vals):
                acc[k % 4] = acc[k % 4] + v
            return ((acc[0] + acc[1]) + acc[2]) + acc[3]

        if stacked[0].numel() != 1:
            return four([stacked[i] for i in range(m)])
        W = 1
        while W * 2 <= m:
            W *= 2
        t = [four([stacked[i] for i in range(x, m, W)]) for x in range(W)]
        off = W // 2
        while off > 0:
            for x in range(off):
                t[x] = t[x] + t[x + off]
            off //= 2
        return t[0]


HOST = TorchSemantics("cpu")
CUDA = TorchSemantics("cuda")


# --------------------------------------------------------------------------------------------
# schedules
# --------------------------------------------------------------------------------------------
def sd_betas(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
             trained_betas=None) -> torch.Tensor:
    """scheduler_ppo.py:99-108 (+ betas_for_alpha_bar :25-45)."""
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=F32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=F32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=F32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        def abar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        b = [min(1 - abar((i + 1) / num_train_timesteps) / abar(i / num_train_timesteps), 0.999)
             for i in range(num_train_timesteps)]
        return torch.tensor(b, dtype=F32)
    raise NotImplementedError(f"{beta_schedule} schedule not implemented.")


def sd_alphas_cumprod(betas: torch.Tensor) -> torch.Tensor:
    """scheduler_ppo.py:110-111; fp32 cumprod."""
    return torch.cumprod(1.0 - betas, dim=0)


def sd_timesteps(num_inference_steps: int, num_train_timesteps=1000, timestep_spacing="leading",
                 steps_offset=0) -> np.ndarray:
    """scheduler_ppo.py:142-163."""
    n, T = num_inference_steps, num_train_timesteps
    if n > T:
        raise ValueError("num_inference_steps > num_train_timesteps")
    if timestep_spacing == "linspace":
        return np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
    if timestep_spacing == "leading":
        ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64)
        return ts + steps_offset
    if timestep_spacing == "trailing":
        return np.round(np.arange(T, 0, -(T / n))).astype(np.int64) - 1
    raise ValueError(f"Unsupported timestep_spacing: {timestep_spacing}.")


def sd_prev_timestep(t: int, num_inference_steps: int, num_train_timesteps=1000) -> int:
    """scheduler_ppo.py:203 — NOT the next grid point: t - (T // n)."""
    return int(t) - num_train_timesteps // num_inference_steps


def ddim_scalars(alphas_cumprod: torch.Tensor, t: int, prev_t: int) -> Tuple[torch.Tensor, ...]:
    """scheduler_ppo.py:309-312 and the `** 0.5` terms of :317,:323,:329-330, as fp32 0-d tensors:
    (sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)); prev_t<0 uses alphas_cumprod[0] (:114,:310)."""
    a_t = alphas_cumprod[int(t)]
    a_p = alphas_cumprod[int(prev_t)] if prev_t >= 0 else alphas_cumprod[0]
    b_t = 1 - a_t
    b_p = 1 - a_p
    return a_t ** 0.5, b_t ** 0.5, a_p ** 0.5, b_p ** 0.5


def fm_sigmas(num_inference_steps=None, sigmas=None, mu=None, timesteps=None, *, num_train_timesteps=1000,
              shift=1.0, use_dynamic_shifting=False, time_shift_type="exponential", shift_terminal=None,
              invert_sigmas=False, use_karras_sigmas=False, use_exponential_sigmas=False,
              sigma_min=None, sigma_max=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """edit_ppo/scheduler_fmppo.py:171-245 (+:142-151 for the ctor's default grid, :489-499, :516-530,
    :546-550).  Returns (timesteps[n], sigmas[n+1]) fp32.  Beta sigmas (scipy) are not restated: no shipped
    config uses them."""
    if use_dynamic_shifting and mu is None:
        raise ValueError("`mu` must be passed when `use_dynamic_shifting` is set to be `True`")
    if sigmas is not None and timesteps is not None and len(sigmas) != len(timesteps):
        raise ValueError("`sigmas` and `timesteps` should have the same length")
    if num_inference_steps is not None:
        if (sigmas is not None and len(sigmas) != num_inference_steps) or (
                timesteps is not None and len(timesteps) != num_inference_steps):
            raise ValueError("`sigmas`/`timesteps` length must equal num_inference_steps")
    else:
        num_inference_steps = len(sigmas) if sigmas is not None else len(timesteps)
    ts_given = timesteps is not None
    if ts_given:
        timesteps = np.array(timesteps).astype(np.float32)
    if sigmas is None:
        if timesteps is None:
            # ctor grid (:142-151): sigma_max/min come from the constructor's shifted 1000-point grid
            t0 = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
            s0 = torch.from_numpy(t0).to(F32) / num_train_timesteps
            if not use_dynamic_shifting:
                s0 = shift * s0 / (1 + (shift - 1) * s0)
            smax = s0[0].item() if sigma_max is None else sigma_max
            smin = s0[-1].item() if sigma_min is None else sigma_min
            timesteps = np.linspace(smax * num_train_timesteps, smin * num_train_timesteps, num_inference_steps)
        sig = timesteps / num_train_timesteps
    else:
        sig = np.array(sigmas).astype(np.float32)
        num_inference_steps = len(sig)
    if use_dynamic_shifting:
        if time_shift_type == "exponential":
            sig = math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)
        else:
            sig = mu / (mu + (1 / sig - 1) ** 1.0)
    else:
        sig = shift * sig / (1 + (shift - 1) * sig)
    if shift_terminal:
        omz = 1 - sig
        sig = 1 - omz / (omz[-1] / (1 - shift_terminal))
    if use_karras_sigmas:
        rho, ramp = 7.0, np.linspace(0, 1, num_inference_steps)
        lo, hi = sig[-1].item() ** (1 / rho), sig[0].item() ** (1 / rho)
        sig = (hi + ramp * (lo - hi)) ** rho
    elif use_exponential_sigmas:
        sig = np.exp(np.linspace(math.log(sig[0].item()), math.log(sig[-1].item()), num_inference_steps))
    sig_t = torch.from_numpy(np.asarray(sig)).to(dtype=F32)
    ts_t = torch.from_numpy(timesteps).to(dtype=F32) if ts_given else sig_t * num_train_timesteps
    if invert_sigmas:
        sig_t = 1.0 - sig_t
        ts_t = sig_t * num_train_timesteps
        sig_t = torch.cat([sig_t, torch.ones(1)])
    else:
        sig_t = torch.cat([sig_t, torch.zeros(1)])
    return ts_t, sig_t


# --------------------------------------------------------------------------------------------
# policy (FactorNetPPO)
# --------------------------------------------------------------------------------------------
def action_dims(variant: str, order_dim: int, scaler_dim: int, mu_dim: int = 0) -> int:
    """factor_net_ppo.py:66 (sd) / edit_ppo/factor_net_ppo.py:67 (fm)."""
    return order_dim + scaler_dim - 1 + (mu_dim if variant == "fm" else 0)


def action_value_table(variant: str, num_actions: int, order_dim: int, scaler_dim: int, mu_dim: int = 0) -> torch.Tensor:
    """Bin values [A, K]: factor_net_ppo.py:87-102 (sd) / edit_ppo/factor_net_ppo.py:92-110 (fm)."""
    K = num_actions
    A = action_dims(variant, order_dim, scaler_dim, mu_dim)
    rows = []
    for i in range(A):
        if i == 0:
            rows.append(torch.linspace(0, 2 if variant == "sd" else 1, K))
        elif i == 1 and (variant == "sd" or i < order_dim - 1):
            rows.append(torch.linspace(-2, 0, K))
        elif i < order_dim - 1:
            rows.append(torch.linspace(-1, 1, K))
        elif variant == "sd" or i < order_dim + scaler_dim - 1:
            rows.append(torch.linspace(-0.05, 0.05, K))
        else:
            rows.append(torch.cat((torch.tensor([0.0]), torch.linspace(0.5, 0.99, K - 1))))
    return torch.stack(rows)


def cosine_features(epsilon: torch.Tensor, order_dim: int) -> torch.Tensor:
    """factor_net_ppo.py:108-130 — cos-sim of history slot 0 vs slots 1..order_dim-1 (use_conv=True only)."""
    B = epsilon.shape[0]
    flat = epsilon.reshape(B, order_dim, -1)
    first = flat[:, 0, :]
    return torch.cat([torch.nn.functional.cosine_similarity(flat[:, i, :], first, dim=-1).unsqueeze(-1)
                      for i in range(1, order_dim)], dim=-1)


def _linear_lowp(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """nn.Linear under autocast: operands cast to `dt`, fp32 accumulation, bias added before the single rounding of
    the output (cuBLAS; the accumulation order is the library's, so a logit can differ from this by one `dt` ulp)."""
    acc = x.to(dt).double() @ w.to(dt).double().t() + b.to(dt).double()
    return acc.float().to(dt)


def policy_probs(sd: Dict[str, torch.Tensor], x: torch.Tensor, variant: str,
                 epsilon: Optional[torch.Tensor] = None, order_dim: int = 4,
                 sem: Optional["TorchSemantics"] = None) -> torch.Tensor:
    """forward_: factor_net_ppo.py:137-157 (sd: x/999, softmax) / edit_ppo/factor_net_ppo.py:149-169
    (fm: identity normalise, softmax(logits/0.01)).  x: [R, 2] in the model dtype; returns [R, A, K] fp32.
    `epsilon` (stacked history [R, order_dim, ...]) switches on the use_conv feature path.
    `sem.autocast`: the three Linear layers run in that dtype (activations rounded to it after every layer), the
    softmax in fp32 on the upcast logits; a policy whose parameters are already 16-bit (gen_ppo.py:194-195) is the
    same computation."""
    sem = sem or HOST
    av = sd["action_values"]
    A, K = av.shape
    xn = sem.sdiv(x.float(), 999.0) if variant == "sd" else x.float()
    if epsilon is not None:
        xn = torch.cat([xn, cosine_features(epsilon, order_dim)], dim=-1)
    dt = sem.autocast
    if dt is None:
        h = torch.relu(torch.nn.functional.linear(xn, sd["mlp.0.weight"].float(), sd["mlp.0.bias"].float()))
        h = torch.relu(torch.nn.functional.linear(h, sd["mlp.2.weight"].float(), sd["mlp.2.bias"].float()))
        logits = torch.nn.functional.linear(h, sd["mlp.4.weight"].float(), sd["mlp.4.bias"].float()).view(-1, A, K)
    else:
        h = torch.relu(_linear_lowp(xn, sd["mlp.0.weight"], sd["mlp.0.bias"], dt))
        h = torch.relu(_linear_lowp(h, sd["mlp.2.weight"], sd["mlp.2.bias"], dt))
        logits = _linear_lowp(h, sd["mlp.4.weight"], sd["mlp.4.bias"], dt).view(-1, A, K)
    if variant == "fm":
        logits = sem.sdiv(logits, 0.01)
    return torch.softmax(logits.float(), dim=-1)


def sample_indices(probs: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """factor_net_ppo.py:161 — torch.multinomial(p.view(-1,K), 1) is argmax(p / q), q ~ Exp(1) drawn as
    `empty_like(p).exponential_(1)` from the default generator (ATen multinomial n_sample==1 fast path).
    probs [B, A, K], q [B*A, K] -> idx [B, A] int64."""
    B, A, K = probs.shape
    return torch.argmax(probs.reshape(-1, K) / q, dim=-1).view(B, A)


def draw_q(B: int, A: int, K: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """The RNG contract of the path: one exponential_ draw of shape [B*A, K] fp32 per step."""
    return torch.empty(B * A, K, dtype=F32).exponential_(1, generator=generator)


def gather_actions(sd: Dict[str, torch.Tensor], probs: torch.Tensor, idx: torch.Tensor):
    """factor_net_ppo.py:164-168: (bin values, probabilities of the sampled bins), both [B, A]."""
    B = idx.shape[0]
    av = sd["action_values"].unsqueeze(0).expand(B, -1, -1)
    return (torch.gather(av, 2, idx.unsqueeze(-1)).squeeze(-1),
            probs.gather(2, idx.unsqueeze(-1)).squeeze(-1))


def action_indices_from_values(sd, actions: torch.Tensor) -> torch.Tensor:
    """factor_net_ppo.py:174-178 — nearest-bin recovery on the PPO-update side."""
    av = sd["action_values"]
    idx = torch.zeros_like(actions, dtype=torch.long)
    for d in range(av.shape[0]):
        idx[:, d] = (actions[:, d].unsqueeze(-1) - av[d]).abs().argmin(dim=-1)
    return idx


def action_probs_entropy(sd, x: torch.Tensor, actions: torch.Tensor, variant: str):
    """get_action_probs: factor_net_ppo.py:170-184 -> (selected probs [R,A], normalised entropy [R,A])."""
    probs = policy_probs(sd, x, variant)
    idx = action_indices_from_values(sd, actions)
    # reference divides by torch.log(tensor(K, dtype)) (fp32); do the same to stay bit-identical
    ent = torch.distributions.Categorical(probs=probs).entropy() / torch.log(
        torch.as_tensor(probs.shape[2], dtype=probs.dtype))
    return probs.gather(2, idx.unsqueeze(-1)).squeeze(-1), ent


def step_masks(B: int, A: int, n_hist: int, order_dim: int) -> torch.Tensor:
    """scheduler_ppo.py:248-249."""
    m = torch.ones(B, A, dtype=F32)
    m[:, n_hist - 1:order_dim - 1] = 0
    return m


def coefficients(actions: torch.Tensor, n_hist: int, order_dim: int, scaler_dim: int,
                 sem: Optional["TorchSemantics"] = None):
    """scheduler_ppo.py:253-259 + set_default_coefficients :165-175 (same in edit_ppo/scheduler_fmppo.py
    :249-268).  Returns (coef list of n_hist tensors [B] newest first — for n_hist==1 the reference
    bypasses the coefficient and uses the estimate itself (:263-265) so coef is None —, scale list).
    With 16-bit actions (policy cast to the pipeline dtype) under autocast, torch.sum returns fp32, so the LAST
    coefficient is an fp32 tensor while the others keep the 16-bit dtype."""
    sem = sem or HOST
    a = [actions[:, i] for i in range(order_dim - 1)]
    s = [actions[:, i] for i in range(order_dim - 1, order_dim + scaler_dim - 1)]
    a.append(a[-1] if a else None)  # placeholder (:166)
    a[0] = a[0] + 1
    if n_hist > 1:
        a[n_hist - 1] = 1 - sem.sum0(torch.stack(a[:n_hist - 1]))
    s = [v + 1 for v in s]
    return (a[:n_hist] if n_hist > 1 else None), s


# --------------------------------------------------------------------------------------------
# per-element update (what the fused CUDA step kernel computes)
# --------------------------------------------------------------------------------------------
def cfg_combine(uncond: torch.Tensor, cond: torch.Tensor, guidance: float) -> torch.Tensor:
    """denoise_ppo.py:97-100: u + g*(c - u)."""
    return uncond + guidance * (cond - uncond)


def _bc(v: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return v.view(-1, *([1] * (like.dim() - 1)))


def combine_history(hist: Sequence[torch.Tensor], coef, scale, sample: torch.Tensor):
    """scheduler_ppo.py:263-280: eff = sum_i c_i * e_i (python sum, left to right from 0), then the
    optional (1+s0) / (1+s1) scalings.  Returns (eff, sample')."""
    if coef is None:
        eff = hist[0]
    else:
        eff = sum(_bc(c, e) * e for c, e in zip(coef, hist))
    if len(scale) >= 1:
        eff = eff * _bc(scale[0], eff)
    if len(scale) == 2:
        sample = sample * _bc(scale[1], sample)
    elif len(scale) > 2:
        raise NotImplementedError
    return eff, sample


def ddim_update(sample, eff, scalars, prediction_type="epsilon", sem: Optional["TorchSemantics"] = None):
    """scheduler_ppo.py:306-332 (_get_prev_sample), eta=0, no clipping.  The four scalars are 0-d HOST tensors in the
    reference: `sem` says how torch combines them with the (CPU or CUDA) latents."""
    sem = sem or HOST
    sa_t, sb_t, sa_p, sb_p = scalars
    if prediction_type == "v_prediction":
        eff = sem.smul(sa_t, eff) + sem.smul(sb_t, sample)
    elif prediction_type != "epsilon":
        raise ValueError(f"Unsupported prediction_type: {prediction_type}")
    x0 = sem.sdiv(sample - sem.smul(sb_t, eff), sa_t)
    return sem.smul(sa_p, x0) + sem.smul(sb_p, eff)


def fm_update(sample, eff, dt, out_dtype):
    """edit_ppo/scheduler_fmppo.py:354,:429,:436: fp32 sample + dt*eff, cast to the model dtype."""
    return (sample + dt * eff).to(out_dtype)


# --------------------------------------------------------------------------------------------
# stateful runners with the reference's step() contract (used by tests and the CPU baseline)
# --------------------------------------------------------------------------------------------
class OracleSDScheduler:
    """PPOScheduler restated (scheduler_ppo.py:81-332).  `step()` returns the reference's 5-tuple.
    `forced_idx` (optional [B, A] int64) replaces the sampled indices (for latent parity with injected
    actions); `q` (optional) replaces the default-generator draw."""

    def __init__(self, state_dict, *, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02,
                 beta_schedule="linear", trained_betas=None, prediction_type="epsilon",
                 timestep_spacing="leading", steps_offset=0, order_dim=4, scaler_dim=2, use_conv=False,
                 sem: Optional[TorchSemantics] = None, policy_dtype: Optional[torch.dtype] = None):
        """`sem`: where the reference being restated ran (TorchSemantics; default: CPU tensors, no autocast).
        `policy_dtype`: the policy — bin buffer included — was cast to this dtype (gen_ppo.py:194-195)."""
        self.sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
        if policy_dtype is not None:
            self.sd = {k: v.to(policy_dtype) for k, v in self.sd.items()}
        self.sem = sem or HOST
        self.T, self.order_dim, self.scaler_dim = num_train_timesteps, order_dim, scaler_dim
        self.prediction_type, self.spacing, self.offset = prediction_type, timestep_spacing, steps_offset
        self.use_conv = use_conv
        self.alphas_cumprod = sd_alphas_cumprod(
            sd_betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas))
        self.n = None
        self.hist: List[torch.Tensor] = []

    def set_timesteps(self, n: int):
        self.timesteps = torch.from_numpy(sd_timesteps(n, self.T, self.spacing, self.offset))
        self.n = n
        self.hist = []

    def step(self, model_output, timestep, sample, q=None, forced_idx=None):
        if self.n is None:
            raise ValueError("Number of inference steps is 'None'. Call 'set_timesteps' first.")
        t = int(timestep)
        prev_t = sd_prev_timestep(t, self.n, self.T)
        B = model_output.shape[0]
        x = torch.tensor([[t, prev_t]], dtype=model_output.dtype).repeat(B, 1)
        self.hist = ([model_output] + self.hist)[: self.order_dim]      # newest first
        n_hist = len(self.hist)
        eps_stack = torch.stack(self.hist, dim=1)     # :222-232, built (and paid for) on every step
        if n_hist < self.order_dim:
            pad = torch.zeros(B, self.order_dim - n_hist, *model_output.shape[1:], dtype=model_output.dtype)
            eps_stack = torch.cat([eps_stack, pad], dim=1)
        probs = policy_probs(self.sd, x, "sd", eps_stack if self.use_conv else None, self.order_dim, sem=self.sem)
        A, K = probs.shape[1:]
        if forced_idx is None:
            if q is None:
                q = draw_q(B, A, K)
            idx = sample_indices(probs, q)
        else:
            idx = forced_idx
        actions, act_probs = gather_actions(self.sd, probs, idx)
        masks = step_masks(B, A, n_hist, self.order_dim)
        coef, scale = coefficients(actions, n_hist, self.order_dim, self.scaler_dim, sem=self.sem)
        eff, smp = combine_history(self.hist, coef, scale, sample)
        prev = ddim_update(smp, eff, ddim_scalars(self.alphas_cumprod, t, prev_t), self.prediction_type, sem=self.sem)
        self.last_idx, self.last_probs_full = idx, probs
        return prev, actions, act_probs, {"x": x, "epsilon": eps_stack}, masks


class OracleFMScheduler:
    """FMPPOScheduler restated (edit_ppo/scheduler_fmppo.py:108-455), per_token_timesteps excluded."""

    def __init__(self, state_dict, *, num_train_timesteps=1000, shift=1.0, use_dynamic_shifting=False,
                 time_shift_type="exponential", shift_terminal=None, invert_sigmas=False,
                 use_karras_sigmas=False, use_exponential_sigmas=False, order_dim=4, scaler_dim=2, mu_dim=1,
                 use_conv=False, sem: Optional[TorchSemantics] = None):
        self.sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
        self.sem = sem or HOST
        self.kw = dict(num_train_timesteps=num_train_timesteps, shift=shift,
                       use_dynamic_shifting=use_dynamic_shifting, time_shift_type=time_shift_type,
                       shift_terminal=shift_terminal, invert_sigmas=invert_sigmas,
                       use_karras_sigmas=use_karras_sigmas, use_exponential_sigmas=use_exponential_sigmas)
        self.order_dim, self.scaler_dim, self.mu_dim, self.use_conv = order_dim, scaler_dim, mu_dim, use_conv
        self.n = None
        self.hist: List[torch.Tensor] = []

    def set_timesteps(self, num_inference_steps=None, sigmas=None, mu=None, timesteps=None):
        self.timesteps, self.sigmas = fm_sigmas(num_inference_steps, sigmas, mu, timesteps, **self.kw)
        self.n = len(self.timesteps)
        self.step_index = None
        self.begin_index = None
        self.hist = []

    def set_begin_index(self, i=0):
        self.begin_index = i

    def step(self, model_output, timestep, sample, q=None, forced_idx=None):
        if self.n is None:
            raise ValueError("Number of inference steps is 'None'. Call 'set_timesteps' first.")
        if isinstance(timestep, int) or (torch.is_tensor(timestep) and not timestep.is_floating_point()):
            raise ValueError("Passing integer indices as timesteps to `step()` is not supported.")
        if self.step_index is None:
            if self.begin_index is None:
                hits = (self.timesteps == timestep).nonzero()
                self.step_index = hits[1 if len(hits) > 1 else 0].item()
            else:
                self.step_index = self.begin_index
        sample = sample.to(F32)
        self.hist = ([model_output] + self.hist)[: self.order_dim]
        n_hist = len(self.hist)
        cur, nxt = self.sigmas[self.step_index], self.sigmas[self.step_index + 1]
        dt = nxt - cur
        B = model_output.shape[0]
        x = torch.tensor([[cur, nxt]], dtype=model_output.dtype).repeat(B, 1)
        eps_stack = torch.stack(self.hist, dim=1)
        if n_hist < self.order_dim:
            pad = torch.zeros(B, self.order_dim - n_hist, *model_output.shape[1:], dtype=model_output.dtype)
            eps_stack = torch.cat([eps_stack, pad], dim=1)
        probs = policy_probs(self.sd, x, "fm", eps_stack if self.use_conv else None, self.order_dim, sem=self.sem)
        A, K = probs.shape[1:]
        if forced_idx is None:
            if q is None:
                q = draw_q(B, A, K)
            idx = sample_indices(probs, q)
        else:
            idx = forced_idx
        actions, act_probs = gather_actions(self.sd, probs, idx)
        masks = step_masks(B, A, n_hist, self.order_dim)
        coef, scale = coefficients(actions, n_hist, self.order_dim, self.scaler_dim, sem=self.sem)  # mu params unused (:409,:440)
        eff, smp = combine_history(self.hist, coef, scale, sample)
        prev = fm_update(smp, eff, dt, model_output.dtype)
        self.step_index += 1
        self.last_idx, self.last_probs_full = idx, probs
        return prev, actions, act_probs, {"x": x, "epsilon": eps_stack}, masks


class OracleFMGeneralScheduler:
    """The reference's training-free flow-matching baselines restated:
    FlowMatchGeneralDiscreteScheduler.step, edit_ppo/scheduler_fm.py:384-488, `type` in
    euler / heun / dpm-solver / dpm-solver-multistep.  The sigma schedule is the same code as FMPPOScheduler's
    (edit_ppo/scheduler_fm.py:259-353).  The two-stage kinds keep the first stage's fp32 sample, model output and
    step size between calls; that state is NOT cleared by set_timesteps in the reference (its one-line
    set_timesteps override at :141-145 is shadowed by the full definition at :259), and is not cleared here."""

    KINDS = ("euler", "heun", "dpm-solver", "dpm-solver-multistep")

    def __init__(self, *, kind="euler", num_train_timesteps=1000, shift=1.0, use_dynamic_shifting=False,
                 time_shift_type="exponential", shift_terminal=None, invert_sigmas=False,
                 use_karras_sigmas=False, use_exponential_sigmas=False):
        self.kw = dict(num_train_timesteps=num_train_timesteps, shift=shift,
                       use_dynamic_shifting=use_dynamic_shifting, time_shift_type=time_shift_type,
                       shift_terminal=shift_terminal, invert_sigmas=invert_sigmas,
                       use_karras_sigmas=use_karras_sigmas, use_exponential_sigmas=use_exponential_sigmas)
        self.kind = kind
        self.n = None
        self.prev_dt = self.prev_sample = self.prev_model_output = None

    def set_timesteps(self, num_inference_steps=None, sigmas=None, mu=None, timesteps=None):
        self.timesteps, self.sigmas = fm_sigmas(num_inference_steps, sigmas, mu, timesteps, **self.kw)
        self.n = len(self.timesteps)
        self.step_index = None
        self.begin_index = None

    def set_begin_index(self, i=0):
        self.begin_index = i

    def step(self, model_output, timestep, sample):
        if self.step_index is None:
            if self.begin_index is None:
                hits = (self.timesteps == timestep).nonzero()
                self.step_index = hits[1 if len(hits) > 1 else 0].item()
            else:
                self.step_index = self.begin_index
        sample = sample.to(F32)                                           # :399
        i, sg = self.step_index, self.sigmas
        if self.kind == "euler":                                          # :405-410
            nxt = sg[i + 1] if i + 1 < len(sg) else sg[-1]
            prev = sample + (nxt - sg[i]) * model_output
        elif self.kind == "heun":                                         # :412-430
            if i % 2 == 0:
                nxt = sg[i + 2] if i + 2 < len(sg) else sg[-1]
                self.prev_dt, self.prev_sample, self.prev_model_output = nxt - sg[i], sample, model_output
                prev = sample + self.prev_dt * model_output
            else:
                prev = self.prev_sample + 0.5 * self.prev_dt * (self.prev_model_output + model_output)
        elif self.kind == "dpm-solver":                                   # :431-452
            if i % 2 == 0:
                self.prev_dt, self.prev_sample, self.prev_model_output = sg[i + 1] - sg[i], sample, model_output
                prev = sample + self.prev_dt * model_output
            else:
                prev = self.prev_sample + (self.prev_dt + (sg[i + 1] - sg[i])) * model_output
        elif self.kind == "dpm-solver-multistep":                         # :454-483
            if i == 0:
                self.prev_dt, self.prev_sample, self.prev_model_output = sg[1] - sg[0], sample, model_output
                prev = sample + self.prev_dt * model_output
            else:
                prev = self.prev_sample + (self.prev_dt + (sg[i + 1] - sg[i])) * model_output
                self.prev_dt, self.prev_sample = sg[i + 1] - sg[i], sample
        else:
            # the reference falls through with `prev_sample` unbound (UnboundLocalError); a clear error here
            raise ValueError(f"unknown solver type {self.kind!r}")
        self.step_index += 1
        return prev.to(model_output.dtype)                                # :488


class OracleDPMSolverAMED:
    """The AMED baseline restated: the reference's plugin `diffusers_amed_plugin_dpmpp.py`
    (custom-timestep set_timesteps :29-68, first-order update :70-138, second-order update :140-262, step :350-436,
    `scale_dir` use :417-423) on top of diffusers 0.26.3 `DPMSolverMultistepScheduler`.  diffusers is a third-party
    dependency ABSENT from the reference tree and from this image (pinned `diffusers==0.26.3`, env.yaml:52); the
    inherited pieces (beta tables, `_sigma_to_alpha_sigma_t`, `convert_model_output`, stock `set_timesteps`) are a
    restatement of the published library algorithm.  PARITY PIN: tests/golden/amed_*.npz are produced by running
    the UNMODIFIED plugin file over a stand-in of that base (oracle/ref_shim.py::_dpm_base) — this pins every line
    that lives in the reference; the inherited pieces are pinned only to the restatement ("parity unpinned" for
    those).  ODE variants only (dpmsolver / dpmsolver++), solver_order 1-3, no thresholding."""

    def __init__(self, *, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, solver_order=2, prediction_type="epsilon", algorithm_type="dpmsolver++",
                 solver_type="midpoint", lower_order_final=True, euler_at_final=False, final_sigmas_type="zero",
                 timestep_spacing="linspace", steps_offset=0, scale_dirs=None, scale_times=None):
        if algorithm_type not in ("dpmsolver", "dpmsolver++") or solver_order not in (1, 2, 3):
            raise NotImplementedError
        self.T = num_train_timesteps
        self.alphas_cumprod = sd_alphas_cumprod(sd_betas(num_train_timesteps, beta_start, beta_end, beta_schedule,
                                                         trained_betas))
        self.order, self.prediction_type, self.algo, self.solver_type = (solver_order, prediction_type,
                                                                         algorithm_type, solver_type)
        self.lower_order_final, self.euler_at_final, self.final_sigmas_type = (lower_order_final, euler_at_final,
                                                                               final_sigmas_type)
        self.spacing, self.offset = timestep_spacing, steps_offset
        self.scale_dirs, self.scale_times = scale_dirs, scale_times
        self.n = None

    def set_timesteps(self, num_inference_steps=None, timesteps=None):
        all_sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        if timesteps is None:                         # stock diffusers grid (restated; basic interpolation only)
            n, T = num_inference_steps, self.T
            if self.spacing == "linspace":
                ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
            elif self.spacing == "leading":
                ts = (np.arange(0, n + 1) * (T // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + self.offset
            elif self.spacing == "trailing":
                ts = np.arange(T, 0, -(T / n)).round().copy().astype(np.int64) - 1
            else:
                raise ValueError(self.spacing)
            sig = np.interp(ts, np.arange(0, len(all_sigmas)), all_sigmas)
            last = all_sigmas[0] if self.final_sigmas_type == "sigma_min" else 0
            self.sigmas = torch.from_numpy(np.concatenate([sig, [last]]).astype(np.float32))
            self.timesteps = torch.from_numpy(ts)
            self.n = len(ts)
        else:                                         # plugin :47-60
            self.sigmas = torch.from_numpy(all_sigmas[timesteps])
            self.timesteps = torch.tensor(timesteps[:-1], dtype=torch.int64)
            for i in range(len(self.scale_times)):
                if i % 2 == 1:
                    target = self.sigmas[i] * self.scale_times[i]
                    src = torch.tensor(all_sigmas[timesteps[i + 1] + 1:timesteps[i - 1]])
                    self.timesteps[i] = timesteps[i + 1] + 1 + torch.argmin(torch.abs(src - target))
            self.n = len(timesteps)                   # :60 counts the trailing 0 as well
        self.hist: List[torch.Tensor] = []            # converted model outputs, newest first
        self.lower_order_nums = 0
        self.step_index = None

    @staticmethod
    def _alpha_sigma(sigma):
        alpha = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha, sigma * alpha

    def _convert(self, e, sample):
        alpha, sig = self._alpha_sigma(self.sigmas[self.step_index])
        pp = self.algo == "dpmsolver++"
        if self.prediction_type == "epsilon":
            return (sample - sig * e) / alpha if pp else e
        if self.prediction_type == "sample":
            return e if pp else (sample - alpha * e) / sig
        if self.prediction_type == "v_prediction":
            return alpha * sample - sig * e if pp else alpha * e + sig * sample
        raise ValueError(self.prediction_type)

    def step(self, model_output, timestep, sample):
        if self.n is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if self.step_index is None:
            hits = (self.timesteps == timestep).nonzero()
            self.step_index = (len(self.timesteps) - 1 if len(hits) == 0 else
                               hits[1].item() if len(hits) > 1 else hits[0].item())
        i, nts = self.step_index, len(self.timesteps)
        final_first = (i == nts - 1) and (self.euler_at_final or (self.lower_order_final and nts < 15)
                                          or self.final_sigmas_type == "zero")                         # :394-398
        m0 = self._convert(model_output, sample)
        self.hist = ([m0] + self.hist)[: self.order]
        sample = sample.to(F32)                                                                          # :409
        sd = 1.0 if self.scale_dirs is None else self.scale_dirs[i]                                     # :417
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i + 1])
        alpha_s, sigma_s = self._alpha_sigma(self.sigmas[i])
        h = (torch.log(alpha_t) - torch.log(sigma_t)) - (torch.log(alpha_s) - torch.log(sigma_s))
        pp = self.algo == "dpmsolver++"
        if self.order == 1 or self.lower_order_nums < 1 or final_first:                                 # :418-419
            if pp:
                x = (sigma_t / sigma_s) * sample - sd * (alpha_t * (torch.exp(-h) - 1.0)) * m0          # :121
            else:
                x = (alpha_t / alpha_s) * sample - sd * (sigma_t * (torch.exp(h) - 1.0)) * m0           # :123
        elif not (self.order == 2 or self.lower_order_nums < 2 or
                  ((i == nts - 2) and self.lower_order_final and nts < 15)):                            # :422-423
            alpha_p, sigma_p = self._alpha_sigma(self.sigmas[i - 1])                                    # third order
            alpha_q, sigma_q = self._alpha_sigma(self.sigmas[i - 2])                                    # :307-346
            lam_s = torch.log(alpha_s) - torch.log(sigma_s)
            lam_p = torch.log(alpha_p) - torch.log(sigma_p)
            lam_q = torch.log(alpha_q) - torch.log(sigma_q)
            h_0, h_1 = lam_s - lam_p, lam_p - lam_q
            r0, r1 = h_0 / h, h_1 / h
            m1, m2 = self.hist[1], self.hist[2]
            d1_0, d1_1 = (1.0 / r0) * (m0 - m1), (1.0 / r1) * (m1 - m2)
            d1 = d1_0 + (r0 / (r0 + r1)) * (d1_0 - d1_1)
            d2 = (1.0 / (r0 + r1)) * (d1_0 - d1_1)
            if pp:
                x = ((sigma_t / sigma_s) * sample - sd * (alpha_t * (torch.exp(-h) - 1.0)) * m0
                     + sd * (alpha_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * d1
                     - sd * (alpha_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5)) * d2)
            else:
                x = ((alpha_t / alpha_s) * sample - sd * (sigma_t * (torch.exp(h) - 1.0)) * m0
                     - sd * (sigma_t * ((torch.exp(h) - 1.0) / h - 1.0)) * d1
                     - sd * (sigma_t * ((torch.exp(h) - 1.0 - h) / h ** 2 - 0.5)) * d2)
        else:                                                                                           # :420-421
            alpha_p, sigma_p = self._alpha_sigma(self.sigmas[i - 1])
            lam_s = torch.log(alpha_s) - torch.log(sigma_s)
            h_0 = lam_s - (torch.log(alpha_p) - torch.log(sigma_p))
            r0 = h_0 / h
            d1 = (1.0 / r0) * (m0 - self.hist[1])                                                       # :201
            if pp and self.solver_type == "midpoint":                                                   # :205-209
                x = ((sigma_t / sigma_s) * sample - sd * (alpha_t * (torch.exp(-h) - 1.0)) * m0
                     - sd * 0.5 * (alpha_t * (torch.exp(-h) - 1.0)) * d1)
            elif pp:                                                                                    # :211-215
                x = ((sigma_t / sigma_s) * sample - sd * (alpha_t * (torch.exp(-h) - 1.0)) * m0
                     + sd * (alpha_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * d1)
            elif self.solver_type == "midpoint":                                                        # :219-223
                x = ((alpha_t / alpha_s) * sample - sd * (sigma_t * (torch.exp(h) - 1.0)) * m0
                     - sd * 0.5 * (sigma_t * (torch.exp(h) - 1.0)) * d1)
            else:                                                                                       # :225-229
                x = ((alpha_t / alpha_s) * sample - sd * (sigma_t * (torch.exp(h) - 1.0)) * m0
                     - sd * (sigma_t * ((torch.exp(h) - 1.0) / h - 1.0)) * d1)
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1
        self.step_index += 1
        return x.to(m0.dtype)                                                                            # :429


def run_sd_preview(sched: OracleSDScheduler, x_T: torch.Tensor, pairs: Sequence[torch.Tensor], guidance: float,
                   qs: Optional[Sequence[torch.Tensor]] = None, forced_idx=None):
    """The caller loop of denoise_ppo.py:62-113 with the denoiser replaced by given CFG pairs
    ([2B, ...] each, unconditional half first, :66/:97).  Returns (final latents, per-step records)."""
    x = x_T
    rec = []
    for i, t in enumerate(sched.timesteps):
        u, c = pairs[i].chunk(2)
        eps = cfg_combine(u, c, guidance)
        out = sched.step(eps, t, x, q=None if qs is None else qs[i],
                         forced_idx=None if forced_idx is None else forced_idx[i])
        x = out[0]
        rec.append(out)
    return x, rec


# --------------------------------------------------------------------------------------------
# PPO update side
# --------------------------------------------------------------------------------------------
def ppo_loss_replicated(sd: Dict[str, torch.Tensor], x: torch.Tensor, actions: torch.Tensor, old_probs: torch.Tensor,
                        masks: torch.Tensor, rewards: torch.Tensor, variant: str, clip_range: float,
                        entropy_coef: float) -> torch.Tensor:
    """train_ppo.py:376-427 restated literally on the reference's layout: x [B,n',2] (B replicated condition rows
    per step), actions / old_probs / masks [B,n',A], rewards [B,1].  `sd` entries may require grad."""
    B, n1, A = actions.shape
    adv = (rewards - rewards.mean()) / (rewards.std() + 1e-8) * 10                 # :376
    adv = adv.repeat(1, n1).reshape(B * n1, -1)                                    # :377-379
    adv = adv * masks.reshape(B * n1, A)                                           # :390
    cur, ent = action_probs_entropy(sd, x.reshape(B * n1, 2), actions.reshape(B * n1, A), variant)   # :408
    logp = (cur + 1e-9).log().sum(dim=1).unsqueeze(1)                              # :410
    old = (old_probs.reshape(B * n1, A) + 1e-9).log().sum(dim=1).unsqueeze(1)      # :411
    ratio = (logp - old).exp()
    clipped = torch.clamp(ratio, 1 - clip_range, 1 + clip_range)
    policy_loss = -torch.min(adv * ratio, adv * clipped).mean()                    # :416-420
    return policy_loss + (-entropy_coef * ent.mean())                              # :424-427
