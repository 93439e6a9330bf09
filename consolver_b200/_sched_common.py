"""Pieces shared by PPOScheduler and FMPPOScheduler: the per-trajectory device state, the draw-source selection
(in-kernel RNG / torch launch / replay), the lazy `conds` value and the rollout-record views."""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .config_utils import LazyConds
from .factor_net import alloc_policy_outputs


#: default of `scheduler.reference_device` for newly built schedulers ("cuda" unless CONSOLVER_REFERENCE_DEVICE=cpu)
DEFAULT_REFERENCE_DEVICE = os.environ.get("CONSOLVER_REFERENCE_DEVICE", "cuda")


class Trajectory:
    """Per-(set_timesteps, batch shape) device state: policy outputs for every step, the Exp(1) buffer, the
    history ring, the condition rows of the whole grid.  Allocated once per trajectory; a step allocates only the
    latent it returns, and its `actions / probs / masks` are views of row i of these buffers."""

    def __init__(self, fn, n_steps: int, order_dim: int, B: int, shape, dtype, device, rows: Sequence[Sequence[float]]):
        self.key = (B, tuple(shape), dtype, device)
        self.n = max(int(n_steps), 1)
        self.order_dim = order_dim
        A, K = fn.action_dims, fn.num_actions
        self.out = alloc_policy_outputs(B, A, K, order_dim, device, lead=(self.n,))
        # rows 0..n-1: tables of the grid (one table-kernel launch per pass); row n: scratch for an off-grid timestep
        self.out["probs_table"] = torch.empty(self.n + 1, A, K, device=device, dtype=torch.float32)
        self.last_table_row = 0
        # row pointers by arithmetic: indexing a tensor costs ~2 us of host time, a step needs seven of them
        self._row = {k: (v.data_ptr(), v.stride(0) * v.element_size()) for k, v in self.out.items()}
        self.q = torch.empty((B * A, K), device=device, dtype=torch.float32)
        self.ring = None                       # [order_dim, B, *shape], allocated on the first fused-CFG step
        self.ring_epoch = 0                    # bumped by rewind(): stale conds['epsilon'] views can tell
        self.grid_rows: List[int] = []         # condx row used by each step of this pass (mid-grid starts, repeats)
        self.ring_shape = (order_dim, B, *shape)
        self.dtype, self.device = dtype, device
        # conds['x'] rows of the whole grid, rounded through the model dtype as the reference's
        # torch.tensor([[a, b]], dtype=model_output.dtype) does: one H2D copy per trajectory, not one per step
        host = torch.tensor([list(map(float, r)) for r in rows], dtype=dtype)
        self.condx = host.to(device, non_blocking=True)
        self.condx_f32 = host.float().to(device, non_blocking=True)    # policy-table kernel input [n,2]
        self.condx_host = host.float().numpy()
        self.count = 0
        self.table_pass = -1       # pass (count // n) whose probability tables are in out['probs_table']
        self.rng_plan = None       # (nthreads, offset increment) of torch's exponential_ launch for q's numel
        self.graph_rng = None      # device int64[2] {seed, offset} for CUDA-graph replays
        self.graph_rng_used = 0
        self.policy_forked = False  # the policy side stream has been forked off the main stream in this pass
        self._conv = None
        self._fixed: Dict[int, torch.Tensor] = {}

    def p(self, name: str, i: int) -> int:
        base, stride = self._row[name]
        return base + i * stride

    def rewind(self):
        """restart the pass (GraphedPreview): same buffers, tables re-evaluated, RNG bookkeeping reset"""
        self.count = 0
        self.table_pass = -1
        self.graph_rng_used = 0
        self.policy_forked = False
        self.ring_epoch += 1
        self.grid_rows = []

    def slot(self, i: int) -> torch.Tensor:
        if self.ring is None:
            self.ring = torch.empty(self.ring_shape, device=self.device, dtype=self.dtype)
        return self.ring[i % self.ring_shape[0]]

    def conv_buffers(self, fn):
        """use_conv=True scratch: features [B,od-1], reduction workspace, per-sample tables [n,B,A,K]"""
        if self._conv is None:
            from .features import workspace_bytes

            B, od = self.key[0], self.order_dim
            self._conv = (torch.empty(B, od - 1, device=self.device, dtype=torch.float32),
                          torch.empty(workspace_bytes(B, od) // 8 + 1, device=self.device, dtype=torch.float64),
                          torch.empty(self.n, B, fn.action_dims, fn.num_actions, device=self.device))
        return self._conv

    def fixed_rows(self, coef_fn, n_hist: int) -> torch.Tensor:
        """coefficient records [B, od+2] of a fixed-coefficient (baseline) solver at history depth n_hist"""
        if n_hist not in self._fixed:
            c = [float(v) for v in coef_fn(n_hist)]
            if len(c) != n_hist:
                raise ValueError("fixed_coefficients(n_hist) must return n_hist values")
            od, B = self.order_dim, self.key[0]
            row = torch.tensor(c + [0.0] * (od - n_hist) + [1.0, 1.0], dtype=torch.float32)
            self._fixed[n_hist] = row.to(self.device).expand(B, od + 2).contiguous()
        return self._fixed[n_hist]

    def record(self, skip_first: bool = True) -> Dict[str, torch.Tensor]:
        """The rollout record denoise_ppo.py:105-118 assembles with unsqueeze + cat, as views:
        x [B,n',2], probs / actions / masks / idx / logp [B,n',A] for steps 1..count-1 (or 0.. if not skip_first)."""
        lo, hi = (1 if skip_first else 0), min(self.count, self.n)
        B = self.key[0]
        pick = lambda k: self.out[k][lo:hi].transpose(0, 1)  # noqa: E731
        rows = self.grid_rows[lo:hi]
        if len(rows) == hi - lo and rows != list(range(lo, hi)) and all(r is not None for r in rows):
            # the run did not walk the grid from its first point (img2img-style start, repeated timesteps): gather the
            # condition rows the steps actually used instead of slicing by step count
            x_rows = self.condx.index_select(0, torch.tensor(rows, device=self.condx.device))
        else:
            x_rows = self.condx[lo:hi]
        return dict(x=x_rows.unsqueeze(0).expand(B, hi - lo, 2), probs=pick("probs"),
                    actions=pick("actions"), masks=pick("masks"), idx=pick("idx"), logp=pick("logp"))

    def last(self, table_row: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """table [A,K] (or [B,A,K] with use_conv), indices [B,A], coefficient records and log-probs of the latest step"""
        i = (self.count - 1) % self.n
        row = self.last_table_row if table_row is None else table_row
        table = self._conv[2][i] if self._conv is not None else self.out["probs_table"][row]
        return dict(probs_table=table, idx=self.out["idx"][i], coef=self.out["coef"][i], logp=self.out["logp"][i])


def next_rng(sched, tr: Trajectory, device):
    """consolver_rng_t for this step's in-kernel draw, or None to use the torch exponential_ launch."""
    from . import rng as _rng

    if not sched.use_fused_rng:
        return None
    if torch.cuda.is_current_stream_capturing():
        if tr.graph_rng is None or tr.rng_plan is None:
            return None
        nthreads, inc = tr.rng_plan
        r = _lib.Rng(0, tr.graph_rng_used * inc, tr.graph_rng.data_ptr(), nthreads)
        tr.graph_rng_used += 1
    else:
        if not _rng.fused_rng_available(device):
            return None
        if tr.rng_plan is None:
            tr.rng_plan = _lib.philox_plan(tr.q.numel())
        nthreads, inc = tr.rng_plan
        seed, off = _rng.take(device, inc)
        r = _lib.Rng(seed, off, None, nthreads)
    tr._rng_keepalive = r
    return r


def draw_source(sched, tr: Trajectory, device, fused_ok: bool):
    """Where this step's categorical draw comes from -> (q_ptr, idx_ptr, rng_arg), exactly one non-null unless the
    scheduler runs with fixed coefficients:
      * in-kernel RNG (torch's exponential_ stream regenerated by the sample kernel) when `fused_ok`,
      * else the torch launch `q.exponential_(1)` — the draw torch.multinomial makes (factor_net_ppo.py:161),
      * `replay['idx']`: forced bins,  `replay['q']`: supplied Exp(1) values."""
    if sched.fixed_coefficients is not None:
        return None, None, None                       # baseline solvers draw nothing
    rp = sched.replay
    if rp is None:
        r = next_rng(sched, tr, device) if fused_ok else None
        if r is not None:
            return None, None, ctypes.byref(r)
        tr.q.exponential_(1)
        return tr.q.data_ptr(), None, None
    if rp.get("idx") is not None:
        forced = rp["idx"][tr.count].to(device=device, dtype=torch.int64).contiguous()
        tr._forced_keepalive = forced
        return None, forced.data_ptr(), None
    tr.q.copy_(rp["q"][tr.count].reshape(tr.q.shape))
    return tr.q.data_ptr(), None, None


def lazy_conds(conds_x: torch.Tensor, hist_now: List[torch.Tensor], order_dim: int, tr: "Trajectory" = None,
               ring: bool = False) -> LazyConds:
    """`conds` of the reference's return: 'x' now, 'epsilon' (newest-first zero-padded stack, scheduler_ppo.py:222-237)
    only when somebody reads it.  `ring`: the history tensors are slots of the trajectory's ring (step_cfg); the stack is
    only valid until the oldest of them is overwritten, i.e. while fewer than order_dim - len(history) + 1 further
    steps have run."""
    def stack():
        s = torch.stack(hist_now, dim=1)
        if len(hist_now) < order_dim:
            pad = s.new_zeros(s.shape[0], order_dim - len(hist_now), *s.shape[2:])
            s = torch.cat([s, pad], dim=1)
        return s

    valid = None
    if ring and tr is not None:
        made_at, depth, slots, pass_id = tr.count, len(hist_now), tr.ring_shape[0], tr.ring_epoch
        # the oldest referenced slot was written at step made_at - depth and is rewritten at step made_at - depth + slots
        valid = lambda: tr.ring_epoch == pass_id and tr.count <= made_at - depth + slots  # noqa: E731
    return LazyConds(conds_x, stack, valid)


class SolverOptions:
    """Optional knobs shared by both schedulers (none of them changes a number)."""

    def _init_solver_options(self):
        #: Which execution of the reference the arithmetic reproduces bit for bit.  "cuda" (default): the reference run
        #: on CUDA tensors, the way its drivers run it — ATen hands the host-resident schedule scalars to the kernels as
        #: fp32 values and turns `t / scalar` into `t * (1/scalar)` (pinned by tests/golden/cuda_*).  "cpu": the
        #: reference run on CPU tensors (true divisions, scalars rounded to a 16-bit tensor's dtype first; pinned by the
        #: CPU-made fixtures).  The two differ by an ulp here and there, also in fp32.
        self.reference_device = DEFAULT_REFERENCE_DEVICE
        #: True (default): a CUDA `timestep` tensor is NOT read back; the value is taken from the host copy of the grid
        #: at the current step count (pipelines step in grid order).  False: `.item()` it (one sync per step).
        self.sync_free = True
        #: link the policy and step kernels with programmatic dependent launch (CONSOLVER_NO_PDL=1 turns it off for a
        #: whole process: the first thing to try when a result looks order-dependent — profiles/live_fuzz_r02.md)
        self.use_pdl = os.environ.get("CONSOLVER_NO_PDL") != "1"
        #: generate the Exp(1) draw inside the sample kernel (bit-identical to torch's exponential_, see rng.py)
        self.use_fused_rng = True
        #: optional side stream for the policy kernels (set by GraphedPreview): the sample kernel needs nothing from the
        #: step kernels, so in a captured graph the sample chain becomes a parallel branch
        self.policy_stream: Optional[torch.cuda.Stream] = None
        #: solver-only replays (GraphedPreview): chain consecutive step kernels with programmatic dependent launch
        #: (CONSOLVER_FLAG_CHAIN).  Only valid when the model outputs are NOT produced by the kernel right before.
        self.chain_steps = False
        #: replay instead of sampling: {'idx': per-step [B,A] int64} forces the bins (PPO replay, parity tests with
        #: injected actions); {'q': per-step [B*A,K] fp32} supplies the Exp(1) draw
        self.replay: Optional[Dict] = None
        #: None, or a callable n_hist -> n_hist multipliers (newest first, summing to 1): the policy is bypassed and the
        #: fused step runs with these fixed coefficients (consolver_b200.baselines); `fixed_depth` caps the history
        self.fixed_coefficients = None
        self.fixed_depth: Optional[int] = None
        self._hist: List[torch.Tensor] = []    # model outputs, NEWEST FIRST (references or ring slots)
        self._traj: Optional[Trajectory] = None

    def _semantics(self, model_dtype: torch.dtype):
        """-> (step flags, policy flags, dtype the policy MLP runs in or None) for this call, from `reference_device`,
        the autocast state and the policy's own dtype.

        * autocast (gen_ppo.py:309, train_ppo.py:353, edit_ppo/train_ppo.py:289): nn.Linear runs in the autocast dtype,
          softmax and torch.sum in fp32.  A policy whose parameters are 16-bit only runs under autocast in the
          reference (fp32 inputs into a 16-bit Linear raise otherwise) and is treated the same.
        * a policy cast to the pipeline dtype, bin buffer included (gen_ppo.py:193-195): the per-sample coefficients
          are 16-bit tensors (CONSOLVER_FLAG_LOWP_COEF / CONSOLVER_POLICY_COEF_*)."""
        if self.reference_device not in ("cuda", "cpu"):
            raise ValueError("reference_device must be 'cuda' or 'cpu'")
        host = self.reference_device == "cpu"
        sflags = _lib.FLAG_HOST_SCALARS if host else 0
        pflags = _lib.POLICY_HOST_DIV if host else 0
        fn = self.factor_net_module
        act = None
        wdt = fn._kparams_dtype()
        if wdt in (torch.float16, torch.bfloat16):
            act = wdt
        elif torch.is_autocast_enabled("cuda"):
            ad = torch.get_autocast_dtype("cuda")
            if ad in (torch.float16, torch.bfloat16):
                act = ad
        if act is not None:
            pflags |= _lib.POLICY_ACT_F16 if act == torch.float16 else _lib.POLICY_ACT_BF16
        bdt = fn.action_values.dtype
        if bdt in (torch.float16, torch.bfloat16):
            pflags |= _lib.POLICY_COEF_F16 if bdt == torch.float16 else _lib.POLICY_COEF_BF16
            if bdt == model_dtype:
                sflags |= _lib.FLAG_LOWP_COEF
        return sflags, pflags, act

    @property
    def factor_net_module(self):
        fn = self.factor_net
        return fn.module if hasattr(fn, "module") else fn    # DDP-wrapped during training (scheduler_ppo.py:239)

    @property
    def ets(self) -> List[torch.Tensor]:
        """History oldest-first, the reference's attribute name (scheduler_ppo.py:123)."""
        return self._hist[::-1]

    def _history_depth(self, order_dim: int) -> int:
        if self.fixed_coefficients is None:
            return order_dim
        return min(order_dim, self.fixed_depth or order_dim)

    def trajectory(self, skip_first: bool = True) -> Dict[str, torch.Tensor]:
        if self._traj is None:
            raise ValueError("no trajectory recorded; call step() first")
        return self._traj.record(skip_first)


def strided_model_outputs(e0: torch.Tensor, older: List[torch.Tensor]):
    """Model outputs for `consolver_step_fm_strided` -> (e0, older, e_stride).  When the newest output and the history
    are all views with contiguous samples and ONE common sample stride (the `noise_pred[:, :L]` slices of a wider
    transformer output, edit_ppo/denoise_diffusion.py:140) they are consumed in place (e_stride = that stride);
    contiguous tensors give e_stride 0; any other mix is made contiguous (one copy each)."""
    every = [e0] + list(older)
    if all(t.is_contiguous() for t in every):
        return e0, older, 0
    n = e0[0].numel()
    stride = e0.stride(0)
    if stride % 8 == 0 and stride >= n and all(
            t.dim() > 1 and t[0].is_contiguous() and t.stride(0) == stride and t.shape == e0.shape for t in every):
        return e0, older, stride
    return e0.contiguous(), [h.contiguous() for h in older], 0
