"""PPO update side of ConsistencySolver: the clipped-ratio policy loss on a rollout record and the
data-parallel gradient exchange.  Reference: train_ppo.py:376-437 (advantages, loss), factor_net_ppo.py:170-184
(`get_action_probs`), accelerate/DDP gradient all-reduce (train_ppo.py:257,:430; edit_ppo/train_ppo.py:177,:382).

What is different from the reference, none of it numerical:
  * the MLP is evaluated on the n-1 DISTINCT condition rows of a rollout, not on B*(n-1) replicated rows
    (`conds['x']` is `[[t, prev_t]].repeat(B, 1)`, scheduler_ppo.py:207-210); the per-sample probabilities are gathers
    from those tables, so forward values are identical and the gradient is the same sum;
  * bins are taken from the recorded indices (the reference re-derives them with an argmin over |a - values|,
    factor_net_ppo.py:174-178, which round-trips exactly — SURVEY §8a T1);
  * gradients live in ONE flat fp32 buffer (75 041 floats for the production policy); the exchange is a single
    all-reduce(AVG) on it over NCCL/NVLink — latency-bound, no bucketing — instead of DDP's hook machinery;
  * the random step count of a rollout comes from a shared seeded RNG instead of a broadcast
    (edit_ppo/train_ppo.py:275-283).
This module is training-side torch code (autograd); it is not the sampling hot path and runs on any device."""
from __future__ import annotations

import random
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------------------------
# flat parameter / gradient storage
# ---------------------------------------------------------------------------------------------------------------
class FlatParams:
    """Re-homes the parameters (and their .grad) of a module into two contiguous fp32 buffers, so the optimizer
    sees ordinary parameters while collectives and checksums touch one tensor."""

    def __init__(self, module: torch.nn.Module):
        ps = [p for p in module.parameters() if p.requires_grad]
        if not ps:
            raise ValueError("module has no trainable parameters")
        dev, dt = ps[0].device, ps[0].dtype
        n = sum(p.numel() for p in ps)
        self.flat = torch.empty(n, device=dev, dtype=dt)
        self.grad = torch.zeros(n, device=dev, dtype=dt)
        o = 0
        for p in ps:
            k = p.numel()
            self.flat[o:o + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[o:o + k].view_as(p)
            p.grad = self.grad[o:o + k].view_as(p)
            o += k
        self.params = ps
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()
        for p, (o, k) in zip(self.params, self._spans()):
            if p.grad is None or p.grad.data_ptr() != self.grad[o:o + k].data_ptr():
                p.grad = self.grad[o:o + k].view_as(p)      # an optimizer's zero_grad(set_to_none=True) detached it

    def _spans(self):
        o = 0
        for p in self.params:
            yield o, p.numel()
            o += p.numel()

    def checksum(self) -> float:
        """the reference's DDP self-check: a per-rank parameter sum (train_ppo.py:452-455)"""
        return float(self.flat.double().sum())


def broadcast_parameters(flat: FlatParams, src: int = 0):
    """C2 of SURVEY §2.2: one broadcast of the weights from rank 0 at start."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat.flat, src=src)


def allreduce_gradients(flat: FlatParams):
    """C1 of SURVEY §2.2: average the flat gradient buffer over ranks — one latency-bound collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat.grad, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
        flat.grad.div_(dist.get_world_size())


class PeerGradExchange:
    """One-shot all-reduce (AVG) of the flat policy gradient over NVLink peer memory, FUSED into the PPO reduction kernel
    (csrc/ppo.cu::ppo_reduce_allreduce_kernel) — the B200-native replacement for the DDP gradient all-reduce of
    train_ppo.py:257,:430 / edit_ppo/train_ppo.py:382.  300 KB is latency-bound: NCCL costs 23-35 us at 2-8 GPUs; here
    every rank stores its gradient into a peer-mapped buffer, signals, and sums all ranks' buffers in rank order (every
    rank gets the bit-identical average), inside the kernel that produced the gradient.

    Buffers come from torch's symmetric-memory allocator (cudaMalloc'd / fabric memory mapped into every peer over
    NVLink/NVSwitch); the kernel only sees raw pointers (consolver_peers_t).  Collective: construct on all ranks."""

    def __init__(self, flat: FlatParams, group=None, policy=None):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGradExchange needs an initialised process group")
        group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("one-shot exchange supports up to 16 ranks of one NVLink domain")
        dev = flat.grad.device
        self.stride = (flat.numel + 63) // 64 * 64                     # floats per parity, 256-byte granules
        # one flag per (rank, CTA of the exchange kernel): 1024 gradient elements per CTA (+ the statistics slot)
        self.pad_words = pad_words = (self.world * ((flat.numel + 1 + 1023) // 1024) + 63) // 64 * 64
        _lib.load()
        self.buf = symm_mem.empty(2 * self.stride + pad_words, dtype=torch.float32, device=dev)
        self.buf.zero_()
        try:
            self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        except Exception:  # noqa: BLE001  (older torch: the group has to be enabled explicitly first)
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                symm_mem.enable_symm_mem_for_group(group.group_name)
            self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        torch.cuda.synchronize(dev)
        self.hdl.barrier()                                              # everybody's pad is zero before anyone signals
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._buf_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self._sig_ptrs = torch.tensor([p + 2 * self.stride * 4 for p in ptrs], dtype=torch.int64, device=dev)
        self.epoch = 0

    def next_peers(self):
        """consolver_peers_t for the next fused update (every rank must make the same sequence of calls)"""
        from . import _lib

        self.epoch += 1
        p = _lib.Peers(self._buf_ptrs.data_ptr(), self._sig_ptrs.data_ptr(), self.rank, self.world,
                       self.epoch & 0xFFFFFFFF, self.stride, self.pad_words)
        self._keepalive = p
        return p


def shared_step_count(step: int, seed: int, lo: int = 2, hi: int = 15) -> int:
    """Number of inference steps of rollout `step`, identical on every rank without a collective
    (train_ppo.py:345 draws random.choice(range(2,16)) under identical seeds; the FLUX driver broadcasts it)."""
    return random.Random(seed * 1_000_003 + step).choice(list(range(lo, hi + 1)))


# ---------------------------------------------------------------------------------------------------------------
# advantages and loss
# ---------------------------------------------------------------------------------------------------------------
def advantages_from_rewards(rewards: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    """train_ppo.py:376-390: standardise within the rank's batch, x10, repeat per step, zero the unused dims.
    rewards [B,1], masks [B,n',A] -> advantages [B,n',A]."""
    adv = (rewards - rewards.mean()) / (rewards.std() + 1e-8) * 10
    return adv.view(-1, 1, 1) * masks


def ppo_loss(factor_net, x_rows: torch.Tensor, idx: torch.Tensor, old_probs: torch.Tensor,
             advantages: torch.Tensor, clip_range: float = 0.2, entropy_coef: float = 0.0
             ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Clipped PPO loss of train_ppo.py:406-427 on one rollout record.

      x_rows     [n',2]     the distinct condition rows (t, prev_t) of steps 1..n-1 (`trajectory()['x'][0]`)
      idx        [B,n',A]   sampled bin indices            old_probs [B,n',A]  their probabilities at rollout time
      advantages [B,n',A]   from advantages_from_rewards
    """
    fn = factor_net.module if hasattr(factor_net, "module") else factor_net
    tables = fn.forward_({"x": x_rows})                                       # [n',A,K], autograd
    n1, A, K = tables.shape
    B = idx.shape[0]
    cur = tables.unsqueeze(0).expand(B, n1, A, K).gather(3, idx.unsqueeze(-1)).squeeze(-1)   # [B,n',A]
    logp = (cur + 1e-9).log().sum(dim=2, keepdim=True)                         # joint over action dims
    old_logp = (old_probs + 1e-9).log().sum(dim=2, keepdim=True)
    ratio = (logp - old_logp).exp()
    clipped = torch.clamp(ratio, 1 - clip_range, 1 + clip_range)
    policy_loss = -torch.min(advantages * ratio, advantages * clipped).mean()
    # Categorical(probs).entropy() / log K (factor_net_ppo.py:180-181); its mean over the B*n'*A replicated
    # entries equals the mean over the distinct rows
    ent = torch.distributions.Categorical(probs=tables).entropy() / torch.log(
        torch.as_tensor(K, dtype=tables.dtype, device=tables.device))
    entropy_loss = -entropy_coef * ent.mean()
    loss = policy_loss + entropy_loss
    return loss, dict(policy_loss=policy_loss.detach(), entropy=ent.mean().detach(), ratio_mean=ratio.mean().detach())


def ppo_loss_grad_cuda(factor_net, flat: FlatParams, x_rows: torch.Tensor, idx_rba: torch.Tensor,
                       old_rba: torch.Tensor, adv_rba: torch.Tensor, clip_range: float, entropy_coef: float,
                       workspace: Optional[torch.Tensor] = None,
                       exchange: Optional["PeerGradExchange"] = None) -> torch.Tensor:
    """Hand-written CUDA forward + loss + backward (csrc/ppo.cu): writes d loss / d params into `flat.grad` and
    returns stats [4] = {loss, policy_loss, mean entropy, mean ratio} (device tensor, no sync).
    Inputs in the trajectory buffers' native layout: idx / old probs / advantages as [R, B, A]."""
    from . import _lib

    fn = factor_net.module if hasattr(factor_net, "module") else factor_net
    if fn.use_conv:
        raise NotImplementedError("the PPO kernel covers the shared-row policy (use_conv=False)")
    lib = _lib.load()
    R, B, A = idx_rba.shape
    dev = idx_rba.device
    need = int(lib.consolver_ppo_workspace(R, fn.hidden_dim, fn.action_dims, fn.num_actions))
    if workspace is None or workspace.numel() * 4 < need:
        workspace = torch.empty((need + 3) // 4, device=dev, dtype=torch.float32)
    stats = torch.empty(4, device=dev, dtype=torch.float32)
    w = fn.kernel_weights()
    expected = sum(p.numel() for p in fn.parameters())
    if flat.grad.numel() != expected or flat.grad.dtype != torch.float32:
        raise ValueError("flat gradient buffer does not match the policy's parameters")
    import ctypes

    peers = ctypes.byref(exchange.next_peers()) if exchange is not None and exchange.world > 1 else None
    rc = lib.consolver_ppo_loss_grad_allreduce_f32(
        *w[:6], x_rows.data_ptr(), R, fn.x_div, fn.temperature, fn.hidden_dim, fn.action_dims, fn.num_actions,
        idx_rba.data_ptr(), old_rba.data_ptr(), adv_rba.data_ptr(), B, float(clip_range), float(entropy_coef),
        workspace.data_ptr(), flat.grad.data_ptr(), stats.data_ptr(), peers,
        torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "consolver_ppo_loss_grad_allreduce_f32")
    return stats


def ppo_update(factor_net, flat: FlatParams, optimizer, record: Dict[str, torch.Tensor], rewards: torch.Tensor,
               ppo_epochs: int = 1, clip_range: float = 0.2, entropy_coef: float = 0.0,
               max_grad_norm: Optional[float] = 1.0, native: Optional[bool] = None,
               exchange: Optional["PeerGradExchange"] = None, read_back: bool = True) -> Dict[str, float]:
    """ppo_epochs x (loss, backward, flat all-reduce, clip, optimizer step) — train_ppo.py:406-437.
    `record` is `scheduler.trajectory()` (views); it is detached/cloned here because the next rollout reuses the
    buffers.  `native` (default: on for CUDA tensors with the shared-row policy) runs forward + loss + backward as the
    hand-written kernels of csrc/ppo.cu; otherwise torch autograd on the distinct rows.  `exchange` (PeerGradExchange, native
    path only): the gradient all-reduce is fused into the reduction kernel over NVLink peer memory instead of a NCCL call.
    `read_back=False` (native path): no host synchronisation at all — the statistics come back as DEVICE tensors
    (`stats` [4] = loss, policy loss, entropy, mean ratio; `grad_norm` 0-d), for loops that log every k-th iteration (the
    reference reads the loss back every iteration, train_ppo.py:458)."""
    fn0 = factor_net.module if hasattr(factor_net, "module") else factor_net
    if getattr(fn0, "use_conv", False):
        # use_conv makes the policy input per-sample (cosine features of the model-output history), which the rollout
        # record does not keep — the scheduler's ring is overwritten as the trajectory advances (the reference keeps
        # hundreds of MiB of conds['epsilon'] for this, denoise_ppo.py:105-118; no shipped config trains with use_conv)
        raise NotImplementedError("ppo_update: training a use_conv=True policy is not supported (the rollout record holds "
                                  "no per-step model-output history); sample with it, or train with use_conv=False")
    if getattr(fn0, "continuous", False):
        raise NotImplementedError("ppo_update covers the discrete policy; for the continuous extension build the loss from "
                                  "factor_net(conds, actions) (density, entropy) with autograd")
    x_rows = record["x"][0].detach().float().clone()
    idx = record["idx"].detach().clone()
    old_probs = record["probs"].detach().clone()
    adv = advantages_from_rewards(rewards.detach(), record["masks"].detach()).clone()
    fn = factor_net.module if hasattr(factor_net, "module") else factor_net
    if native is None:
        native = bool(idx.is_cuda and not fn.use_conv and flat.grad.dtype == torch.float32)
    stats = {}
    if native:
        idx_r, old_r, adv_r = (t.transpose(0, 1).contiguous() for t in (idx, old_probs, adv))   # [R,B,A]
        x_rows = x_rows.contiguous()
        ws = None
        for _ in range(ppo_epochs):
            for p, (o, k) in zip(flat.params, flat._spans()):       # keep p.grad aliased to the flat buffer
                if p.grad is None or p.grad.data_ptr() != flat.grad[o:o + k].data_ptr():
                    p.grad = flat.grad[o:o + k].view_as(p)
            st = ppo_loss_grad_cuda(fn, flat, x_rows, idx_r, old_r, adv_r, clip_range, entropy_coef, ws, exchange)
            if exchange is None:
                allreduce_gradients(flat)      # NCCL; with `exchange` the all-reduce already happened inside the kernel
            if max_grad_norm is not None:
                norm = flat.grad.norm()
                flat.grad.mul_(torch.clamp(max_grad_norm / (norm + 1e-6), max=1.0))
            optimizer.step()
        if not read_back:
            stats.update(stats=st)
            if max_grad_norm is not None:
                stats["grad_norm"] = norm
            return stats
        vals = st.tolist()      # one read-back per update, after the last epoch
        stats.update(loss=vals[0], policy_loss=vals[1], entropy=vals[2], ratio_mean=vals[3])
        if max_grad_norm is not None:
            stats["grad_norm"] = float(norm)
        return stats
    for _ in range(ppo_epochs):
        flat.zero_grad()
        loss, info = ppo_loss(factor_net, x_rows, idx, old_probs, adv, clip_range, entropy_coef)
        loss.backward()
        allreduce_gradients(flat)
        if max_grad_norm is not None:
            norm = flat.grad.norm()
            flat.grad.mul_(torch.clamp(max_grad_norm / (norm + 1e-6), max=1.0))    # clip_grad_norm_ on the flat buffer
            stats["grad_norm"] = float(norm)
        optimizer.step()
        stats.update(loss=float(loss.detach()), **{k: float(v) for k, v in info.items()})
    return stats


# ---------------------------------------------------------------------------------------------------------------
# rollout (SD): B replicas of ONE (noise, target) pair differing only by the sampled actions
# ---------------------------------------------------------------------------------------------------------------
def rollout_sd(scheduler, denoiser, noise_one: torch.Tensor, batch: int, cfg: float, num_inference_steps: int):
    """data_processing.py:65-80 (`repeat_random_sample`) + denoise_ppo.py:62-118: replicate one noise sample B times,
    run the CFG sampling loop, return (final latents [B,...], record views)."""
    from .denoise import denoise_loop

    noise = noise_one.unsqueeze(0).expand(batch, *noise_one.shape).contiguous()
    with torch.no_grad():
        return denoise_loop(scheduler, denoiser, noise, cfg=cfg, num_inference_steps=num_inference_steps)


class GraphedRollouts:
    """Rollouts as CUDA graphs, one per step count.  train_ppo.py:345 draws the number of inference steps of every rollout
    from 2..15, so the whole CFG sampling loop (denoiser included) is captured once per count (`GraphedDenoiseLoop`) and a
    rollout is ONE graph launch instead of ~10 eager launches per step — eager stepping costs ~175 us of host time per
    solver step against ~15 us of GPU time at batch 80.  Every count has its own scheduler object (its own trajectory
    buffers and history ring) sharing ONE `factor_net`: the graphs re-evaluate the probability tables from the live
    weights at every replay, so optimizer steps between rollouts are seen.  The default generator is consumed exactly as
    by eager rollouts (same draws, same order).

    `denoiser(model_in [2B,...], t, i)` must be capturable.  Build it AFTER `FlatParams(factor_net)` (the graphs hold the
    parameter addresses)."""

    def __init__(self, scheduler, denoiser, noise_one: torch.Tensor, batch: int, cfg: float,
                 step_counts=range(2, 16), prebuild: bool = True):
        self.proto, self.denoiser, self.cfg, self.batch = scheduler, denoiser, float(cfg), int(batch)
        self.noise = noise_one.unsqueeze(0).expand(batch, *noise_one.shape).contiguous()
        self.loops: Dict[int, object] = {}
        if prebuild:
            for n in step_counts:
                self._loop(int(n))

    def _loop(self, n: int):
        from .denoise import GraphedDenoiseLoop

        if n not in self.loops:
            s = type(self.proto)(**dict(self.proto.config))
            s.factor_net = self.proto.factor_net                     # shared, live weights
            for k in ("reference_device", "use_pdl", "use_fused_rng"):
                setattr(s, k, getattr(self.proto, k))
            self.loops[n] = GraphedDenoiseLoop(s, self.denoiser, self.noise, self.cfg, n)
        return self.loops[n]

    def rollout(self, num_inference_steps: int, noise_one: Optional[torch.Tensor] = None):
        """-> (final latents [B,...], rollout record as views — valid until the next rollout with the same step count)"""
        loop = self._loop(int(num_inference_steps))
        noise = None
        if noise_one is not None:
            noise = noise_one.unsqueeze(0).expand(self.batch, *noise_one.shape)
        lat = loop.replay(noise)
        return lat, loop.record()


def latent_mse_reward(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Synthetic stand-in for the reference's image-space rewards (edit_ppo/reward_model.py needs pretrained
    weights): negative latent MSE to the teacher latent, shape [B,1] like calculate_reward's outputs."""
    return -((pred.float() - target.float()) ** 2).flatten(1).mean(dim=1, keepdim=True)
