"""PPO-update side on CPU: the distinct-row loss equals the reference's replicated-row loss
(restated with the oracle), flat-buffer plumbing, and 2-rank gloo data parallelism."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import consolver_b200 as cb
import consolver_oracle as orc
from consolver_b200 import ppo
from golden_io import Golden


def _policy(seed=0, variant="sd"):
    torch.manual_seed(seed)
    fn = cb.FactorNetPPO(hidden_dim=64, num_actions=11, order_dim=4, scaler_dim=0) if variant == "sd" else \
        cb.FactorNetPPOFM(hidden_dim=64, num_actions=11, order_dim=2, scaler_dim=0, mu_dim=0)
    with torch.no_grad():
        fn.mlp[4].weight.normal_(0, 0.3)
        fn.mlp[4].bias.normal_(0, 0.1)
    return fn


def _fake_record(fn, B, n1, seed, variant="sd"):
    g = torch.Generator().manual_seed(seed)
    A, K = fn.action_dims, fn.num_actions
    if variant == "sd":
        t = torch.tensor([874., 749., 624., 499., 374., 249., 124.])[:n1]
        x = torch.stack([t, t - 125], dim=1)
    else:
        x = torch.rand(n1, 2, generator=g)
    idx = torch.randint(0, K, (B, n1, A), generator=g)
    with torch.no_grad():
        tables = fn.forward_({"x": x})
    old = tables.unsqueeze(0).expand(B, n1, A, K).gather(3, idx.unsqueeze(-1)).squeeze(-1)
    old = (old * (1 + 0.2 * torch.randn(old.shape, generator=g))).clamp(1e-4, 1.0)     # an older policy
    masks = torch.ones(B, n1, A)
    masks[:, 0, 1:] = 0
    rewards = torch.randn(B, 1, generator=g)
    actions = fn.action_values[torch.arange(A).view(1, 1, A).expand(B, n1, A), idx]
    return dict(x=x.unsqueeze(0).expand(B, n1, 2), idx=idx, probs=old, masks=masks, actions=actions), rewards


def _reference_loss(fn, rec, rewards, clip, ent_coef, variant):
    """train_ppo.py:376-427 restated literally on B*(n-1) replicated rows: oracle.ppo_loss_replicated."""
    sd = dict(fn.state_dict())
    for k, v in fn.named_parameters():
        sd[k] = v                                               # keep autograd
    return orc.ppo_loss_replicated(sd, rec["x"], rec["actions"], rec["probs"], rec["masks"], rewards, variant, clip,
                                   ent_coef)


@pytest.mark.parametrize("variant", ["sd", "fm"])
def test_distinct_row_loss_and_gradients_match_the_replicated_row_reference(variant):
    fn = _policy(1, variant)
    rec, rewards = _fake_record(fn, B=6, n1=5, seed=2, variant=variant)
    adv = ppo.advantages_from_rewards(rewards, rec["masks"])
    loss, _ = ppo.ppo_loss(fn, rec["x"][0], rec["idx"], rec["probs"], adv, 0.2, 0.01)
    g1 = torch.autograd.grad(loss, list(fn.parameters()))
    ref = _reference_loss(fn, rec, rewards, 0.2, 0.01, variant)
    g2 = torch.autograd.grad(ref, list(fn.parameters()))
    torch.testing.assert_close(loss, ref, rtol=1e-5, atol=1e-6)
    for a, b in zip(g1, g2):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)


def test_update_side_probs_match_reference_golden():
    g = Golden("update_sd_o4_s0")
    kw = g.meta["kwargs"]
    fn = cb.FactorNetPPO(**kw)
    fn.load_state_dict(g.state_dict)
    p, e = fn(dict(x=g["x"]), g["actions"])
    torch.testing.assert_close(p, g["probs"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(e, g["entropy"], rtol=1e-5, atol=1e-6)
    g = Golden("update_fm_o2_s0_m0")
    fn = cb.FactorNetPPOFM(**g.meta["kwargs"])
    fn.load_state_dict(g.state_dict)
    p, e = fn(dict(x=g["x"]), g["actions"])
    torch.testing.assert_close(p, g["probs"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(e, g["entropy"], rtol=1e-4, atol=1e-6)


def test_flat_params_alias_module_parameters():
    fn = _policy(3)
    before = {k: v.clone() for k, v in fn.state_dict().items()}
    flat = ppo.FlatParams(fn)
    assert flat.numel == sum(p.numel() for p in fn.parameters())
    for k, v in fn.state_dict().items():
        assert torch.equal(v, before[k])
    opt = torch.optim.SGD(fn.parameters(), lr=0.1)
    rec, rewards = _fake_record(fn, 4, 3, 5)
    s0 = flat.checksum()
    stats = ppo.ppo_update(fn, flat, opt, rec, rewards, ppo_epochs=2, clip_range=0.2, entropy_coef=0.01)
    assert flat.checksum() != s0 and "loss" in stats and stats["grad_norm"] > 0
    assert fn.mlp[0].weight.data_ptr() == flat.flat.data_ptr()             # still aliased after the steps
    assert ppo.shared_step_count(7, 42) == ppo.shared_step_count(7, 42) and 2 <= ppo.shared_step_count(7, 42) <= 15


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fn = _policy(100 + rank)                      # different init per rank: the broadcast must fix that
    flat = ppo.FlatParams(fn)
    ppo.broadcast_parameters(flat, 0)
    opt = torch.optim.SGD(fn.parameters(), lr=0.05)
    for it in range(3):
        rec, rewards = _fake_record(fn, 4, 3, seed=10 * it + rank)        # each rank its own rollout
        ppo.ppo_update(fn, flat, opt, rec, rewards, ppo_epochs=1, clip_range=0.2, entropy_coef=0.01)
    q.put((rank, flat.checksum(), flat.flat.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_training_keeps_replicas_identical():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == pytest.approx(res[1][1], rel=0, abs=0)      # the reference's per-rank checksum check
    assert res[0][2] == res[1][2]


def test_graphed_rollouts_build_one_loop_per_step_count_sharing_the_live_policy(monkeypatch):
    """Host logic of ppo.GraphedRollouts without a GPU: the capture itself (denoise.GraphedDenoiseLoop) is replaced by a
    recorder.  One loop per step count, each on its OWN scheduler (own trajectory buffers) built from the prototype's
    config, all of them holding the prototype's factor_net object (the graphs must see optimizer steps) and its
    semantics switches; a rollout replays the loop of its count and hands back that loop's record."""
    from consolver_b200 import denoise

    made = []

    class FakeLoop:
        def __init__(self, scheduler, denoiser, noise, cfg, n):
            self.scheduler, self.denoiser, self.noise, self.cfg, self.n = scheduler, denoiser, noise, cfg, n
            self.replays = []
            made.append(self)

        def replay(self, noise=None):
            self.replays.append(noise)
            return ("latents", self.n)

        def record(self):
            return {"n": self.n}

    monkeypatch.setattr(denoise, "GraphedDenoiseLoop", FakeLoop)
    proto = cb.PPOScheduler(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, steps_offset=1,
                            timestep_spacing="trailing", order_dim=3, scaler_dim=1, prediction_type="v_prediction",
                            factor_net_kwargs=dict(embedding_dim=64, hidden_dim=16, num_actions=5))
    proto.reference_device, proto.use_pdl = "cpu", False
    den = object()
    noise = torch.randn(4, 8, 8)
    rolls = ppo.GraphedRollouts(proto, den, noise, batch=6, cfg=3.0, step_counts=[2, 5])
    assert sorted(rolls.loops) == [2, 5] and len(made) == 2
    for loop in made:
        s = loop.scheduler
        assert s is not proto and s.factor_net is proto.factor_net
        assert dict(s.config) == dict(proto.config)
        assert (s.reference_device, s.use_pdl, s.use_fused_rng) == ("cpu", False, proto.use_fused_rng)
        assert loop.noise.shape == (6, 4, 8, 8) and torch.equal(loop.noise[3], noise) and loop.cfg == 3.0
        assert loop.denoiser is den
    lat, rec = rolls.rollout(5)
    assert lat == ("latents", 5) and rec == {"n": 5} and made[1].replays == [None]
    other = torch.randn(4, 8, 8)
    rolls.rollout(2, noise_one=other)
    assert made[0].replays[0].shape == (6, 4, 8, 8) and torch.equal(made[0].replays[0][5], other)
    rolls.rollout(9)                                   # a count that was not prebuilt is captured on first use
    assert sorted(rolls.loops) == [2, 5, 9] and made[2].n == 9
