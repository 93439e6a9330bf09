"""GPU: PPO rollout through the CUDA scheduler + torch-autograd update (SURVEY §8f N1 / BASELINE config 5)."""
import pytest
import torch

# cpu_reference: these tests check against CPU-made fixtures / the oracle's default (CPU-torch) rules; the product
# default — the reference as executed on CUDA tensors — is covered by tests/test_gpu_cuda_reference.py
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("cpu_reference")]

PROD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000,
            steps_offset=1, timestep_spacing="trailing", order_dim=4, scaler_dim=0, use_conv=False,
            factor_net_kwargs=dict(embedding_dim=64, hidden_dim=256, num_actions=11))


def _denoiser(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(4, 4, device="cuda", generator=g) * 0.3

    def f(x, t, i):                                  # a cheap stand-in with the U-Net's signature/shape
        return torch.einsum("oc,bchw->bohw", w, x) + 0.1 * torch.randn(x.shape, device="cuda", generator=g)
    return f


def test_rollout_record_is_consistent_with_the_autograd_policy_and_update_runs():
    import consolver_b200 as cb
    from consolver_b200 import ppo

    torch.manual_seed(0)
    s = cb.PPOScheduler(**PROD)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    s.factor_net.cuda()
    flat = ppo.FlatParams(s.factor_net)
    opt = torch.optim.AdamW(s.factor_net.parameters(), lr=1e-3)
    B, n = 16, 6
    noise = torch.randn(4, 16, 16, device="cuda")
    target = torch.randn(4, 16, 16, device="cuda")
    lat, rec = ppo.rollout_sd(s, _denoiser(1), noise, B, cfg=3.0, num_inference_steps=n)
    assert lat.shape == (B, 4, 16, 16) and rec["idx"].shape == (B, n - 1, 3)
    assert rec["x"].shape == (B, n - 1, 2) and rec["masks"][:, 0].tolist() == [[1.0, 0.0, 0.0]] * B
    # rows differ across samples only through the sampled actions
    assert len({tuple(r) for r in rec["idx"][:, -1].tolist()}) > 1
    # the probabilities recorded by the CUDA policy kernel are the autograd module's probabilities at those bins
    tables = s.factor_net.forward_({"x": rec["x"][0].float()})
    cur = tables.unsqueeze(0).expand(B, n - 1, 3, 11).gather(3, rec["idx"].unsqueeze(-1)).squeeze(-1)
    torch.testing.assert_close(cur, rec["probs"], rtol=0, atol=1e-6)
    torch.testing.assert_close(rec["logp"], torch.log(rec["probs"] + 1e-9), rtol=0, atol=1e-6)
    rewards = ppo.latent_mse_reward(lat, target.unsqueeze(0).expand_as(lat))
    assert rewards.shape == (B, 1)
    before = flat.checksum()
    stats = ppo.ppo_update(s.factor_net, flat, opt, rec, rewards, ppo_epochs=2, clip_range=0.2, entropy_coef=0.01)
    assert abs(stats["ratio_mean"] - 1.0) < 0.2 and stats["loss"] == stats["loss"]
    assert flat.checksum() != before
    # the next rollout sees the updated weights (tables are re-evaluated per trajectory)
    lat2, rec2 = ppo.rollout_sd(s, _denoiser(1), noise, B, cfg=3.0, num_inference_steps=n)
    t2 = s.factor_net.forward_({"x": rec2["x"][0].float()})
    cur2 = t2.unsqueeze(0).expand(B, n - 1, 3, 11).gather(3, rec2["idx"].unsqueeze(-1)).squeeze(-1)
    torch.testing.assert_close(cur2, rec2["probs"], rtol=0, atol=1e-6)


def test_graphed_preview_matches_eager_and_advances_the_generator():
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview, preview_from_pairs

    s = cb.PPOScheduler(**PROD)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    s.factor_net.cuda()
    B, n = 8, 8
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, 4, 64, 64, device="cuda", generator=g)
    pairs = [torch.randn(2 * B, 4, 64, 64, device="cuda", generator=g) for _ in range(n)]
    gp = GraphedPreview(s, x, pairs, 3.0, n)
    torch.manual_seed(77)
    out_g = gp.replay().clone()
    idx_g = gp.record()["idx"].clone()
    out_g2 = gp.replay().clone()                      # the generator advanced: different actions
    e = cb.PPOScheduler(**PROD)
    e.factor_net.load_state_dict(s.factor_net.state_dict())
    e.factor_net.cuda()
    e.set_timesteps(n, device="cuda")
    torch.manual_seed(77)
    out_e = preview_from_pairs(e, x, pairs, 3.0)
    assert torch.equal(e.trajectory()["idx"], idx_g)
    assert torch.equal(out_e, out_g)
    assert not torch.equal(out_g, out_g2)


def test_graphed_fm_preview_matches_eager():
    import numpy as np
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview, preview_from_outputs

    kw = dict(shift=3.0, use_dynamic_shifting=True, order_dim=2, scaler_dim=0, mu_dim=0,
              factor_net_kwargs=dict(hidden_dim=256, num_actions=11))
    torch.manual_seed(0)
    s, e = cb.FMPPOScheduler(**kw), cb.FMPPOScheduler(**kw)
    e.factor_net.load_state_dict(s.factor_net.state_dict())
    s.factor_net.cuda(), e.factor_net.cuda()
    B, n = 4, 8
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, 4096, 64, device="cuda", generator=g).bfloat16()
    vs = [torch.randn(B, 4096, 64, device="cuda", generator=g).bfloat16() for _ in range(n)]
    tk = dict(sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
    gp = GraphedPreview(s, x, vs, None, n, set_timesteps_kwargs=tk)
    torch.manual_seed(9)
    out_g = gp.replay().clone()
    idx_g = gp.record()["idx"].clone()
    e.set_timesteps(n, device="cuda", **tk)
    e.set_begin_index(0)
    torch.manual_seed(9)
    out_e = preview_from_outputs(e, x, vs)
    assert torch.equal(e.trajectory()["idx"], idx_g)
    assert torch.equal(out_e, out_g)


def test_graphed_preview_stress_matches_eager_over_many_replays():
    """The graph uses a side stream for the policy and PDL-chained step kernels (early loads before the previous
    step finishes): replay it many times at the bench batch size and demand bit equality with eager execution."""
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview, preview_from_pairs

    s, e = cb.PPOScheduler(**PROD), cb.PPOScheduler(**PROD)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    e.factor_net.load_state_dict(s.factor_net.state_dict())
    s.factor_net.cuda(), e.factor_net.cuda()
    B, n = 64, 8
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, 4, 64, 64, device="cuda", generator=g)
    pairs = [torch.randn(2 * B, 4, 64, 64, device="cuda", generator=g) for _ in range(n)]
    gp = GraphedPreview(s, x, pairs, 3.0, n)
    assert gp.used_chain and gp.used_policy_stream and not s.chain_steps and s.policy_stream is None
    torch.manual_seed(123)
    outs = []
    for _ in range(12):
        outs.append(gp.replay().clone())              # back-to-back: replay k+1 is enqueued while k still runs
    torch.cuda.synchronize()
    torch.manual_seed(123)
    for k in range(12):
        e.set_timesteps(n, device="cuda")
        ref = preview_from_pairs(e, x, pairs, 3.0)
        assert torch.equal(ref, outs[k]), f"replay {k}"


@pytest.mark.parametrize("variant", ["sd", "fm"])
def test_native_ppo_kernel_matches_autograd(variant):
    """csrc/ppo.cu (forward + clipped loss + entropy + backward on the distinct rows) vs torch autograd."""
    import consolver_b200 as cb
    from consolver_b200 import ppo

    torch.manual_seed(3)
    if variant == "sd":
        fn = cb.FactorNetPPO(hidden_dim=256, num_actions=11, order_dim=4, scaler_dim=0)
        with torch.no_grad():
            fn.mlp[4].weight.normal_(0, 0.3)
            fn.mlp[4].bias.normal_(0, 0.1)
        t = torch.tensor([874., 749., 624., 499., 374.])
        x_rows = torch.stack([t, t - 125], 1)
    else:
        fn = cb.FactorNetPPOFM(hidden_dim=256, num_actions=11, order_dim=4, scaler_dim=2, mu_dim=0)
        with torch.no_grad():
            fn.mlp[4].weight.mul_(0.05)          # temperature 0.01: keep the softmax away from one-hot
        x_rows = torch.rand(5, 2)
    fn.cuda()
    x_rows = x_rows.cuda()
    flat = ppo.FlatParams(fn)
    B, R, A, K = 48, x_rows.shape[0], fn.action_dims, fn.num_actions
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, K, (B, R, A), device="cuda", generator=g)
    with torch.no_grad():
        tables = fn.forward_({"x": x_rows})
    old = tables.unsqueeze(0).expand(B, R, A, K).gather(3, idx.unsqueeze(-1)).squeeze(-1)
    old = (old * (1 + 0.3 * torch.randn(old.shape, device="cuda", generator=g))).clamp(1e-4, 1.0)
    masks = torch.ones(B, R, A, device="cuda")
    masks[:, 0, 1:] = 0
    adv = ppo.advantages_from_rewards(torch.randn(B, 1, device="cuda", generator=g), masks)
    for clip, ent in ((0.2, 0.01), (0.05, 0.0)):
        flat.zero_grad()
        loss, info = ppo.ppo_loss(fn, x_rows, idx, old, adv, clip, ent)
        loss.backward()
        g_ref = flat.grad.clone()
        flat.grad.zero_()
        st = ppo.ppo_loss_grad_cuda(fn, flat, x_rows, *(t.transpose(0, 1).contiguous() for t in (idx, old, adv)), clip, ent)
        scale = g_ref.abs().max()
        assert scale > 0
        torch.testing.assert_close(flat.grad, g_ref, rtol=2e-3, atol=float(scale) * 2e-4)
        torch.testing.assert_close(st[0], loss.detach(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(st[1], info["policy_loss"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(st[2], info["entropy"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(st[3], info["ratio_mean"], rtol=1e-4, atol=1e-6)
    # deterministic: same inputs, same bits
    a = flat.grad.clone()
    ppo.ppo_loss_grad_cuda(fn, flat, x_rows, *(t.transpose(0, 1).contiguous() for t in (idx, old, adv)), 0.05, 0.0)
    assert torch.equal(a, flat.grad)


def test_native_ppo_kernel_matches_the_oracle_restatement_of_the_training_loss():
    """Kernel (distinct rows, recorded indices) vs the oracle's literal restatement of train_ppo.py:376-427 on
    B*(n-1) replicated rows with bins re-derived from the action values — loss and gradients."""
    import consolver_b200 as cb
    import consolver_oracle as orc
    from consolver_b200 import ppo

    torch.manual_seed(11)
    fn = cb.FactorNetPPO(hidden_dim=256, num_actions=11, order_dim=4, scaler_dim=0)
    with torch.no_grad():
        fn.mlp[4].weight.normal_(0, 0.3)
    B, R, A, K = 24, 6, 3, 11
    g = torch.Generator().manual_seed(2)
    t = torch.tensor([874., 749., 624., 499., 374., 249.])
    x_rows = torch.stack([t, t - 125], 1)
    idx = torch.randint(0, K, (B, R, A), generator=g)
    actions = fn.action_values[torch.arange(A).view(1, 1, A).expand(B, R, A), idx]
    old = torch.rand(B, R, A, generator=g) * 0.3 + 0.02
    masks = torch.ones(B, R, A)
    masks[:, 0, 1:] = 0
    masks[:, 1, 2:] = 0
    rewards = torch.randn(B, 1, generator=g)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k != "action_values")
          for k, v in fn.state_dict().items()}
    ref = orc.ppo_loss_replicated(sd, x_rows.unsqueeze(0).expand(B, R, 2), actions, old, masks, rewards, "sd", 0.2, 0.01)
    names = [n for n, _ in fn.named_parameters()]
    g_ref = torch.cat([gr.reshape(-1) for gr in torch.autograd.grad(ref, [sd[n] for n in names])])
    fn.cuda()
    flat = ppo.FlatParams(fn)
    adv = ppo.advantages_from_rewards(rewards, masks).cuda()
    st = ppo.ppo_loss_grad_cuda(fn, flat, x_rows.cuda(), idx.transpose(0, 1).contiguous().cuda(),
                                old.transpose(0, 1).contiguous().cuda(), adv.transpose(0, 1).contiguous(), 0.2, 0.01)
    torch.testing.assert_close(st[0].cpu(), ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(flat.grad.cpu(), g_ref, rtol=2e-3, atol=float(g_ref.abs().max()) * 2e-4)


def test_preview_pool_concurrent_replays_match_serial_execution():
    """4 independent previews in flight on 4 streams: same bits as running them one after the other."""
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview, PreviewPool, preview_from_pairs

    B, n = 32, 8
    g = torch.Generator(device="cuda").manual_seed(9)
    batches, previews, eager = [], [], []
    for j in range(4):
        s = cb.PPOScheduler(**PROD)
        with torch.no_grad():
            torch.manual_seed(50 + j)
            s.factor_net.mlp[4].weight.normal_(0, 0.05)
        s.factor_net.cuda()
        e = cb.PPOScheduler(**PROD)
        e.factor_net.load_state_dict(s.factor_net.state_dict())
        e.factor_net.cuda()
        x = torch.randn(B, 4, 64, 64, device="cuda", generator=g)
        pairs = [torch.randn(2 * B, 4, 64, 64, device="cuda", generator=g) for _ in range(n)]
        batches.append((x, pairs))
        previews.append(GraphedPreview(s, x, pairs, 3.0, n))
        eager.append(e)
    pool = PreviewPool(previews, streams=4)
    assert len(pool.streams) == 4
    torch.manual_seed(321)
    outs = []
    for rnd in range(3):
        for j in range(4):
            outs.append(pool.submit(j))
        pool.join()
        outs = [o.clone() for o in outs[:-4]] + [o.clone() for o in outs[-4:]]
    torch.cuda.synchronize()
    torch.manual_seed(321)
    k = 0
    for rnd in range(3):
        for j in range(4):
            eager[j].set_timesteps(n, device="cuda")
            ref = preview_from_pairs(eager[j], *batches[j], 3.0)
            assert torch.equal(ref, outs[k]), f"round {rnd} preview {j}"
            k += 1


@pytest.mark.parametrize("parallel", [True, False])
@pytest.mark.parametrize("kind", ["sd", "fm"])
def test_preview_group_one_graph_for_several_previews_matches_serial_execution(kind, parallel):
    """PreviewGroup: g independent previews as ONE CUDA graph (one branch each, a shared device-resident generator
    state).  Replays — back to back, and interleaved with other users of the generator — give the bits of running the g
    previews eagerly one after the other, and leave the default generator where eager execution leaves it."""
    import numpy as np
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview, PreviewGroup, preview_from_outputs, preview_from_pairs

    g_ = torch.Generator(device="cuda").manual_seed(19)
    n, G = 6, 3
    previews, eager, batches = [], [], []
    for j in range(G):
        if kind == "sd":
            B, shape = 16, (4, 32, 32)
            s, e = cb.PPOScheduler(**PROD), cb.PPOScheduler(**PROD)
            tk = None
        else:
            B, shape = 4, (256, 64)
            kw = dict(shift=3.0, use_dynamic_shifting=True, order_dim=2, scaler_dim=0, mu_dim=0,
                      factor_net_kwargs=dict(hidden_dim=64, num_actions=11))
            s, e = cb.FMPPOScheduler(**kw), cb.FMPPOScheduler(**kw)
            tk = dict(sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
        with torch.no_grad():
            torch.manual_seed(70 + j)
            s.factor_net.mlp[4].weight.normal_(0, 0.05)
        e.factor_net.load_state_dict(s.factor_net.state_dict())
        s.factor_net.cuda(), e.factor_net.cuda()
        dt = torch.float32 if kind == "sd" else torch.bfloat16
        x = torch.randn(B, *shape, device="cuda", generator=g_).to(dt)
        outs_in = [torch.randn((2 * B if kind == "sd" else B), *shape, device="cuda", generator=g_).to(dt) for _ in range(n)]
        previews.append(GraphedPreview(s, x, outs_in, 3.0 if kind == "sd" else None, n, set_timesteps_kwargs=tk))
        eager.append((e, tk))
        batches.append((x, outs_in))
    group = PreviewGroup(previews, parallel=parallel)
    assert len(group) == G

    def run_eager(j):
        e, tk = eager[j]
        e.set_timesteps(n, device="cuda", **(tk or {}))
        if kind == "sd":
            return preview_from_pairs(e, *batches[j], 3.0)
        e.set_begin_index(0)
        return preview_from_outputs(e, *batches[j])

    torch.manual_seed(77)
    got = []
    for rnd in range(3):
        got.append([o.clone() for o in group.replay()])
        if rnd == 1:
            torch.rand(5, device="cuda")                 # somebody else draws: the group must notice and re-seed its state
    single = previews[1].replay().clone()                # the member previews stay usable on their own
    end_state = torch.cuda.get_rng_state()
    torch.manual_seed(77)
    for rnd in range(3):
        for j in range(G):
            assert torch.equal(run_eager(j), got[rnd][j]), f"round {rnd} preview {j}"
        if rnd == 1:
            torch.rand(5, device="cuda")
    assert torch.equal(run_eager(1), single)
    assert torch.equal(torch.cuda.get_rng_state(), end_state)


def test_two_preview_groups_in_rotation_on_two_streams_match_serial_execution():
    """the bench's launch form: groups replayed round-robin, one stream each (PreviewPool of PreviewGroups), device
    generator states advancing by the whole rotation — same bits and same generator consumption as serial eager runs"""
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview, PreviewGroup, PreviewPool, preview_from_pairs

    gen = torch.Generator(device="cuda").manual_seed(5)
    n, B = 8, 16
    previews, eager, batches = [], [], []
    for j in range(4):
        s, e = cb.PPOScheduler(**PROD), cb.PPOScheduler(**PROD)
        with torch.no_grad():
            torch.manual_seed(90 + j)
            s.factor_net.mlp[4].weight.normal_(0, 0.05)
        e.factor_net.load_state_dict(s.factor_net.state_dict())
        s.factor_net.cuda(), e.factor_net.cuda()
        x = torch.randn(B, 4, 32, 32, device="cuda", generator=gen)
        pairs = [torch.randn(2 * B, 4, 32, 32, device="cuda", generator=gen) for _ in range(n)]
        previews.append(GraphedPreview(s, x, pairs, 3.0, n))
        eager.append(e)
        batches.append((x, pairs))
    groups = [PreviewGroup(previews[:2], rotation=2, parallel=False), PreviewGroup(previews[2:], rotation=2, parallel=False)]
    pool = PreviewPool(groups, streams=2)
    assert len(pool.streams) == 2
    torch.manual_seed(11)
    got = []
    for rnd in range(3):
        for gi in range(2):
            outs = pool.submit(gi)
            pool.join() if rnd == 2 else None
            got.append(outs)
        if rnd < 2:
            pool.join()
        got = got[:-2] + [[o.clone() for o in outs] for outs in got[-2:]]
    torch.cuda.synchronize()
    end_state = torch.cuda.get_rng_state()
    torch.manual_seed(11)
    k = 0
    for rnd in range(3):
        for gi in range(2):
            for m in range(2):
                j = 2 * gi + m
                eager[j].set_timesteps(n, device="cuda")
                assert torch.equal(preview_from_pairs(eager[j], *batches[j], 3.0), got[k][m]), f"round {rnd} preview {j}"
            k += 1
    assert torch.equal(torch.cuda.get_rng_state(), end_state)
    assert groups[0]._k <= 1 and groups[1]._k <= 1      # steady rotation: one initial state refresh per group, then none


def test_graphed_denoise_loop_with_a_denoiser_matches_eager():
    """Whole CFG loop (denoiser included) in one CUDA graph == eager denoise_loop, bit for bit, seeds included."""
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedDenoiseLoop, denoise_loop

    s, e = cb.PPOScheduler(**PROD), cb.PPOScheduler(**PROD)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    e.factor_net.load_state_dict(s.factor_net.state_dict())
    s.factor_net.cuda(), e.factor_net.cuda()
    @torch.no_grad()
    def den(x, t, i):      # element-wise only: library convolutions may pick different algorithms under capture
        scale = 1.0 + 0.0005 * float(t)      # python scalar in both modes (eager hands a 0-d tensor, the graph an int)
        return 0.6 * x + 0.3 * torch.roll(x, 1, dims=2) - 0.2 * torch.roll(x, 1, dims=1) * scale

    noise = torch.randn(2, 4, 32, 32, device="cuda")
    noise2 = torch.randn_like(noise)                 # drawn up front: nothing else may touch the generator below
    g = GraphedDenoiseLoop(s, den, noise, cfg=3.0, num_inference_steps=6)
    torch.manual_seed(4)
    out1 = g.replay().clone()
    idx1 = g.record()["idx"].clone()
    out2 = g.replay(noise2).clone()
    torch.manual_seed(4)
    ref1, rec1 = denoise_loop(e, den, noise, cfg=3.0, num_inference_steps=6)
    assert torch.equal(rec1["idx"], idx1) and torch.equal(ref1, out1)
    ref2, _ = denoise_loop(e, den, noise2, cfg=3.0, num_inference_steps=6)
    assert torch.equal(ref2, out2)


def test_graphed_rollouts_equal_eager_rollouts_and_see_weight_updates():
    """GraphedRollouts (one CUDA graph of the whole sampling loop per step count, schedulers sharing one factor_net) vs
    eager ppo.rollout_sd: same latents, same record, same generator consumption — also after optimizer steps between
    the rollouts (the graphs re-evaluate the probability tables from the live weights) and with the sync-free update."""
    import consolver_b200 as cb
    from consolver_b200 import ppo

    def make():
        torch.manual_seed(0)
        s = cb.PPOScheduler(**PROD)
        with torch.no_grad():
            s.factor_net.mlp[4].weight.normal_(0, 0.05)
        s.factor_net.cuda()
        flat = ppo.FlatParams(s.factor_net)
        return s, flat, torch.optim.AdamW(s.factor_net.parameters(), lr=1e-2)

    w = torch.randn(4, 4, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 0.3
    den = lambda x, t, i: torch.einsum("oc,bchw->bohw", w, x)  # noqa: E731
    g = torch.Generator(device="cuda").manual_seed(4)
    noise, target = (torch.randn(4, 16, 16, device="cuda", generator=g) for _ in range(2))
    B, counts = 12, [5, 2, 9, 5, 3, 9]
    (s_g, flat_g, opt_g), (s_e, flat_e, opt_e) = make(), make()
    assert torch.equal(flat_g.flat, flat_e.flat)
    rolls = ppo.GraphedRollouts(s_g, den, noise, B, 3.0, step_counts=sorted(set(counts)))
    outs = {}
    for mode, s, flat, opt in (("graph", s_g, flat_g, opt_g), ("eager", s_e, flat_e, opt_e)):
        torch.manual_seed(99)
        res = []
        for n in counts:
            lat, rec = rolls.rollout(n) if mode == "graph" else ppo.rollout_sd(s, den, noise, B, 3.0, n)
            res.append((lat.clone(), {k: v.clone() for k, v in rec.items()}))
            r = ppo.latent_mse_reward(lat, target.unsqueeze(0).expand_as(lat))
            st = ppo.ppo_update(s.factor_net, flat, opt, rec, r, ppo_epochs=2, entropy_coef=0.01,
                                read_back=mode == "eager")
            if mode == "graph":
                assert isinstance(st["stats"], torch.Tensor) and st["stats"].is_cuda and st["grad_norm"].is_cuda
        outs[mode] = (res, flat.flat.clone(), torch.cuda.get_rng_state())
    for k, ((lat_g, rec_g), (lat_e, rec_e)) in enumerate(zip(outs["graph"][0], outs["eager"][0])):
        assert torch.equal(lat_g, lat_e), f"rollout {k} (n={counts[k]}): latents"
        for key in ("idx", "actions", "probs", "masks", "x"):
            assert torch.equal(rec_g[key], rec_e[key]), f"rollout {k} (n={counts[k]}): {key}"
    assert torch.equal(outs["graph"][1], outs["eager"][1]), "weights after the updates"
    assert torch.equal(outs["graph"][2], outs["eager"][2]), "default generator consumed differently"
