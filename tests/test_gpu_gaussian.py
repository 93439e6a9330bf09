"""Continuous (Gaussian) policy — ppo_type != "discrete".  EXTENSION, PARITY UNPINNED: the reference ships no source for
`FactorNetPPOContinous` (scheduler_ppo.py:23,:139), so these tests check the self-defined semantics against closed forms
(torch.distributions.Normal, torch.randn's CUDA stream) and the parts that ARE the reference's — masks, coefficient
assembly, the fused step — against the oracle."""
import math

import pytest
import torch

import consolver_oracle as orc

pytestmark = pytest.mark.gpu

PROD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000, steps_offset=1,
            timestep_spacing="trailing", use_conv=False)


def _sched(order_dim=4, scaler_dim=2, hidden=256, seed=0):
    import consolver_b200 as cb

    s = cb.PPOScheduler(order_dim=order_dim, scaler_dim=scaler_dim, ppo_type="continuous",
                        factor_net_kwargs=dict(hidden_dim=hidden), **PROD)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():                      # a non-trivial head (it starts at zero = centre of each range)
        s.factor_net.mlp[4].weight.copy_(torch.randn(s.factor_net.mlp[4].weight.shape, generator=g) * 0.3)
        s.factor_net.mlp[4].bias.add_(torch.randn(s.factor_net.mlp[4].bias.shape, generator=g) * 0.3)
    s.factor_net.cuda()
    return s


def test_constructor_builds_the_extension_and_reports_its_ranges():
    import consolver_b200 as cb

    s = cb.PPOScheduler(ppo_type="continuous", **PROD)                       # reference defaults: order 4, 2 scalers
    fn = s.factor_net
    assert isinstance(fn, cb.FactorNetPPOContinous) and fn.action_dims == 5
    assert torch.allclose(fn.action_range, torch.tensor([[0., 2.], [-2., 0.], [-1., 1.], [-.05, .05], [-.05, .05]]))
    sd = fn.state_dict()
    assert set(sd) == {"action_range", "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight",
                       "mlp.4.bias"} and sd["mlp.4.weight"].shape == (10, 256)


@pytest.mark.parametrize("B", [1, 3, 64, 1000])
def test_in_kernel_normal_draw_is_torch_randn(B):
    """RNG contract of the extension: one torch.randn([B, A]) per step from the default CUDA generator, regenerated
    inside the kernel bit for bit (Philox4x32-10, curand's Box-Muller, ATen's thread->element mapping)."""
    s = _sched()
    fn = s.factor_net
    A = fn.action_dims
    x = torch.randn(B, 4, 8, 8, device="cuda")
    s.set_timesteps(4, device="cuda")
    for i, t in enumerate(s.timesteps):
        e = torch.randn_like(x)                          # model output: drawn BEFORE the seed is set
        torch.manual_seed(100 + i)
        z_ref = torch.randn(B, A, device="cuda")
        torch.manual_seed(100 + i)
        x, actions, probs, conds, masks = s.step(e, t, x, return_dict=False)
        lp = s.last_policy()
        z = (actions - lp["mean"]) / lp["std"]
        torch.testing.assert_close(z, z_ref, rtol=1e-5, atol=1e-5)          # recovered through mean + std*z
        assert torch.equal(actions, torch.addcmul(lp["mean"].expand(B, A), lp["std"].expand(B, A), z_ref)) or \
            torch.allclose(actions, lp["mean"] + lp["std"] * z_ref, rtol=0, atol=1e-6)
    # the generator advanced exactly as the torch launch would have advanced it
    torch.manual_seed(7)
    torch.randn(B, A, device="cuda")
    want = torch.cuda.get_rng_state()
    e = torch.randn(B, 4, 8, 8, device="cuda")
    torch.manual_seed(7)
    s.set_timesteps(4, device="cuda")
    s.step(e, s.timesteps[0], torch.zeros(B, 4, 8, 8, device="cuda"), return_dict=False)
    assert torch.equal(torch.cuda.get_rng_state(), want)


@pytest.mark.parametrize("order_dim,scaler_dim", [(4, 2), (4, 0), (2, 0), (3, 1)])
def test_head_logprob_masks_and_coefficients(order_dim, scaler_dim):
    B, n = 37, 6
    s = _sched(order_dim, scaler_dim, hidden=64, seed=order_dim)
    fn = s.factor_net
    A = fn.action_dims
    s.set_timesteps(n, device="cuda")
    g = torch.Generator().manual_seed(3)
    zs = [torch.randn(B, A, generator=g) for _ in range(n)]
    s.replay = {"z": zs}
    x = torch.randn(B, 4, 8, 8, generator=g).cuda()
    ac = orc.sd_alphas_cumprod(orc.sd_betas(1000, 0.00085, 0.012, "scaled_linear"))
    hist = []
    for i, t in enumerate(s.timesteps):
        e = torch.randn(B, 4, 8, 8, generator=g)
        x_in = x
        x, actions, probs, conds, masks = s.step(e.cuda(), t, x, return_dict=False)
        lp = s.last_policy()
        # head vs torch (fp32 autograd path of the module) and vs the fp64 closed form
        mean_t, std_t = fn.mean_std(conds["x"][:1])
        torch.testing.assert_close(lp["mean"], mean_t[0], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(lp["std"], std_t[0], rtol=1e-5, atol=1e-7)
        mean64, std64 = lp["mean"].double().cpu(), lp["std"].double().cpu()
        z64 = zs[i].double()
        assert torch.allclose(actions.double().cpu(), mean64 + std64 * z64, rtol=0, atol=2e-6)
        logp64 = torch.distributions.Normal(mean64, std64).log_prob(mean64 + std64 * z64)
        assert (lp["logp"].double().cpu() - logp64).abs().max() <= 1e-6 * max(1.0, logp64.abs().max().item())
        assert torch.allclose(probs.double().cpu(), logp64.exp(), rtol=2e-6, atol=1e-9)
        # the reference's parts: masks, coefficient assembly, fused step (default = CUDA-tensor rules)
        hist = ([e] + hist)[:order_dim]
        n_hist = len(hist)
        assert torch.equal(masks.cpu(), orc.step_masks(B, A, n_hist, order_dim))
        a_cpu = actions.cpu()
        coef, scale = orc.coefficients(a_cpu, n_hist, order_dim, scaler_dim)
        c = lp["coef"].cpu()
        for j, cj in enumerate(coef or []):
            assert torch.equal(c[:, j], cj)
        for j, sj in enumerate(scale):
            assert torch.equal(c[:, order_dim + j], sj)
        eff, xs = orc.combine_history(hist, coef, scale, x_in.cpu())
        tt = int(s._timesteps_host[i])
        ref = orc.ddim_update(xs, eff, orc.ddim_scalars(ac, tt, orc.sd_prev_timestep(tt, n)), sem=orc.CUDA)
        assert torch.equal(x.cpu(), ref), f"step {i}"


def test_forced_actions_recover_the_same_logprob_and_update_side_matches():
    """PPO replay: feeding the sampled actions back gives the same log-probs; the autograd side (get_action_probs) agrees
    with the kernel on the density and returns Normal's entropy."""
    B, n = 16, 5
    s = _sched(4, 2, hidden=64, seed=9)
    fn = s.factor_net
    s.set_timesteps(n, device="cuda")
    x = torch.randn(B, 4, 8, 8, device="cuda")
    es = [torch.randn_like(x) for _ in range(n)]
    acts, logps, outs = [], [], []
    for i, t in enumerate(s.timesteps):
        x_next, actions, probs, conds, masks = s.step(es[i], t, x if i == 0 else outs[-1], return_dict=False)
        acts.append(actions.clone()); logps.append(s.last_policy()["logp"].clone()); outs.append(x_next)
        p_t, ent = fn(dict(x=conds["x"]), actions)
        torch.testing.assert_close(p_t, probs, rtol=2e-5, atol=1e-8)
        mean, std = fn.mean_std(conds["x"])
        torch.testing.assert_close(ent, 0.5 + 0.5 * math.log(2 * math.pi) + std.log(), rtol=1e-6, atol=1e-6)
    s.set_timesteps(n, device="cuda")
    s.replay = {"actions": acts}
    for i, t in enumerate(s.timesteps):
        x2 = s.step(es[i], t, x if i == 0 else outs[i - 1], return_dict=False)[0]
        assert torch.equal(x2, outs[i])
        torch.testing.assert_close(s.last_policy()["logp"], logps[i], rtol=0, atol=2e-5)
    rec = s.trajectory()
    assert rec["actions"].shape == (B, n - 1, fn.action_dims) and torch.equal(rec["actions"][:, 0], acts[1])


def test_fm_scheduler_still_refuses_the_continuous_policy():
    import consolver_b200 as cb

    with pytest.raises((NotImplementedError, AssertionError)):      # edit_ppo/scheduler_fmppo.py:169-170 is `assert 0`
        cb.FMPPOScheduler(ppo_type="continuous")
