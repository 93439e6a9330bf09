"""Host-side schedules of the drop-in schedulers against the UNMODIFIED reference, for option combinations the
committed golden vectors do not cover.  Runs only where the reference tree exists (the build container);
skipped elsewhere — the GPU box never reads /root/reference."""
import itertools

import numpy as np
import pytest
import torch

import consolver_b200 as cb
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("spacing,schedule,n", list(itertools.product(
    ["leading", "trailing", "linspace"], ["linear", "scaled_linear", "squaredcos_cap_v2"], [2, 7, 15, 50])))
def test_sd_schedules_match_reference(spacing, schedule, n):
    ref = ref_shim.load_reference()
    kw = dict(timestep_spacing=spacing, beta_schedule=schedule, steps_offset=1, order_dim=4, scaler_dim=0,
              factor_net_kwargs=dict(hidden_dim=8, num_actions=3))
    with ref_shim.quiet():
        r = ref.PPOScheduler(**kw)
    m = cb.PPOScheduler(**kw)
    assert torch.equal(r.alphas_cumprod, m.alphas_cumprod)
    r.set_timesteps(n)
    m.set_timesteps(n)
    assert torch.equal(r.timesteps, m.timesteps)
    x0, noise, t = torch.randn(2, 4, 4, 4), torch.randn(2, 4, 4, 4), torch.tensor([3, 700])
    assert torch.equal(r.add_noise(x0, noise, t), m.add_noise(x0, noise, t))
    # the per-step scalars the kernel receives are the reference's 0-d fp32 values
    for tt in m.timesteps.tolist():
        a = r.alphas_cumprod[tt]
        assert float(a ** 0.5) == float(m._sqrt_abar[tt]) and float((1 - a) ** 0.5) == float(m._sqrt_1m_abar[tt])


FM_VARIANTS = [dict(shift=3.0), dict(use_dynamic_shifting=True), dict(use_karras_sigmas=True),
               dict(use_exponential_sigmas=True), dict(use_beta_sigmas=True), dict(shift=2.0, shift_terminal=0.02),
               dict(invert_sigmas=True), dict(use_dynamic_shifting=True, time_shift_type="linear")]


@pytest.mark.parametrize("variant", FM_VARIANTS)
@pytest.mark.parametrize("mode", ["n", "sigmas", "timesteps"])
def test_fm_schedules_match_reference(variant, mode):
    ref = ref_shim.load_reference()
    kw = dict(order_dim=2, scaler_dim=0, mu_dim=0, factor_net_kwargs=dict(hidden_dim=8, num_actions=3), **variant)
    with ref_shim.quiet():
        r = ref.FMPPOScheduler(**kw)
    m = cb.FMPPOScheduler(**kw)
    assert torch.equal(r.sigmas, m.sigmas) and torch.equal(r.timesteps, m.timesteps)
    assert r.sigma_min == m.sigma_min and r.sigma_max == m.sigma_max
    n = 6
    mu = 0.9 if variant.get("use_dynamic_shifting") else None
    args = dict(n=dict(num_inference_steps=n), sigmas=dict(sigmas=np.linspace(1.0, 1 / n, n)),
                timesteps=dict(timesteps=[900.0, 700.0, 500.0, 300.0, 200.0, 100.0]))[mode]
    r.set_timesteps(mu=mu, **args)
    m.set_timesteps(mu=mu, **args)
    assert torch.equal(r.sigmas, m.sigmas), (r.sigmas, m.sigmas)
    assert torch.equal(r.timesteps, m.timesteps)
    assert r.index_for_timestep(r.timesteps[3]) == m.index_for_timestep(m.timesteps[3])
    lat, noise = torch.randn(2, 8, 4), torch.randn(2, 8, 4)
    assert torch.equal(r.scale_noise(lat, r.timesteps[:2], noise), m.scale_noise(lat, m.timesteps[:2], noise))


def test_policy_modules_match_reference_at_init_and_on_the_update_side():
    ref = ref_shim.load_reference()
    for Ref, Mine, kw in ((ref.FactorNetPPO_SD, cb.FactorNetPPO, dict(hidden_dim=32, num_actions=11, order_dim=4, scaler_dim=2)),
                          (ref.FactorNetPPO_FM, cb.FactorNetPPOFM, dict(hidden_dim=32, num_actions=11, order_dim=3, scaler_dim=1, mu_dim=1)),
                          (ref.FactorNetPPO_SD, cb.FactorNetPPO, dict(hidden_dim=32, num_actions=7, order_dim=4, scaler_dim=0, use_conv=True))):
        with ref_shim.quiet():
            torch.manual_seed(5)
            r = Ref(**kw)
        torch.manual_seed(5)
        m = Mine(**kw)
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs) == list(ms) and all(torch.equal(rs[k], ms[k]) for k in rs)    # same init stream, same buffers
        with torch.no_grad():
            for mod in (r, m):
                torch.manual_seed(9)
                mod.mlp[4].weight.normal_(0, 0.2)
        x = torch.tensor([[874.0, 749.0], [499.0, 374.0]]) if Ref is ref.FactorNetPPO_SD else torch.rand(2, 2)
        d = {"x": x}
        if kw.get("use_conv"):
            d["epsilon"] = torch.randn(2, 4, 4, 8, 8)
        idx = torch.randint(0, kw["num_actions"], (2, r.action_dims))
        actions = r.action_values[torch.arange(r.action_dims), idx]
        with ref_shim.quiet():
            pr, er = r(d, actions)
        pm, em = m(d, actions)
        torch.testing.assert_close(pm, pr, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(em, er, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("variant", FM_VARIANTS)
@pytest.mark.parametrize("kind", ["euler", "heun", "dpm-solver", "dpm-solver-multistep"])
def test_fm_baseline_scheduler_grids_match_reference(variant, kind):
    """FlowMatchGeneralDiscreteScheduler (edit_ppo/scheduler_fm.py): every sigma-grid option against the live class."""
    ref = ref_shim.load_reference()
    r = ref.FlowMatchGeneralDiscreteScheduler(type=kind, **variant)
    m = cb.FlowMatchGeneralDiscreteScheduler(type=kind, **variant)
    assert torch.equal(r.sigmas, m.sigmas) and torch.equal(r.timesteps, m.timesteps)
    mu = 0.9 if variant.get("use_dynamic_shifting") else None
    r.set_timesteps(7, mu=mu)
    m.set_timesteps(7, mu=mu)
    assert torch.equal(r.sigmas, m.sigmas) and torch.equal(r.timesteps, m.timesteps)
    assert r.index_for_timestep(r.timesteps[2]) == m.index_for_timestep(m.timesteps[2])
    lat, noise = torch.randn(2, 8, 4), torch.randn(2, 8, 4)
    assert torch.equal(r.scale_noise(lat, r.timesteps[:2], noise), m.scale_noise(lat, m.timesteps[:2], noise))


AMED_ALL = {   # gen_ppo.py:24-55
    4: ([999, 694, 500, 110, 0], [1.0, 0.991, 1.0, 0.9912, 1.0], [1.0, 1.0333, 1.0, 0.9861, 1.0]),
    6: ([999, 758, 666, 495, 333, 107, 0], [1.0, 0.9924, 1.0, 0.9916, 1.0, 0.9906, 1.0],
        [1.0, 1.052, 1.0, 0.9998, 1.0, 0.9781, 1.0]),
    8: ([999, 831, 749, 623, 500, 394, 250, 88, 0], [1.0, 0.9976, 1.0, 0.991, 1.0, 0.9907, 1.0, 0.9905, 1.0],
        [1.0, 1.0257, 1.0, 0.9989, 1.0, 1.0022, 1.0, 0.9747, 1.0]),
    10: ([999, 885, 799, 705, 599, 492, 400, 329, 200, 73, 0],
         [1.0, 0.9974, 1.0, 0.9904, 1.0, 0.991, 1.0, 0.9905, 1.0, 0.9904, 1.0],
         [1.0, 0.9872, 1.0, 1.0152, 1.0, 1.0186, 1.0, 0.9934, 1.0, 0.9731, 1.0]),
    14: ([999, 924, 856, 790, 714, 623, 571, 494, 428, 374, 285, 241, 143, 55, 0],
         [1.0, 0.9922, 1.0, 0.9909, 1.0, 0.9914, 1.0, 0.9908, 1.0, 0.9904, 1.0, 0.9903, 1.0, 0.9904, 1.0],
         [1.0, 0.9835, 1.0, 1.0293, 1.0, 1.0216, 1.0, 1.0241, 1.0, 1.0021, 1.0, 0.9844, 1.0, 0.9714, 1.0]),
}


@pytest.mark.parametrize("solver_order", [2, 3])
@pytest.mark.parametrize("n", sorted(AMED_ALL))
def test_amed_grids_and_step_scalars_match_the_plugin_for_every_shipped_schedule(n, solver_order):
    """All five AMED schedules of gen_ppo.py through the unmodified plugin (over the stand-in of its diffusers base):
    time-scaled timesteps, sigmas, and — via the kernel's documented arithmetic — every latent of a full run."""
    from test_host_cpu import _dpm_kernel_arithmetic
    ref = ref_shim.load_reference()
    cfg = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1,
               solver_order=solver_order)
    ts, dirs, times = AMED_ALL[n]
    r = ref.AMEDDPMSolverMultistepScheduler(**cfg)
    m = cb.DPMSolverMultistepScheduler(**cfg)
    for s in (r, m):
        s.scale_dirs, s.scale_times = dirs, times
        s.set_timesteps(n, timesteps=ts)
    assert torch.equal(r.timesteps, m.timesteps) and torch.equal(r.sigmas, m.sigmas)
    assert r.num_inference_steps == m.num_inference_steps
    g = torch.Generator().manual_seed(n)
    x = torch.randn(2, 4, 8, 8, generator=g)
    xm, hist = x, []
    for t in r.timesteps:
        e = torch.randn(2, 4, 8, 8, generator=g)
        x = r.step(e, t, x, return_dict=False)[0]
        _, plan, order = m._plan_for_step(t)
        xm, hist = _dpm_kernel_arithmetic(plan, order, e, xm, hist)
        m._advance()
        assert torch.equal(x, xm)
