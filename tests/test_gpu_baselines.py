"""GPU: baseline solvers (DDIM, Adams-Bashforth multistep, FM Euler) through the same fused kernels vs closed forms."""
import numpy as np
import pytest
import torch

import consolver_oracle as orc
from golden_io import Golden, names as golden_names

pytestmark = pytest.mark.gpu

SD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, steps_offset=1, timestep_spacing="trailing")


def test_ddim_and_multistep_match_closed_form():
    from consolver_b200 import baselines

    g = torch.Generator().manual_seed(0)
    B, shape, n = 3, (4, 16, 16), 6
    x0 = torch.randn(B, *shape, generator=g)
    pairs = [torch.randn(2 * B, *shape, generator=g) for _ in range(n)]
    ac = orc.sd_alphas_cumprod(orc.sd_betas(1000, 0.00085, 0.012, "scaled_linear"))
    for order, solver in ((1, baselines.ddim_solver(**SD)), (4, baselines.multistep_solver(4, **SD)),
                          (2, baselines.multistep_solver(2, **SD))):
        solver.set_timesteps(n, device="cuda")
        x_gpu, x_cpu, hist = x0.cuda(), x0, []
        for i, t in enumerate(solver.timesteps.tolist()):
            out = solver.step_cfg(pairs[i].cuda(), t, x_gpu, 3.0)
            assert out[1] is None and out[2] is None
            x_gpu = out[0]
            u, c = pairs[i].chunk(2)
            hist = ([orc.cfg_combine(u, c, 3.0)] + hist)[:order]
            w = baselines.ADAMS_BASHFORTH[min(len(hist), order)]
            coef = None if len(hist) == 1 else [torch.full((B,), float(np.float32(v))) for v in w] + \
                [torch.zeros(B)] * (len(hist) - len(w))
            eff, _ = orc.combine_history(hist, coef, [], x_cpu)
            x_cpu = orc.ddim_update(x_cpu, eff, orc.ddim_scalars(ac, t, orc.sd_prev_timestep(t, n)))
            assert torch.equal(x_gpu.cpu(), x_cpu), f"order {order} step {i}"


def test_flow_euler_matches_reference_formula():
    from consolver_b200 import baselines

    g = torch.Generator().manual_seed(1)
    B, shape, n = 2, (64, 16), 5
    s = baselines.flow_euler_solver(shift=3.0)
    s.set_timesteps(n, device="cuda")
    s.set_begin_index(0)
    x = torch.randn(B, *shape, generator=g).bfloat16()
    x_gpu, x_cpu = x.cuda(), x
    sig = s.sigmas.cpu()
    for i, t in enumerate(s.timesteps):
        v = torch.randn(B, *shape, generator=g).bfloat16()
        x_gpu = s.step(v.cuda(), t, x_gpu, return_dict=False)[0]
        x_cpu = (x_cpu.float() + (sig[i + 1] - sig[i]) * v).to(torch.bfloat16)      # edit_ppo/scheduler_fm.py:405-410
        assert torch.equal(x_gpu.cpu(), x_cpu), f"step {i}"


# ---- the reference's own flow-matching baselines (edit_ppo/scheduler_fm.py:384-488), pinned by golden vectors ------
def _fmgen(g, dev="cuda"):
    import consolver_b200 as cb
    m = g.meta
    s = cb.FlowMatchGeneralDiscreteScheduler(**m["config"])
    if m["config"]["use_dynamic_shifting"]:
        s.set_timesteps(m["n"], device=dev, sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
    else:
        s.set_timesteps(m["n"], device=dev)
    if m["use_begin_index"]:
        s.set_begin_index(0)
    return s


@pytest.mark.parametrize("name", golden_names("fmgen_"))
def test_fm_baseline_schedulers_bit_exact_on_golden(name):
    """euler / heun / dpm-solver / dpm-solver-multistep through the fused step kernel: every latent of every step
    identical to the unmodified reference's, fp32, fp16 and bf16 (the 16-bit cases exercise torch's evaluation of
    `dt * v` and `0.5 * dt * (v1 + v2)` in the model dtype — CONSOLVER_FLAG_LOWP_COMBINE)."""
    g = Golden(name)
    s = _fmgen(g)
    x = g["x_T"].cuda()
    for i, t in enumerate(s.timesteps):
        out = s.step(g[f"v_{i}"].cuda(), t, x)
        x = out.prev_sample
        ref = g[f"prev_{i}"]
        assert x.dtype == ref.dtype
        assert torch.equal(x.cpu(), ref), f"{name} step {i}: latent not bit-identical"
    assert s.step_index == g.meta["n"]


@pytest.mark.parametrize("kind", ["euler", "heun", "dpm-solver", "dpm-solver-multistep"])
def test_fm_baseline_schedulers_full_size_against_oracle(kind):
    """FLUX shape [B, 4096, 64] bf16, 8 steps, against the CPU oracle's restatement; also the tuple return and the
    second destination."""
    import consolver_b200 as cb
    kw = dict(shift=3.0, use_dynamic_shifting=True)
    s = cb.FlowMatchGeneralDiscreteScheduler(type=kind, **kw)
    o = orc.OracleFMGeneralScheduler(kind=kind, **kw)
    sig = np.linspace(1.0, 1 / 8, 8)
    s.set_timesteps(8, device="cuda", sigmas=sig, mu=1.15)
    o.set_timesteps(8, sigmas=sig, mu=1.15)
    s.set_begin_index(0)
    o.set_begin_index(0)
    gen = torch.Generator().manual_seed(5)
    x_ref = torch.randn(2, 4096, 64, generator=gen).bfloat16()
    x = x_ref.cuda()
    wide = torch.zeros(2, 4096 + 512, 64, device="cuda", dtype=torch.bfloat16)     # [latents | image latents]
    for i, t in enumerate(s.timesteps):
        v = torch.randn(2, 4096, 64, generator=gen).bfloat16()
        (x,) = s.step(v.cuda(), t, x, return_dict=False, out2=wide)
        x_ref = o.step(v, o.timesteps[i], x_ref)
        assert torch.equal(x.cpu(), x_ref), f"{kind} step {i}"
        assert torch.equal(wide[:, :4096], x) and not wide[:, 4096:].any()


def test_fm_baseline_second_stage_without_first_is_an_error():
    import consolver_b200 as cb
    s = cb.FlowMatchGeneralDiscreteScheduler(type="heun")
    s.set_timesteps(4, device="cuda")
    s.set_begin_index(1)
    v = torch.zeros(1, 8, 8, device="cuda")
    with pytest.raises(RuntimeError, match="first stage"):
        s.step(v, s.timesteps[1], v)
    bad = cb.FlowMatchGeneralDiscreteScheduler(type="rk4")
    bad.set_timesteps(4, device="cuda")
    with pytest.raises(ValueError, match="unknown solver type"):
        bad.step(v, bad.timesteps[0], v)
