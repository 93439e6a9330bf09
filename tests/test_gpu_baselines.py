"""GPU: baseline solvers (DDIM, Adams-Bashforth multistep, FM Euler) through the same fused kernels vs closed forms."""
import numpy as np
import pytest
import torch

import consolver_oracle as orc
from golden_io import Golden, names as golden_names

# cpu_reference: these tests check against CPU-made fixtures / the oracle's default (CPU-torch) rules; the product
# default — the reference as executed on CUDA tensors — is covered by tests/test_gpu_cuda_reference.py
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("cpu_reference")]

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
        (x,) = s.step(v.cuda(), t, x, return_dict=False, out2=wide[:, :4096])
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


# ---- AMED-scaled multistep DPM-Solver(++) (diffusers_amed_plugin_dpmpp.py), one fused kernel per step --------------
def _amed(g, dev="cuda"):
    import consolver_b200 as cb
    m = g.meta
    s = cb.DPMSolverMultistepScheduler(**m["config"])
    if m["amed"]:
        s.scale_dirs, s.scale_times = m["scale_dirs"], m["scale_times"]
        s.set_timesteps(m["n"], device=dev, timesteps=m["schedule"])
    else:
        s.set_timesteps(m["n"], device=dev)
    return s


@pytest.mark.parametrize("name", golden_names("amed_"))
def test_amed_dpm_solver_bit_exact_on_golden(name):
    """Every latent of every step identical to the plugin's (fp32): first / second order, midpoint / heun,
    dpmsolver / dpmsolver++, epsilon / sample / v-prediction, AMED and stock grids, ragged sizes."""
    g = Golden(name)
    s = _amed(g)
    assert torch.equal(s.timesteps.cpu(), g["timesteps"])
    x = g["x_T"].cuda()
    for i, t in enumerate(s.timesteps):          # CUDA scalars, as a pipeline passes them
        x = s.step(g[f"eps_{i}"].cuda(), t, x).prev_sample
        assert torch.equal(x.cpu(), g[f"prev_{i}"]), f"{name} step {i}: latent not bit-identical"


def _amed_pair(kind="amed", **over):
    import consolver_b200 as cb
    cfg = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1, **over)
    ts = [999, 831, 749, 623, 500, 394, 250, 88, 0]
    dirs = [1.0, 0.9976, 1.0, 0.991, 1.0, 0.9907, 1.0, 0.9905, 1.0]
    times = [1.0, 1.0257, 1.0, 0.9989, 1.0, 1.0022, 1.0, 0.9747, 1.0]
    s = cb.DPMSolverMultistepScheduler(**cfg)
    s.scale_dirs, s.scale_times = dirs, times
    s.set_timesteps(8, device="cuda", timesteps=ts)
    o = orc.OracleDPMSolverAMED(scale_dirs=dirs, scale_times=times, **cfg)
    o.set_timesteps(8, timesteps=ts)
    return s, o


def test_amed_step_cfg_full_size_matches_oracle_and_feeds_next_input():
    """SD1.5 latent shape, B=8, 8 AMED steps, CFG fused; the next latent also lands in both halves of the next
    [2B] denoiser input (out2 written twice is the caller's business: here one half)."""
    s, o = _amed_pair()
    gen = torch.Generator().manual_seed(11)
    B, shape = 8, (4, 64, 64)
    x_ref = torch.randn(B, *shape, generator=gen)
    x = x_ref.cuda()
    nxt = torch.zeros(2 * B, *shape, device="cuda")
    for i, t in enumerate(s.timesteps):
        pair = torch.randn(2 * B, *shape, generator=gen)
        (x,) = s.step_cfg(pair.cuda(), t, x, 7.5, out2=nxt[B:])
        u, c = pair.chunk(2)
        x_ref = o.step(orc.cfg_combine(u, c, 7.5), o.timesteps[i], x_ref)
        assert torch.equal(x.cpu(), x_ref), f"step {i}"
        assert torch.equal(nxt[B:], x) and not nxt[:B].any()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_amed_16bit_io_tracks_fp32(dtype):
    """16-bit model outputs / latents: fp32 arithmetic in registers, e, m0 and x' rounded once each.  The plugin's
    own 16-bit path rounds after every torch op, so this is compared to the fp32 oracle on the same (rounded)
    inputs within the dtype's resolution, not bit for bit."""
    s, o = _amed_pair()
    gen = torch.Generator().manual_seed(12)
    x_ref = torch.randn(2, 4, 32, 32, generator=gen).to(dtype).float()
    x = x_ref.to(dtype).cuda()
    eps_rel = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    for i, t in enumerate(s.timesteps):
        e = torch.randn(2, 4, 32, 32, generator=gen).to(dtype)
        x = s.step(e.cuda(), t, x, return_dict=False)[0]
        assert x.dtype == dtype
        x_ref = o.step(e.float(), o.timesteps[i], x_ref)
        err = (x.float().cpu() - x_ref).abs().max() / x_ref.abs().max()
        assert err < 6 * eps_rel * (i + 1), f"step {i}: {err}"
        x_ref = x.float().cpu()                 # re-anchor: compare single steps, not accumulated drift
