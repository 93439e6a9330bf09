"""LIVE differential tests on the GPU: the unmodified reference (byte-identical copies staged under oracle/_ref by
oracle/stage_ref.py — they travel with the working tree) and the drop-in run SIDE BY SIDE on the same device, same seeds,
same inputs.  Beyond the committed fixtures this covers (a) the reference's own caller loop `denoise_ppo.denoise_diffusion`
driving either scheduler — the drop-in claim at the call site — and (b) randomly drawn configurations.
Skipped when oracle/_ref is absent (nothing here reads /root/reference)."""
import importlib.util
import os
import random
import types

import numpy as np
import pytest
import torch

import ref_shim

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shim.reference_available(),
                                                  reason="oracle/_ref not staged (python oracle/stage_ref.py)")]

SD_PROD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000,
               steps_offset=1, timestep_spacing="trailing", use_conv=False)


def _seed_policy(fn, seed, last_std):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in fn.named_parameters():
            if name.startswith("mlp.4"):
                p.copy_(torch.randn(p.shape, generator=g) * (last_std if name.endswith("weight") else 0.1))
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.3)


def _pair(kind, seed, last_std=0.5, **cfg):
    """(reference scheduler, drop-in scheduler) with identical weights, both on cuda"""
    import consolver_b200 as cb

    ref = ref_shim.load_reference()
    fkw = dict(hidden_dim=64, num_actions=11)
    with ref_shim.quiet():
        if kind == "sd":
            r = ref.PPOScheduler(factor_net_kwargs=dict(embedding_dim=64, **fkw), **cfg)
            o = cb.PPOScheduler(factor_net_kwargs=dict(embedding_dim=64, **fkw), **cfg)
        else:
            r = ref.FMPPOScheduler(factor_net_kwargs=dict(fkw), **cfg)
            o = cb.FMPPOScheduler(factor_net_kwargs=dict(fkw), **cfg)
    _seed_policy(r.factor_net, seed, last_std)
    o.factor_net.load_state_dict(r.factor_net.state_dict())
    r.factor_net.cuda(), o.factor_net.cuda()
    return r, o


def _load_caller():
    path = os.path.join(ref_shim.REFERENCE_ROOT, "denoise_ppo.py")
    if not os.path.isfile(path):
        pytest.skip("denoise_ppo.py not staged")
    spec = importlib.util.spec_from_file_location("_ref_denoise_ppo", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Tok:
    model_max_length = 8

    def __call__(self, text, **kw):
        ids = torch.tensor([[(len(t) * 7 + j) % 50 for j in range(self.model_max_length)] for t in text])
        return types.SimpleNamespace(input_ids=ids)


def _text_encoder(ids):
    return (torch.sin(ids.float().unsqueeze(-1) * torch.arange(1, 5, device=ids.device)),)


def _unet(x, t, encoder_hidden_states=None, return_dict=False):
    """deterministic element-wise stand-in for the U-Net (not the product): depends on the latent, the timestep and the
    text embedding, so the two CFG halves differ"""
    c = encoder_hidden_states.mean(dim=(1, 2)).view(-1, 1, 1, 1)
    return (0.55 * x + 0.25 * torch.roll(x, 1, dims=3) - 0.15 * torch.roll(x, 1, dims=1) + 0.1 * c +
            0.0003 * t.float(),)


@pytest.mark.parametrize("pred,scaler_dim,n", [("epsilon", 0, 8), ("v_prediction", 2, 6), ("epsilon", 1, 15)])
def test_the_references_own_caller_loop_gives_the_same_rollout_with_either_scheduler(pred, scaler_dim, n):
    """denoise_ppo.denoise_diffusion (UNMODIFIED) with the reference's PPOScheduler and with the drop-in: final latents,
    the whole rollout record (conds x / epsilon, actions, masks) bit-identical, probabilities within the measured
    cuBLAS-vs-kernel spread."""
    caller = _load_caller()
    r, o = _pair("sd", seed=n, order_dim=4, scaler_dim=scaler_dim, prediction_type=pred, **SD_PROD)
    noise = torch.randn(5, 4, 16, 16, generator=torch.Generator().manual_seed(1)).cuda()
    text = [f"prompt {i}" * (i + 1) for i in range(5)]
    outs = []
    for sched in (r, o):
        torch.manual_seed(1234)
        with ref_shim.quiet(), torch.no_grad():
            outs.append(caller.denoise_diffusion(_text_encoder, sched, _unet, noise, text, _Tok(), cfg=3.0,
                                                 num_inference_steps=n))
    (lat_r, conds_r, probs_r, act_r, masks_r, _), (lat_o, conds_o, probs_o, act_o, masks_o, _) = outs
    assert torch.equal(act_r, act_o), "sampled actions differ"
    assert torch.equal(masks_r, masks_o) and torch.equal(conds_r["x"], conds_o["x"])
    assert torch.equal(conds_r["epsilon"], conds_o["epsilon"])
    torch.testing.assert_close(probs_o, probs_r, rtol=0, atol=1e-6)
    assert lat_o.dtype == lat_r.dtype and torch.equal(lat_r, lat_o), "final latents differ"


def _random_sd_case(rng):
    od = rng.choice([2, 3, 4, 4])
    return dict(order_dim=od, scaler_dim=rng.choice([0, 0, 1, 2]),
                prediction_type=rng.choice(["epsilon", "epsilon", "v_prediction"]),
                timestep_spacing=rng.choice(["trailing", "leading", "linspace"]),
                beta_schedule=rng.choice(["scaled_linear", "linear", "squaredcos_cap_v2"]),
                beta_start=0.00085, beta_end=0.012, steps_offset=rng.choice([0, 1]), use_conv=False)


@pytest.mark.parametrize("case", range(32))
def test_random_sd_configurations_step_for_step(case):
    """randomly drawn scheduler configs, step counts, batch / latent shapes and dtype flows (fp32; fp16 / bf16 outputs
    with fp32 or 16-bit latents; the autocast rollout; gen_ppo.py's fp16 policy under autocast): every step bit-identical
    to the reference stepping next to it on the same GPU with the same seed"""
    rng = random.Random(1000 + case)
    cfg = _random_sd_case(rng)
    n = rng.choice([2, 3, 5, 8, 13])
    B = rng.choice([1, 2, 3, 7])
    # 4-D latents only: the reference's coefficient tensors are [B,1,1,1] (scheduler_ppo.py:253)
    shape = rng.choice([(4, 8, 8), (3, 5, 7), (4, 16, 16), (1, 1, 33)])
    flow = rng.choice(["f32", "f32", "f16_out", "bf16_out", "f16_pipeline", "autocast_f16", "genppo_f16", "genppo_bf16"])
    r, o = _pair("sd", seed=case, **cfg)
    mdt = {"f32": torch.float32, "f16_out": torch.float16, "bf16_out": torch.bfloat16, "f16_pipeline": torch.float16,
           "autocast_f16": torch.float16, "genppo_f16": torch.float16, "genppo_bf16": torch.bfloat16}[flow]
    xdt = mdt if flow in ("f16_pipeline", "genppo_f16", "genppo_bf16") else torch.float32
    ac = {"autocast_f16": torch.float16, "genppo_f16": torch.float16, "genppo_bf16": torch.bfloat16}.get(flow)
    if flow.startswith("genppo"):
        r.factor_net.to("cuda", dtype=mdt), o.factor_net.to("cuda", dtype=mdt)
    r.set_timesteps(n, device="cuda"), o.set_timesteps(n, device="cuda")
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).to(xdt).cuda()
    ctx = (lambda: torch.autocast("cuda", ac)) if ac is not None else __import__("contextlib").nullcontext
    for i in range(n):
        e = torch.randn(B, *shape, generator=g).to(mdt).cuda()
        torch.manual_seed(77 + i)
        with ref_shim.quiet(), ctx(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(e, r.timesteps[i], xr, return_dict=False)
        rng_after_ref = torch.cuda.get_rng_state()
        torch.manual_seed(77 + i)
        with ctx(), torch.no_grad():
            xo, ao, po, co, mo = o.step(e, o.timesteps[i], xo, return_dict=False)
        tag = f"case {case} ({flow}, {cfg}) step {i}"
        assert ao.dtype == ar.dtype and torch.equal(ao, ar), tag + ": actions"
        assert torch.equal(mo, mr) and torch.equal(co["x"], cr["x"]), tag
        torch.testing.assert_close(po, pr, rtol=1.2e-4 if ac is not None else 0, atol=6e-6 if ac is not None else 1e-6)
        assert xo.dtype == xr.dtype, tag + f": latent dtype {xo.dtype} vs {xr.dtype}"
        assert torch.equal(xo, xr), tag + ": latent"
        assert torch.equal(torch.cuda.get_rng_state(), rng_after_ref), tag + ": default generator consumed differently"


@pytest.mark.parametrize("case", range(12))
def test_random_fm_configurations_step_for_step(case):
    rng = random.Random(2000 + case)
    od = rng.choice([2, 2, 3, 4])
    cfg = dict(shift=3.0, use_dynamic_shifting=True, order_dim=od, scaler_dim=rng.choice([0, 0, 1, 2]),
               mu_dim=rng.choice([0, 0, 1]))
    n = rng.choice([3, 5, 8])
    B = rng.choice([1, 2, 4])
    shape = rng.choice([(16, 8), (5, 7), (64, 16)])
    dt = rng.choice([torch.bfloat16, torch.bfloat16, torch.float32, torch.float16])
    r, o = _pair("fm", seed=50 + case, last_std=0.02, **cfg)
    for s in (r, o):
        s.set_timesteps(n, device="cuda", sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
        s.set_begin_index(0)
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).to(dt).cuda()
    for i in range(n):
        v = torch.randn(B, *shape, generator=g).to(dt).cuda()
        torch.manual_seed(5 + i)
        with ref_shim.quiet(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(v, r.timesteps[i], xr, return_dict=False)
        torch.manual_seed(5 + i)
        with torch.no_grad():
            xo, ao, po, co, mo = o.step(v, o.timesteps[i], xo, return_dict=False)
        tag = f"case {case} ({cfg}, {dt}) step {i}"
        assert torch.equal(ao, ar), tag + ": actions"
        assert torch.equal(mo, mr) and torch.equal(co["x"], cr["x"]), tag
        torch.testing.assert_close(po, pr, rtol=8e-6, atol=1e-6)
        assert xo.dtype == xr.dtype and torch.equal(xo, xr), tag + ": latent"


# ---- the baseline solvers, live on the GPU -------------------------------------------------------------
@pytest.mark.parametrize("kind", ["euler", "heun", "dpm-solver", "dpm-solver-multistep"])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_fm_baseline_solvers_next_to_the_reference(kind, dt):
    """FlowMatchGeneralDiscreteScheduler (edit_ppo/scheduler_fm.py:384-488): its sigmas live on the device, so no host
    scalar is involved and CPU- and CUDA-made results coincide; checked live anyway"""
    import consolver_b200 as cb

    ref = ref_shim.load_reference()
    n = 8
    kw = dict(shift=3.0, use_dynamic_shifting=True, type=kind)
    r, o = ref.FlowMatchGeneralDiscreteScheduler(**kw), cb.FlowMatchGeneralDiscreteScheduler(**kw)
    for s in (r, o):
        s.set_timesteps(n, device="cuda", sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
        s.set_begin_index(0)
    g = torch.Generator().manual_seed(3)
    xr = xo = torch.randn(2, 64, 16, generator=g).to(dt).cuda()
    for i in range(n):
        v = torch.randn(2, 64, 16, generator=g).to(dt).cuda()
        xr = r.step(v, r.timesteps[i], xr, return_dict=False)[0]
        xo = o.step(v, o.timesteps[i], xo, return_dict=False)[0]
        assert xo.dtype == xr.dtype and torch.equal(xo, xr), f"{kind} {dt} step {i}"


AMED8 = ([999, 831, 749, 623, 500, 394, 250, 88, 0], [1.0, 0.9976, 1.0, 0.991, 1.0, 0.9907, 1.0, 0.9905, 1.0],
         [1.0, 1.0257, 1.0, 0.9989, 1.0, 1.0022, 1.0, 0.9747, 1.0])


@pytest.mark.parametrize("over", [dict(), dict(algorithm_type="dpmsolver"), dict(prediction_type="v_prediction"),
                                  dict(solver_order=3), dict(solver_type="heun")])
def test_amed_plugin_next_to_the_reference(over):
    """AMED plugin (diffusers_amed_plugin_dpmpp.py, unmodified) over the restated stand-in of its absent diffusers base
    class, run on the GPU next to the drop-in: every step bit-identical (the division by the host-resident alpha_t in
    convert_model_output is a reciprocal multiply on CUDA tensors — CONSOLVER_DPM_CONVERT_DIV_RECIP).  The inherited half
    of this baseline stays pinned only to the restatement of the published algorithm."""
    import consolver_b200 as cb

    ref = ref_shim.load_reference()
    cfg = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1, **over)
    ts, dirs, times = AMED8
    r, o = ref.AMEDDPMSolverMultistepScheduler(**cfg), cb.DPMSolverMultistepScheduler(**cfg)
    assert o.reference_device == "cuda"
    for s in (r, o):
        s.scale_dirs, s.scale_times = dirs, times
        s.set_timesteps(8, device="cuda", timesteps=ts)
    g = torch.Generator().manual_seed(8)
    xr = xo = torch.randn(2, 4, 16, 16, generator=g).cuda()
    for i in range(8):
        e = torch.randn(2, 4, 16, 16, generator=g).cuda()
        xr = r.step(e, r.timesteps[i], xr, return_dict=False)[0]
        xo = o.step(e, o.timesteps[i], xo, return_dict=False)[0]
        assert torch.equal(xo, xr), f"{over} step {i}: {float((xo - xr).abs().max() / xr.abs().max()):.2e}"


@pytest.mark.parametrize("case", range(6))
def test_use_conv_policies_step_for_step(case):
    """use_conv=True (factor_net_ppo.py:108-130,:146-149): per-sample cosine features of the history feed the MLP.  fp32
    model outputs, SD and FM: own sampling, latents bit-identical; tables within the reduction-order spread of the
    feature kernel vs torch's cosine_similarity."""
    rng = random.Random(3000 + case)
    kind = "sd" if case % 2 == 0 else "fm"
    od = rng.choice([3, 4])
    n, B = rng.choice([5, 8]), rng.choice([2, 3])
    if kind == "sd":
        cfg = dict(order_dim=od, scaler_dim=rng.choice([0, 2]), prediction_type=rng.choice(["epsilon", "v_prediction"]),
                   **dict(SD_PROD, use_conv=True))
        shape = (4, 16, 16)
    else:
        cfg = dict(shift=3.0, use_dynamic_shifting=True, order_dim=od, scaler_dim=0, mu_dim=0, use_conv=True)
        shape = (64, 16)
    r, o = _pair(kind, seed=300 + case, last_std=0.5 if kind == "sd" else 0.02, **cfg)
    for s in (r, o):
        if kind == "sd":
            s.set_timesteps(n, device="cuda")
        else:
            s.set_timesteps(n, device="cuda", sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
            s.set_begin_index(0)
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).cuda()
    for i in range(n):
        e = torch.randn(B, *shape, generator=g).cuda()
        torch.manual_seed(9 + i)
        with ref_shim.quiet(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(e, r.timesteps[i], xr, return_dict=False)
        torch.manual_seed(9 + i)
        with torch.no_grad():
            xo, ao, po, co, mo = o.step(e, o.timesteps[i], xo, return_dict=False)
        tag = f"case {case} ({kind}, {cfg}) step {i}"
        assert torch.equal(ao, ar), tag + ": actions"
        torch.testing.assert_close(po, pr, rtol=2e-4, atol=2e-6)
        assert torch.equal(xo, xr), tag + ": latent"


# ---- wider random space (order_dim up to 8, K = 3..161, batch up to 64, full-size latents, every way a caller hands over the
# timestep, step() and step_cfg()).  CONSOLVER_FUZZ_CASES scales the number of drawn cases for one-off deep runs
# (profiles/live_fuzz_r02.md records one with several hundred); the default keeps the GPU suite short. ----------------------
WIDE_CASES = int(os.environ.get("CONSOLVER_FUZZ_CASES", "24"))


def _wide_pair(kind, seed, hidden, K, last_std, **cfg):
    import consolver_b200 as cb

    ref = ref_shim.load_reference()
    fkw = dict(hidden_dim=hidden, num_actions=K)
    with ref_shim.quiet():
        if kind == "sd":
            r = ref.PPOScheduler(factor_net_kwargs=dict(embedding_dim=64, **fkw), **cfg)
            o = cb.PPOScheduler(factor_net_kwargs=dict(embedding_dim=64, **fkw), **cfg)
        else:
            r = ref.FMPPOScheduler(factor_net_kwargs=dict(fkw), **cfg)
            o = cb.FMPPOScheduler(factor_net_kwargs=dict(fkw), **cfg)
    _seed_policy(r.factor_net, seed, last_std)
    o.factor_net.load_state_dict(r.factor_net.state_dict())
    r.factor_net.cuda(), o.factor_net.cuda()
    return r, o


def _hand_over(style, sched, i):
    """the ways callers pass the timestep: a 0-d view of scheduler.timesteps (`for t in scheduler.timesteps`), a python
    int, a clone on the device, a CPU tensor"""
    t = sched.timesteps[i]
    return {"view": lambda: t, "int": lambda: int(t), "clone": lambda: t.clone(), "cpu": lambda: t.cpu()}[style]()


@pytest.mark.parametrize("case", range(WIDE_CASES))
def test_wide_random_sd_configurations_step_for_step(case):
    rng = random.Random(7000 + case)
    od = rng.choice([2, 3, 4, 4, 5, 6, 8])
    cfg = dict(order_dim=od, scaler_dim=rng.choice([0, 0, 1, 2]),
               prediction_type=rng.choice(["epsilon", "epsilon", "v_prediction"]),
               timestep_spacing=rng.choice(["trailing", "leading", "linspace"]),
               beta_schedule=rng.choice(["scaled_linear", "linear", "squaredcos_cap_v2"]),
               beta_start=0.00085, beta_end=0.012, steps_offset=rng.choice([0, 1]), use_conv=False)
    K = rng.choice([3, 11, 11, 161])
    hidden = rng.choice([16, 64, 256])
    n = rng.choice([1, 2, 4, 8, 15, 21])
    B = rng.choice([1, 2, 5, 33, 64])
    shape = rng.choice([(4, 8, 8), (3, 5, 7), (1, 1, 33), (4, 32, 32)] + ([(4, 64, 64)] if B <= 5 else []))
    flow = rng.choice(["f32", "f32", "f32", "f16_out", "bf16_out", "f16_pipeline", "bf16_pipeline", "autocast_f16",
                       "autocast_bf16", "genppo_f16", "genppo_bf16"])
    style = rng.choice(["view", "view", "int", "clone", "cpu"])
    fused_cfg = rng.choice([False, True])
    guidance = rng.choice([3.0, 7.5, 1.0])
    last_std = rng.choice([0.5, 0.05, 2.0])
    # which timesteps the caller walks: the grid; the grid from a later point (img2img-style `timesteps[t_start:]`); or
    # integers of its own that are not grid points (the reference honours whatever it is handed, scheduler_ppo.py:203-207)
    walk = rng.choice(["grid", "grid", "grid", "from_k", "off_grid"])
    r, o = _wide_pair("sd", case, hidden, K, last_std, **cfg)
    mdt = torch.float32 if flow == "f32" else (torch.float16 if "f16" in flow and "bf16" not in flow else torch.bfloat16)
    xdt = mdt if flow.endswith("_pipeline") or flow.startswith("genppo") else torch.float32
    ac = mdt if flow.startswith(("autocast", "genppo")) else None
    if flow.startswith("genppo"):
        r.factor_net.to("cuda", dtype=mdt), o.factor_net.to("cuda", dtype=mdt)
    r.set_timesteps(n, device="cuda"), o.set_timesteps(n, device="cuda")
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).to(xdt).cuda()
    ctx = (lambda: torch.autocast("cuda", ac)) if ac is not None else __import__("contextlib").nullcontext
    tag0 = (f"wide case {case} ({flow}, K={K}, H={hidden}, std={last_std}, n={n}, B={B}, {shape}, t as {style}, "
            f"walk={walk}, cfg={fused_cfg}, {cfg})")
    first = rng.randrange(n) if walk == "from_k" else 0
    own = sorted(rng.sample(range(1000), n), reverse=True) if walk == "off_grid" else None
    for i in range(first, n):
        pair = torch.randn(2 * B, *shape, generator=g).to(mdt).cuda()
        u, c = pair.chunk(2)
        e = u + guidance * (c - u)                       # the caller's combine, denoise_ppo.py:96-100
        torch.manual_seed(77 + i)
        t_r = own[i] if own else _hand_over(style, r, i)
        t_o = own[i] if own else _hand_over(style, o, i)
        with ref_shim.quiet(), ctx(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(e, t_r, xr, return_dict=False)
        rng_after_ref = torch.cuda.get_rng_state()
        torch.manual_seed(77 + i)
        with ctx(), torch.no_grad():
            if fused_cfg:
                xo, ao, po, co, mo = o.step_cfg(pair, t_o, xo, guidance)
                assert torch.equal(o.ets[-1], e), tag0 + f" step {i}: ring slot != caller-side combine"
            else:
                xo, ao, po, co, mo = o.step(e, t_o, xo, return_dict=False)
        tag = tag0 + f" step {i}"
        assert ao.dtype == ar.dtype and torch.equal(ao, ar), tag + ": actions"
        assert torch.equal(mo, mr) and torch.equal(co["x"], cr["x"]), tag
        # probabilities: the measured bars of DESIGN §4 belong to policies of the production scale (last layer std <= 0.5).
        # At std 2.0 the logits are ~4x larger: an fp32 rounding difference of a logit moves p by up to 4x more, and under
        # autocast a logit that lands on the other side of a 16-bit rounding boundary moves by a whole 16-bit ulp of a
        # larger number (observed up to 4.6e-4).  The hard checks — actions, latents, generator state — do not loosen.
        # Hidden layers here are U(-0.3, 0.3) at widths up to 256 (pre-activations several times those of a default-
        # initialised policy) and K goes to 161; over 400 drawn cases the worst fp32 difference was 1.1e-6 at std <= 0.5
        # and 3.2e-6 at std 2.0 (tolerances: 2x those).
        big = last_std > 0.5
        if ac is not None:
            torch.testing.assert_close(po, pr, rtol=2e-3, atol=2e-3 if big else 5e-4)
        else:
            torch.testing.assert_close(po, pr, rtol=0, atol=8e-6 if big else 2.2e-6)
        assert xo.dtype == xr.dtype, tag + f": latent dtype {xo.dtype} vs {xr.dtype}"
        assert torch.equal(xo, xr), tag + ": latent"
        assert torch.equal(torch.cuda.get_rng_state(), rng_after_ref), tag + ": default generator consumed differently"


@pytest.mark.parametrize("case", range(max(WIDE_CASES // 2, 1)))
def test_wide_random_fm_configurations_step_for_step(case):
    rng = random.Random(8000 + case)
    od = rng.choice([2, 2, 3, 4, 6])
    cfg = dict(shift=rng.choice([3.0, 1.0]), use_dynamic_shifting=rng.choice([True, True, False]), order_dim=od,
               scaler_dim=rng.choice([0, 0, 1, 2]), mu_dim=rng.choice([0, 0, 1]))
    K = rng.choice([3, 11, 11, 161])
    hidden = rng.choice([16, 64, 256])
    n = rng.choice([1, 2, 5, 8, 15])
    B = rng.choice([1, 2, 4, 16])
    shape = rng.choice([(16, 8), (5, 7), (64, 16), (256, 64)] + ([(4096, 64)] if B <= 2 else []))
    dt = rng.choice([torch.bfloat16, torch.bfloat16, torch.float32, torch.float16])
    style = rng.choice(["view", "view", "clone", "float"])
    last_std = rng.choice([0.02, 0.002, 0.1])
    # sigma-grid options (edit_ppo/scheduler_fmppo.py:218-236) and how the step index is found (:249-268)
    opt = rng.choice([None, None, None, "use_karras_sigmas", "use_exponential_sigmas", "use_beta_sigmas"])
    if opt:
        cfg[opt] = True
    cfg["shift_terminal"] = rng.choice([None, None, 0.02])
    if n == 1:
        cfg["shift_terminal"] = None      # a one-point grid stretched to a terminal value is 0/0 in the reference (and here)
    cfg["invert_sigmas"] = rng.choice([False, False, False, True])
    cfg["time_shift_type"] = rng.choice(["exponential", "exponential", "linear"])
    begin = rng.choice([0, 0, None, "mid"])
    r, o = _wide_pair("fm", 50 + case, hidden, K, last_std, **cfg)
    # softmax(logits / 0.01): a rounding difference of the logits is multiplied by 100, and the logits grow with the last
    # layer's scale — the measured 8e-6 belongs to last_std 0.02
    p_rtol = 8e-6 * max(1.0, last_std / 0.02)
    for s in (r, o):
        if cfg["use_dynamic_shifting"]:
            s.set_timesteps(n, device="cuda", sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
        else:
            s.set_timesteps(n, device="cuda")
    assert torch.equal(r.timesteps, o.timesteps) and torch.equal(r.sigmas, o.sigmas), f"wide fm case {case}: grid {cfg}"
    # the step index: set by the pipeline (0, or mid-grid for image-to-image), or looked up from the first timestep
    unique = len(set(r.timesteps.tolist())) == n
    first = rng.randrange(n) if begin == "mid" else 0
    for s in (r, o):
        if begin is not None or not unique:
            s.set_begin_index(first)
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).to(dt).cuda()
    hand = lambda s, i: float(s.timesteps[i]) if style == "float" else _hand_over(style, s, i)  # noqa: E731
    for i in range(first, n):
        v = torch.randn(B, *shape, generator=g).to(dt).cuda()
        torch.manual_seed(5 + i)
        with ref_shim.quiet(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(v, hand(r, i), xr, return_dict=False)
        rng_after_ref = torch.cuda.get_rng_state()
        torch.manual_seed(5 + i)
        with torch.no_grad():
            xo, ao, po, co, mo = o.step(v, hand(o, i), xo, return_dict=False)
        tag = (f"wide fm case {case} (K={K}, H={hidden}, n={n}, B={B}, {shape}, {dt}, t as {style}, begin={begin}, {cfg}) "
               f"step {i}")
        assert torch.equal(ao, ar), tag + ": actions"
        assert torch.equal(mo, mr) and torch.equal(co["x"], cr["x"]), tag
        torch.testing.assert_close(po, pr, rtol=p_rtol, atol=1e-6)
        assert xo.dtype == xr.dtype and torch.equal(xo, xr), tag + ": latent"
        assert torch.equal(torch.cuda.get_rng_state(), rng_after_ref), tag + ": default generator consumed differently"


@pytest.mark.parametrize("case", range(max(WIDE_CASES // 2, 1)))
def test_wide_random_use_conv_configurations_step_for_step(case):
    """use_conv=True over the wider space (fp32 model outputs; order 3-6, K to 161, batch to 16, fused CFG): the per-sample
    policy kernel (cosine features -> MLP -> draw) feeds the step kernel through the same programmatic dependent launch."""
    rng = random.Random(9000 + case)
    kind = rng.choice(["sd", "sd", "fm"])
    od = rng.choice([3, 4, 4, 5, 6])
    K = rng.choice([11, 11, 161])
    hidden = rng.choice([16, 64, 256])
    n = rng.choice([2, 5, 8, 12])
    B = rng.choice([1, 2, 5, 16])
    fused_cfg = kind == "sd" and rng.choice([False, True])
    if kind == "sd":
        cfg = dict(order_dim=od, scaler_dim=rng.choice([0, 1, 2]), prediction_type=rng.choice(["epsilon", "v_prediction"]),
                   **dict(SD_PROD, use_conv=True))
        shape = rng.choice([(4, 16, 16), (4, 32, 32), (3, 5, 7)])
    else:
        cfg = dict(shift=3.0, use_dynamic_shifting=True, order_dim=od, scaler_dim=rng.choice([0, 2]), mu_dim=0, use_conv=True)
        shape = rng.choice([(64, 16), (256, 64), (5, 7)])
    last_std = 0.5 if kind == "sd" else 0.02
    r, o = _wide_pair(kind, 300 + case, hidden, K, last_std, **cfg)
    for s in (r, o):
        if kind == "sd":
            s.set_timesteps(n, device="cuda")
        else:
            s.set_timesteps(n, device="cuda", sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
            s.set_begin_index(0)
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).cuda()
    for i in range(n):
        pair = torch.randn(2 * B, *shape, generator=g).cuda()
        u, c = pair.chunk(2)
        e = u + 3.0 * (c - u)
        torch.manual_seed(9 + i)
        with ref_shim.quiet(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(e, r.timesteps[i], xr, return_dict=False)
        torch.manual_seed(9 + i)
        with torch.no_grad():
            if fused_cfg:
                xo, ao, po, co, mo = o.step_cfg(pair, o.timesteps[i], xo, 3.0)
            else:
                xo, ao, po, co, mo = o.step(e, o.timesteps[i], xo, return_dict=False)
        tag = f"wide use_conv case {case} ({kind}, K={K}, H={hidden}, n={n}, B={B}, {shape}, cfg={fused_cfg}, {cfg}) step {i}"
        assert torch.equal(ao, ar), tag + ": actions"
        torch.testing.assert_close(po, pr, rtol=2e-4, atol=4e-6)
        assert torch.equal(xo, xr), tag + ": latent"


@pytest.mark.parametrize("case", range(max(WIDE_CASES // 2, 1)))
def test_our_caller_loop_equals_the_references_caller_loop(case):
    """BOTH sides of the boundary swapped: consolver_b200.denoise.denoise_loop (no torch.cat of the CFG-doubled input, the
    step kernel writes the next latent into both halves; rollout record as views) driving the drop-in scheduler, against
    the reference's denoise_ppo.denoise_diffusion driving the reference scheduler — random configurations, fp32 and 16-bit
    denoiser outputs: final latents and the whole rollout record."""
    from consolver_b200.denoise import denoise_loop

    caller = _load_caller()
    rng = random.Random(10000 + case)
    od = rng.choice([2, 3, 4, 4, 6])
    cfg = dict(order_dim=od, scaler_dim=rng.choice([0, 0, 1, 2]), prediction_type=rng.choice(["epsilon", "v_prediction"]),
               **SD_PROD)
    n, B = rng.choice([2, 3, 8, 15]), rng.choice([1, 2, 5, 16])
    shape = rng.choice([(4, 16, 16), (4, 8, 8), (4, 32, 32)])
    odt = rng.choice([torch.float32, torch.float32, torch.bfloat16, torch.float16])   # the denoiser's output dtype
    guidance = rng.choice([3.0, 7.5])
    r, o = _wide_pair("sd", 700 + case, rng.choice([64, 256]), rng.choice([11, 161]), 0.5, **cfg)
    noise = torch.randn(B, *shape, generator=torch.Generator().manual_seed(case)).cuda()
    text = [f"prompt {i}" * (i + 1) for i in range(B)]
    unet = lambda x, t, encoder_hidden_states=None, return_dict=False: (  # noqa: E731
        _unet(x, t, encoder_hidden_states)[0].to(odt),)
    tok = _Tok()
    ids = lambda t_: tok(t_).input_ids.cuda()  # noqa: E731
    embeds = torch.cat([_text_encoder(ids([""] * B))[0], _text_encoder(ids(text))[0]])
    torch.manual_seed(4321)
    with ref_shim.quiet(), torch.no_grad():
        lat_r, conds_r, probs_r, act_r, masks_r, _ = caller.denoise_diffusion(
            _text_encoder, r, unet, noise, text, tok, cfg=guidance, num_inference_steps=n)
    torch.manual_seed(4321)
    with torch.no_grad():
        lat_o, rec = denoise_loop(o, lambda x, t, i: unet(x, t, embeds)[0], noise, cfg=guidance, num_inference_steps=n)
    tag = f"caller case {case} (n={n}, B={B}, {shape}, out {odt}, g={guidance}, {cfg})"
    assert torch.equal(rec["actions"], act_r), tag + ": actions"
    assert torch.equal(rec["masks"], masks_r) and torch.equal(rec["x"].to(conds_r["x"].dtype), conds_r["x"]), tag
    torch.testing.assert_close(rec["probs"], probs_r, rtol=0, atol=2.2e-6)
    assert lat_o.dtype == lat_r.dtype and torch.equal(lat_o, lat_r), tag + ": final latents"


USE_CONV_16BIT_STATS = {"draws": 0, "mismatches": 0}


@pytest.mark.parametrize("case", range(max(WIDE_CASES // 2, 1)))
def test_use_conv_with_16bit_outputs_and_autocast_next_to_the_reference(case):
    """use_conv=True on bf16 model outputs / under bf16 autocast (formerly "no fixture").  The reference evaluates
    F.cosine_similarity in bf16 (every op rounded to it), the feature kernel reduces in fp32 / fp64, so the MLP inputs agree
    to bf16 precision only and a draw that sits on a near-tie of p/q can legitimately pick another bin.
    (fp16 outputs are not drawn: there the REFERENCE itself dies — for the zero-padded history slots of the first steps
    cosine_similarity's eps^2 = 1e-16 underflows to 0 in fp16, the features are 0 * inf = NaN, and torch.multinomial's
    device-side assert "probability tensor contains inf, nan" takes the CUDA context down; nothing to reproduce.)
    Every step starts from the REFERENCE's latent (one differing draw must not contaminate the later steps); checked: the
    draws coincide but for a small fraction (< 3 % over the run), and wherever a sample's draws coincide its latent is
    bit-identical, dtype included."""
    rng = random.Random(14000 + case)
    od = rng.choice([3, 4])
    cfg = dict(order_dim=od, scaler_dim=rng.choice([0, 0, 2]), prediction_type=rng.choice(["epsilon", "v_prediction"]),
               **dict(SD_PROD, use_conv=True))
    n, B = rng.choice([5, 8]), rng.choice([2, 5, 16])
    shape = rng.choice([(4, 16, 16), (4, 32, 32)])
    flow = rng.choice(["bf16_out", "autocast_bf16"])
    mdt = torch.bfloat16
    ac = mdt if flow.startswith("autocast") else None
    r, o = _wide_pair("sd", 900 + case, 64, 11, 0.5, **cfg)
    r.set_timesteps(n, device="cuda"), o.set_timesteps(n, device="cuda")
    g = torch.Generator().manual_seed(case)
    x = torch.randn(B, *shape, generator=g).cuda()
    ctx = (lambda: torch.autocast("cuda", ac)) if ac is not None else __import__("contextlib").nullcontext
    for i in range(n):
        e = torch.randn(B, *shape, generator=g).to(mdt).cuda()
        torch.manual_seed(21 + i)
        with ref_shim.quiet(), ctx(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(e, r.timesteps[i], x, return_dict=False)
        state = torch.cuda.get_rng_state()
        torch.manual_seed(21 + i)
        with ctx(), torch.no_grad():
            xo, ao, po, co, mo = o.step(e, o.timesteps[i], x, return_dict=False)
        tag = f"use_conv 16-bit case {case} ({flow}, n={n}, B={B}, {shape}, {cfg}) step {i}"
        assert torch.equal(torch.cuda.get_rng_state(), state), tag + ": default generator consumed differently"
        assert xo.dtype == xr.dtype and torch.equal(mo, mr), tag
        same = (ao == ar).all(dim=1)
        USE_CONV_16BIT_STATS["draws"] += B
        USE_CONV_16BIT_STATS["mismatches"] += int((~same).sum())
        assert torch.equal(xo[same], xr[same]), tag + ": latent of samples whose draws coincide"
        torch.testing.assert_close(po[same], pr[same], rtol=5e-2, atol=5e-3)
        x = xr
    st = USE_CONV_16BIT_STATS
    print(f"use_conv 16-bit draws so far: {st['mismatches']} / {st['draws']} samples with a differing bin")
    assert st["mismatches"] <= max(3, 0.03 * st["draws"]), st
