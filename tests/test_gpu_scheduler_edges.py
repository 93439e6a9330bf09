"""GPU: scheduler-level edge cases of the drop-in API (dtypes, layouts, off-grid timesteps, timestep read-back,
batch changes, empty batches)."""
import pytest
import torch

import consolver_oracle as orc
from golden_io import Golden

# cpu_reference: these tests check against CPU-made fixtures / the oracle's default (CPU-torch) rules; the product
# default — the reference as executed on CUDA tensors — is covered by tests/test_gpu_cuda_reference.py
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("cpu_reference")]


def _pair(name="sd_eps_s0_n8_B3"):
    import consolver_b200 as cb
    g = Golden(name)
    m = g.meta
    s = cb.PPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
    s.factor_net.load_state_dict(g.state_dict)
    s.factor_net.cuda()
    o = orc.OracleSDScheduler(g.state_dict, **m["config"])
    return g, m, s, o


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_16bit_pipeline_dtype(dtype):
    """A 16-bit pipeline (gen_ppo.py: fp16 latents and U-Net output) through the fused-CFG step: the reference's torch
    ops, run on the same dtypes by the oracle, decide everything — the first latent comes back 16-bit, torch
    promotion makes every later one fp32 — and the kernels reproduce it bit for bit, dtype trajectory included."""
    g, m, s, o = _pair()
    s.set_timesteps(m["n"], device="cuda")
    o.set_timesteps(m["n"])
    s.replay = {"idx": [g[f"idx_{i}"] for i in range(m["n"])]}
    x = g["x_T"].to(dtype)
    xg = x.cuda()
    for i, t in enumerate(o.timesteps):
        pair = g[f"pair_{i}"].to(dtype)
        xg = s.step_cfg(pair.cuda(), s.timesteps[i], xg, m["guidance"])[0]
        u, c = pair.chunk(2)
        x = o.step(orc.cfg_combine(u, c, m["guidance"]), t, x, forced_idx=g[f"idx_{i}"])[0]
        assert xg.dtype == x.dtype == (dtype if i == 0 else torch.float32)
        assert torch.equal(xg.cpu(), x), f"step {i}"


def test_non_contiguous_inputs_and_off_grid_timesteps():
    g, m, s, o = _pair()
    s.set_timesteps(m["n"], device="cuda")
    o.set_timesteps(m["n"])
    x = g["x_T"]
    xg = x.cuda()
    for i in range(4):
        t = int(o.timesteps[i]) - 7                    # not a grid point: the fused MLP+sample kernel is used
        eps = g[f"eps_{i}"]
        q = g[f"q_{i}"]
        s.replay = {"q": {s._traj.count if s._traj else 0: q.cuda()}}
        eps_nc = eps.cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)      # channels_last strides
        assert not eps_nc.is_contiguous()
        xg, actions, probs, conds, masks = s.step(eps_nc, t, xg, return_dict=False)
        x, a_o, p_o, _, m_o = o.step(eps, t, x, q=q)
        assert torch.equal(actions.cpu(), a_o) and torch.equal(masks.cpu(), m_o)
        torch.testing.assert_close(probs.cpu(), p_o, rtol=0, atol=1e-6)
        assert torch.equal(conds["x"].cpu(), torch.tensor([[t, t - 125]], dtype=torch.float32).repeat(3, 1))
        assert torch.equal(xg.cpu(), x)


def test_timestep_read_back_and_batch_change():
    g, m, s, o = _pair()
    s.set_timesteps(m["n"], device="cuda")
    s.sync_free = False                               # .item() the CUDA timestep like the reference does
    x = g["x_T"].cuda()
    out = s.step(g["eps_0"].cuda(), s.timesteps[0], x, return_dict=True)
    assert out.prev_sample.shape == x.shape and out.actions.shape == (3, 3)
    # a different batch size mid-trajectory re-allocates the per-trajectory buffers and keeps working
    x5 = torch.randn(5, *m["shape"], device="cuda")
    s.set_timesteps(m["n"], device="cuda")
    out5 = s.step(torch.randn_like(x5), s.timesteps[0], x5, return_dict=True)
    assert out5.actions.shape == (5, 3)
    with pytest.raises(TypeError):
        s.step(torch.randn_like(x5).half(), s.timesteps[1], x5.bfloat16())     # two different 16-bit types
    with pytest.raises(ValueError):
        s.step_cfg(torch.randn(7, *m["shape"], device="cuda"), s.timesteps[1], x5, 3.0)


def test_empty_batch_is_rejected_loudly():
    from consolver_b200._lib import ConsolverError
    g, m, s, o = _pair()
    s.set_timesteps(m["n"], device="cuda")
    x0 = torch.zeros(0, *m["shape"], device="cuda")
    with pytest.raises((ConsolverError, ZeroDivisionError, RuntimeError, ValueError)):
        s.step(x0.clone(), s.timesteps[0], x0)


def test_fp16_cast_policy_and_sample_action_api():
    """gen_ppo.py:193-195 casts the loaded policy (and its bin buffer) to fp16 — which the reference can only run under
    autocast (gen_ppo.py:309): fp16 Linear layers, fp32 softmax, fp16 bin values.  Exercises the public
    FactorNetPPO.sample_action(x_dict) (factor_net_ppo.py:159-168), which no longer reads the row back to the host."""
    import consolver_b200 as cb
    g = Golden("sd_eps_s0_n8_B3")
    fn = cb.FactorNetPPO(**{**g.meta["factor_net_kwargs"], "order_dim": 4, "scaler_dim": 0})
    fn.load_state_dict(g.state_dict)
    fn.to("cuda", dtype=torch.float16)
    sd16 = {k: v.to(torch.float16).float() for k, v in g.state_dict.items()}     # what the kernel must see
    x = torch.tensor([[874.0, 749.0]]).repeat(5, 1).cuda()
    torch.manual_seed(3)
    with torch.cuda.device(0):
        torch.cuda.set_sync_debug_mode("error")          # the call must not synchronise with the host
        try:
            actions, probs = fn.sample_action({"x": x})
        finally:
            torch.cuda.set_sync_debug_mode("default")
    assert actions.shape == (5, 3) and probs.shape == (5, 3) and actions.dtype == torch.float16
    ref = orc.policy_probs(g.state_dict, x[:1].cpu(), "sd", sem=orc.TorchSemantics("cuda", torch.float16))[0]
    # every sampled probability is an entry of the fp16-autocast table, every action a (fp16-rounded) bin value
    for b in range(5):
        for a in range(3):
            k = (sd16["action_values"][a] - actions[b, a].float().cpu()).abs().argmin()
            assert sd16["action_values"][a, k] == actions[b, a].float().cpu()
            torch.testing.assert_close(probs[b, a].cpu(), ref[a, k], rtol=1.2e-4, atol=6e-6)
    # after an in-place weight update the cached fp32 copies are refreshed
    with torch.no_grad():
        fn.mlp[4].bias.add_(1.0)
    p2 = fn.sample_action({"x": x})[1]
    assert p2.shape == (5, 3)


def test_cuda_timesteps_starting_mid_grid_like_img2img():
    """A pipeline that starts at timesteps[t_start:] (img2img strength < 1) hands CUDA timestep tensors from the middle
    of the grid: the scheduler locates the start with one read-back and follows the grid from there."""
    g, m, s, o = _pair()
    s.set_timesteps(m["n"], device="cuda")
    o.set_timesteps(m["n"])
    start = 3
    x = g["x_T"]
    xg = x.cuda()
    s.replay = {"q": {k: g[f"q_{start + k}"].cuda() for k in range(m["n"] - start)}}
    for k, i in enumerate(range(start, m["n"])):
        eps = g[f"eps_{i}"]
        xg, actions, _, conds, _ = s.step(eps.cuda(), s.timesteps[i], xg, return_dict=False)     # CUDA 0-d tensor
        x, a_o, _, c_o, _ = o.step(eps, o.timesteps[i], x, q=g[f"q_{i}"])
        assert torch.equal(conds["x"].cpu(), c_o["x"]), f"step {i}: wrong timestep row"
        assert torch.equal(actions.cpu(), a_o)
        assert torch.equal(xg.cpu(), x)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_autocast_style_mixed_precision_matches_reference_promotion(dtype):
    """train_ppo.py:353 runs the rollout under accelerate autocast: the U-Net output is 16-bit, the latents fp32.
    The reference's torch ops then promote: the returned latent is fp32, and from the second step on every product
    is fp32 arithmetic on the upcast model outputs; at the first step (eff is the 16-bit output itself) two products
    are 16-bit products.  The oracle (fed the same mixed dtypes) and the kernel agree bit for bit at every step."""
    g, m, s, o = _pair()
    s.set_timesteps(m["n"], device="cuda")
    o.set_timesteps(m["n"])
    s.replay = {"idx": [g[f"idx_{i}"] for i in range(m["n"])]}
    x = g["x_T"]
    for i, t in enumerate(o.timesteps):
        eps = g[f"eps_{i}"].to(dtype)
        out = s.step(eps.cuda(), s.timesteps[i], x.cuda(), return_dict=False)
        ref = o.step(eps, t, x, forced_idx=g[f"idx_{i}"])
        assert out[0].dtype == torch.float32 and ref[0].dtype == torch.float32
        assert torch.equal(out[0].cpu(), ref[0]), f"step {i}"
        assert torch.equal(out[1].cpu(), ref[1]) and torch.equal(out[4].cpu(), ref[4])
        x = out[0].cpu()
    # the other direction: 16-bit latent with an fp32 model output is promoted to fp32
    s.set_timesteps(m["n"], device="cuda")
    s.replay = {"idx": [g[f"idx_{i}"] for i in range(m["n"])]}
    y = s.step(g["eps_0"].cuda(), s.timesteps[0], g["x_T"].to(dtype).cuda(), return_dict=False)[0]
    assert y.dtype == torch.float32
    with pytest.raises(TypeError):
        s.step(g["eps_1"].half().cuda(), s.timesteps[1], y.bfloat16(), return_dict=False)


@pytest.mark.parametrize("noise_dtype", [torch.float32, torch.float16])
def test_denoise_loop_with_a_16bit_denoiser_output_ends_with_fp32_latents(noise_dtype):
    """The rollout loop (denoise_ppo.py:52-120) with an fp16 denoiser output: under autocast the ping-pong latents are
    fp32 throughout; starting from fp16 noise (gen_ppo.py) they turn fp32 at the second step and the loop re-types
    its buffers."""
    from consolver_b200.denoise import denoise_loop
    g, m, s, _ = _pair()
    w = torch.randn(4, 4, device="cuda") * 0.3
    den = lambda x, t, i: torch.einsum("oc,bchw->bohw", w.to(x.dtype), x).half()  # noqa: E731
    noise = g["x_T"].to(noise_dtype).cuda()
    torch.manual_seed(3)
    lat, rec = denoise_loop(s, den, noise, cfg=3.0, num_inference_steps=6)
    assert lat.dtype == torch.float32 and rec["actions"].shape[:2] == (3, 5)
    # the same loop spelled out, replaying the sampled actions
    idx = s._traj.out["idx"][:6].clone()
    s.set_timesteps(6, device="cuda")
    s.replay = {"idx": [idx[i] for i in range(6)]}
    x = noise
    for i, t in enumerate(s.timesteps):
        x = s.step_cfg(den(torch.cat([x] * 2), t, i), t, x, 3.0)[0]
    assert torch.equal(x, lat)


def test_policy_cast_to_the_pipeline_dtype_promotes_after_the_first_step():
    """gen_ppo.py:193-195 casts the policy, bin buffer included, to the pipeline's fp16 and runs under autocast: the
    reference's coefficients are fp16 tensors EXCEPT the closing one (torch.sum returns fp32 under autocast), so the
    latent is fp16 after the first step only and fp32 from the second on.  The dtype trajectory here; the bit-for-bit
    check of this flow is tests/test_gpu_cuda_reference.py on the cuda_genppo_* fixtures (made by the reference on a
    B200)."""
    g, m, s, _ = _pair()
    s.factor_net.to("cuda", dtype=torch.float16)
    s.set_timesteps(m["n"], device="cuda")
    s.replay = {"idx": [g[f"idx_{i}"] for i in range(m["n"])]}
    x = g["x_T"].half().cuda()
    with torch.autocast("cuda", torch.float16):
        for i in range(m["n"]):
            eps = g[f"eps_{i}"].half().cuda()
            want = torch.float16 if i == 0 else torch.float32
            assert s.next_latent_dtype(eps.dtype, x.dtype) == want
            x, actions = s.step(eps, s.timesteps[i], x, return_dict=False)[:2]
            assert x.dtype == want, f"step {i}"
            assert actions.dtype == torch.float16          # bin values come back in the policy's dtype, as the reference's do


@pytest.mark.parametrize("kind", ["learned", "heun"])
def test_flow_denoise_loop_writes_into_the_packed_transformer_input(kind):
    """edit_ppo/denoise_diffusion.py:84-160 with reference-image latents appended along the sequence axis: the ready-made
    loop (no per-step torch.cat) against the reference's loop spelled out with torch.cat and noise_pred[:, :L]."""
    import numpy as np
    import consolver_b200 as cb
    from consolver_b200.denoise import denoise_loop_flow
    B, L, Li, D, n = 2, 64, 32, 16, 6
    kw = dict(shift=3.0, use_dynamic_shifting=True)
    if kind == "learned":
        s = cb.FMPPOScheduler(order_dim=2, scaler_dim=0, mu_dim=0, factor_net_kwargs=dict(hidden_dim=32, num_actions=11),
                              **kw)
        with torch.no_grad():
            s.factor_net.mlp[4].weight.normal_(0, 0.02)
        s.factor_net.cuda()
    else:
        s = cb.FlowMatchGeneralDiscreteScheduler(type="heun", **kw)
    gen = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.randn(B, L, D, device="cuda", generator=gen).bfloat16()
    img = torch.randn(B, Li, D, device="cuda", generator=gen).bfloat16()
    w = (torch.randn(D, D, device="cuda", generator=gen) * 0.2).bfloat16()
    seen = []

    def den(inp, t, i):
        assert inp.shape == (B, L + Li, D)
        seen.append(inp.clone())
        return (inp @ w) * (1.0 + 0.01 * i)

    sig = np.linspace(1.0, 1 / n, n)
    torch.manual_seed(5)
    lat, rec = denoise_loop_flow(s, den, x0, img, num_inference_steps=n, sigmas=sig, mu=1.15)
    assert lat.dtype == torch.bfloat16 and all(torch.equal(v[:, L:], img) for v in seen)
    if kind == "learned":
        assert rec["actions"].shape == (B, n - 1, 1) and rec["x"].shape == (B, n - 1, 2)
        idx = s._traj.out["idx"][:n].clone()
        s.replay = {"idx": [idx[i] for i in range(n)]}
    else:
        assert rec is None
    # the reference's loop, spelled out
    s.set_timesteps(n, device="cuda", sigmas=sig, mu=1.15)
    s.set_begin_index(0)
    x = x0
    for i, t in enumerate(s.timesteps):
        inp = torch.cat([x, img], dim=1)
        assert torch.equal(inp, seen[i]), f"step {i}: transformer input differs"
        x = s.step(den(inp, t, i)[:, :L], t, x, return_dict=False)[0]
    assert torch.equal(x, lat)


def _no_sync():
    """context: any synchronising torch call (.item(), nonzero, blocking copies ...) raises"""
    import contextlib

    @contextlib.contextmanager
    def ctx():
        prev = torch.cuda.get_sync_debug_mode()
        torch.cuda.set_sync_debug_mode("error")
        try:
            yield
        finally:
            torch.cuda.set_sync_debug_mode(prev)
    return ctx()


def test_pipeline_style_loops_never_synchronise_with_the_host():
    """`for t in scheduler.timesteps: scheduler.step(model_output, t, latents)` — the loop every diffusers pipeline
    runs, with `t` a CUDA 0-d tensor — completes without a single host synchronisation for all four schedulers,
    first step of the trajectory included (the reference reads the timestep back several times per step).
    `set_timesteps` (one upload of the grid, as in the reference) stays outside the checked region."""
    import numpy as np
    import consolver_b200 as cb
    g, m, s, _ = _pair()
    x0 = g["x_T"].cuda()
    eps = [g[f"eps_{i}"].cuda() for i in range(m["n"])]
    fm = cb.FMPPOScheduler(shift=3.0, use_dynamic_shifting=True, order_dim=2, scaler_dim=0, mu_dim=0,
                           factor_net_kwargs=dict(hidden_dim=32, num_actions=11))
    fm.factor_net.cuda()
    base = cb.FlowMatchGeneralDiscreteScheduler(shift=3.0, use_dynamic_shifting=True, type="heun")
    dpm = cb.DPMSolverMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    v = torch.randn(2, 64, 16, device="cuda").bfloat16()
    sig = np.linspace(1.0, 1 / 6, 6)

    def loop(sched, outputs, x, start=0):
        for i, t in enumerate(sched.timesteps[start:]):          # CUDA 0-d views of the grid tensor
            x = sched.step(outputs[i], t, x, return_dict=False)[0]
        return x

    cases = [
        (s, lambda: s.set_timesteps(m["n"], device="cuda"), eps, x0, 0),
        (s, lambda: s.set_timesteps(m["n"], device="cuda"), eps, x0, 3),       # a pipeline that starts mid-grid
        (fm, lambda: fm.set_timesteps(6, device="cuda", sigmas=sig, mu=1.15), [v] * 6, v, 0),   # no set_begin_index:
        (base, lambda: base.set_timesteps(6, device="cuda", sigmas=sig, mu=1.15), [v] * 6, v, 0),  # step() finds t
        (dpm, lambda: dpm.set_timesteps(6, device="cuda"), eps, x0, 0),
    ]
    for sched, prepare, outputs, x, start in cases:
        prepare()
        loop(sched, outputs, x, start)         # warm-up pass: library load, lazy one-off state
        prepare()
        torch.cuda.synchronize()
        with _no_sync():
            loop(sched, outputs, x, start)
        torch.cuda.synchronize()


# ---- behaviours fixed after the round-1 review ------------------------------------------------------------------------
def test_a_cuda_timestep_that_is_not_a_view_of_the_grid_is_honoured():
    """A caller that clones, reorders or repeats timesteps gets steps at the values it PASSED (the reference reads every
    timestep back, scheduler_ppo.py:205) — not at grid[step_count].  Two schedulers: one fed views in grid order, one fed
    clones in a permuted order with a repeat; each step must equal the same (t, history) evaluated through ints."""
    g, m, s, _ = _pair()
    n = m["n"]
    order = [0, 3, 3, 1, 5, 2]                                   # a Heun-style repeat and a custom subset
    idx = [g[f"idx_{i}"] for i in range(n)]
    outs = {}
    for how in ("clone", "int"):
        s.set_timesteps(n, device="cuda")
        s.replay = {"idx": [idx[k] for k in range(len(order))]}
        x = g["x_T"].cuda()
        res = []
        for k, j in enumerate(order):
            t = s.timesteps[j].clone() if how == "clone" else int(s._timesteps_host[j])
            x = s.step(g[f"eps_{k}"].cuda(), t, x, return_dict=False)[0]
            res.append(x.clone())
        outs[how] = res
        rec = s.trajectory(skip_first=False)["x"][0].cpu()
        want = torch.tensor([[float(s._timesteps_host[j]), float(s._timesteps_host[j] - s._stride)] for j in order])
        assert torch.equal(rec, want), "the rollout record must hold the condition rows the steps actually used"
    for a, b in zip(outs["clone"], outs["int"]):
        assert torch.equal(a, b)


def test_reading_conds_epsilon_after_the_ring_was_reused_raises():
    g, m, s, _ = _pair()
    s.set_timesteps(m["n"], device="cuda")
    s.replay = {"idx": [g[f"idx_{i}"] for i in range(m["n"])]}
    x = g["x_T"].cuda()
    kept = []
    od = m["config"]["order_dim"]
    for i, t in enumerate(s.timesteps):
        out = s.step_cfg(g[f"pair_{i}"].cuda(), t, x, m["guidance"])
        x = out[0]
        kept.append(out[3])
        if i == 2:      # step 0 referenced one slot; the ring (order_dim slots) has not wrapped yet: still readable
            e0 = kept[0]["epsilon"]
            assert e0.shape[1] == od and torch.equal(e0[:, 0].cpu(), g["eps_0"]) and not e0[:, 1:].any()
    last = kept[-1]
    assert len(last) == 2 and set(iter(last)) == {"x", "epsilon"} and set(dict(last)) == {"x", "epsilon"}
    assert torch.equal(last["epsilon"][:, 0].cpu(), g[f"eps_{m['n'] - 1}"])
    assert torch.equal(kept[0]["epsilon"], e0)                    # already materialised: a copy, stays valid
    for stale in (kept[1], kept[-2]):                             # never read in time: their slots have been rewritten
        with pytest.raises(RuntimeError, match="history ring had been overwritten"):
            stale["epsilon"]


def test_step_cfg_rejects_a_destination_it_cannot_write_densely():
    g, m, s, _ = _pair()
    s.set_timesteps(m["n"], device="cuda")
    x = g["x_T"].cuda()
    pair = g["pair_0"].cuda()
    with pytest.raises(ValueError, match="out must have the sample's shape"):
        s.step_cfg(pair, s.timesteps[0], x, 3.0, out=torch.empty(x.shape[0], x[0].numel(), device="cuda"))
    wide = torch.empty(*x.shape[:-1], 2 * x.shape[-1], device="cuda")[..., ::2]
    with pytest.raises(ValueError, match="out must have the sample's shape"):
        s.step_cfg(pair, s.timesteps[0], x, 3.0, out=wide)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_can_drive_two_devices():
    """per-device kernel attributes / SM counts and a device guard around the launches: tensors on cuda:1 while cuda:0 is
    current, with a batch large enough for the >48 KB shared-memory path of the policy kernel"""
    import consolver_b200 as cb
    g, m, _, _ = _pair()
    res = []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)
        s = cb.PPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
        s.factor_net.load_state_dict(g.state_dict)
        s.factor_net.to(dev)
        s.set_timesteps(m["n"], device=dev)
        s.replay = {"idx": [g[f"idx_{i}"].to(dev) for i in range(m["n"])]}
        x = g["x_T"].to(dev)
        for i, t in enumerate(s.timesteps):
            x = s.step_cfg(g[f"pair_{i}"].to(dev), t, x, m["guidance"])[0]
        res.append(x.cpu())
    assert torch.equal(res[0], res[1])


@pytest.mark.parametrize("flow", ["genppo_bf16_scalers", "genppo_f16_scalers", "bf16_out_depth6", "f16_pipeline_depth8",
                                  "f32_depth5_ragged", "fm_bf16_depth6"])
def test_pdl_linked_step_reads_the_coefficient_record_the_policy_kernel_wrote(flow):
    """The step kernel is a programmatic dependent launch of the policy kernel and starts while that one is still
    running; its coefficient loads belong after griddepcontrol.wait.  nvcc had hoisted `__ldg` coefficient loads above
    the wait in the 16-bit depth-1 and the runtime-depth instantiations (stale records, non-deterministic; found by the
    live differential fuzzing).  Trajectories with the two kernels linked (use_pdl, the default) must equal the same
    trajectories with ordinary stream order, every time."""
    import numpy as np

    import consolver_b200 as cb

    fm = flow.startswith("fm")
    od = {"genppo_bf16_scalers": 3, "genppo_f16_scalers": 4, "bf16_out_depth6": 6, "f16_pipeline_depth8": 8,
          "f32_depth5_ragged": 5, "fm_bf16_depth6": 6}[flow]
    mdt = torch.bfloat16 if "bf16" in flow else torch.float16 if "f16" in flow else torch.float32
    xdt = mdt if flow.startswith(("genppo", "f16_pipeline")) else torch.float32
    shape = (3, 5, 7) if "ragged" in flow else ((64, 16) if fm else (4, 16, 16))
    B, n = 5, 9
    fkw = dict(hidden_dim=32, num_actions=11)

    def build():
        if fm:
            s = cb.FMPPOScheduler(shift=3.0, use_dynamic_shifting=True, order_dim=od, scaler_dim=2, mu_dim=0,
                                  factor_net_kwargs=dict(fkw))
        else:
            s = cb.PPOScheduler(order_dim=od, scaler_dim=2, prediction_type="v_prediction", timestep_spacing="trailing",
                                beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012,
                                factor_net_kwargs=dict(embedding_dim=64, **fkw))
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for p in s.factor_net.parameters():
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (0.02 if fm else 0.4))
        s.factor_net.cuda()
        if flow.startswith("genppo"):
            s.factor_net.to("cuda", dtype=mdt)
        return s

    def run(s, pdl):
        s.use_pdl = pdl
        if fm:
            s.set_timesteps(n, device="cuda", sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
            s.set_begin_index(0)
        else:
            s.set_timesteps(n, device="cuda")
        g = torch.Generator().manual_seed(11)
        x = torch.randn(B, *shape, generator=g).to(mdt if fm else xdt).cuda()
        outs = []
        ctx = torch.autocast("cuda", mdt) if flow.startswith("genppo") else __import__("contextlib").nullcontext()
        with ctx, torch.no_grad():
            for i in range(n):
                e = torch.randn(B, *shape, generator=g).to(mdt).cuda()
                torch.manual_seed(100 + i)
                x = s.step(e, s.timesteps[i], x, return_dict=False)[0]
                outs.append(x)
        return outs

    base = run(build(), False)
    for rep in range(12):
        got = run(build(), True)
        for i, (a, b) in enumerate(zip(got, base)):
            assert a.dtype == b.dtype and torch.equal(a, b), f"{flow}: repetition {rep}, step {i}"
