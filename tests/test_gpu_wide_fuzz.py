"""GPU: randomly drawn configurations for the parts the live-reference fuzz (tests/test_gpu_live_reference.py) does not
reach — the CUDA-graph forms (PDL-chained step kernels, policy on a side stream) against eager execution, the native PPO
loss/gradient kernel against torch autograd, the fixed-coefficient baseline solvers against the oracle's arithmetic.
CONSOLVER_FUZZ_CASES scales the number of cases."""
import contextlib
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = int(os.environ.get("CONSOLVER_FUZZ_CASES", "24"))


def _seed_policy(fn, seed, last_std):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in fn.named_parameters():
            if name.startswith("mlp.4"):
                p.copy_(torch.randn(p.shape, generator=g) * (last_std if name.endswith("weight") else 0.1))
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.3)


@pytest.mark.parametrize("case", range(max(CASES // 2, 1)))
def test_graphed_previews_equal_eager_execution_over_random_configurations(case):
    """GraphedPreview (one CUDA graph per preview: table kernel, sample kernels on a side stream, step kernels chained as
    programmatic dependent launches that read the previous step's latent and ring slot after griddepcontrol.wait) must
    give the bits of eager stepping, generator consumption included — for every history depth, scaler count, dtype flow
    and for FM, not only the production configuration."""
    import consolver_b200 as cb
    from consolver_b200.denoise import GraphedPreview, preview_from_outputs, preview_from_pairs

    rng = random.Random(11000 + case)
    kind = rng.choice(["sd", "sd", "fm"])
    od = rng.choice([2, 3, 4, 4, 5, 6, 8])
    K, hidden = rng.choice([3, 11, 161]), rng.choice([16, 64, 256])
    n, B = rng.choice([1, 2, 5, 8, 12]), rng.choice([1, 2, 5, 33, 64])
    fkw = dict(hidden_dim=hidden, num_actions=K)
    if kind == "sd":
        cfg = dict(order_dim=od, scaler_dim=rng.choice([0, 0, 1, 2]),
                   prediction_type=rng.choice(["epsilon", "v_prediction"]), timestep_spacing="trailing",
                   beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, steps_offset=1)
        shape = rng.choice([(4, 8, 8), (3, 5, 7), (4, 32, 32)] + ([(4, 64, 64)] if B <= 5 else []))
        flow = rng.choice(["f32", "f32", "bf16_out", "f16_out", "f16_pipeline", "bf16_pipeline"])
        mk = lambda: cb.PPOScheduler(factor_net_kwargs=dict(embedding_dim=64, **fkw), **cfg)  # noqa: E731
        tk = {}
    else:
        cfg = dict(shift=3.0, use_dynamic_shifting=True, order_dim=od, scaler_dim=rng.choice([0, 0, 2]), mu_dim=0)
        shape = rng.choice([(64, 16), (5, 7), (256, 64)])
        flow = rng.choice(["f32", "bf16_pipeline", "bf16_pipeline", "f16_pipeline"])
        mk = lambda: cb.FMPPOScheduler(factor_net_kwargs=dict(fkw), **cfg)  # noqa: E731
        tk = dict(sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
    mdt = torch.float32 if flow == "f32" else torch.float16 if "f16" in flow and "bf16" not in flow else torch.bfloat16
    xdt = mdt if flow.endswith("_pipeline") else torch.float32
    guided = kind == "sd" and rng.choice([True, True, False])
    s, e = mk(), mk()
    _seed_policy(s.factor_net, case, 0.5 if kind == "sd" else 0.02)
    e.factor_net.load_state_dict(s.factor_net.state_dict())
    s.factor_net.cuda(), e.factor_net.cuda()
    g = torch.Generator().manual_seed(case)
    x = torch.randn(B, *shape, generator=g).to(xdt).cuda()
    outs = [torch.randn((2 * B if guided else B), *shape, generator=g).to(mdt).cuda() for _ in range(n)]
    gp = GraphedPreview(s, x, outs, 3.0 if guided else None, n, set_timesteps_kwargs=tk)
    tag = f"graph case {case} ({kind}, {flow}, K={K}, H={hidden}, n={n}, B={B}, {shape}, guided={guided}, {cfg})"
    torch.manual_seed(500 + case)
    got = [gp.replay().clone() for _ in range(3)]          # back to back: replay k+1 is enqueued while k still runs
    idx = gp.record()["idx"].clone() if n > 1 else None
    state_after_graph = torch.cuda.get_rng_state()
    torch.manual_seed(500 + case)
    for k in range(3):
        e.set_timesteps(n, device="cuda", **tk)
        if hasattr(e, "set_begin_index"):
            e.set_begin_index(0)
        ref = preview_from_pairs(e, x, outs, 3.0) if guided else preview_from_outputs(e, x, outs)
        assert ref.dtype == got[k].dtype and torch.equal(ref, got[k]), tag + f": replay {k}"
    if idx is not None:
        assert torch.equal(e.trajectory()["idx"], idx), tag + ": indices of the last replay"
    assert torch.equal(torch.cuda.get_rng_state(), state_after_graph), tag + ": generator consumed differently"


@pytest.mark.parametrize("case", range(max(CASES // 2, 1)))
def test_native_ppo_kernel_matches_autograd_over_random_shapes(case):
    """csrc/ppo.cu vs torch autograd of the same loss (train_ppo.py:406-427) for random policy widths, bin counts, action
    dims, rollout lengths, batch sizes, clip ranges and entropy weights, SD and FM policies."""
    import consolver_b200 as cb
    from consolver_b200 import ppo

    rng = random.Random(12000 + case)
    kind = rng.choice(["sd", "sd", "fm"])
    od, sc = rng.choice([2, 3, 4, 6]), rng.choice([0, 1, 2])
    H, K = rng.choice([16, 64, 256]), rng.choice([3, 11, 161])
    B, R = rng.choice([1, 2, 7, 48, 80]), rng.choice([1, 2, 7, 14])
    torch.manual_seed(case)
    if kind == "sd":
        fn = cb.FactorNetPPO(hidden_dim=H, num_actions=K, order_dim=od, scaler_dim=sc)
        with torch.no_grad():
            fn.mlp[4].weight.normal_(0, 0.3)
            fn.mlp[4].bias.normal_(0, 0.1)
        t = torch.tensor(sorted(rng.sample(range(70, 1000), R), reverse=True), dtype=torch.float32)
        x_rows = torch.stack([t, t - 66], 1)
    else:
        fn = cb.FactorNetPPOFM(hidden_dim=H, num_actions=K, order_dim=od, scaler_dim=sc, mu_dim=rng.choice([0, 1]))
        with torch.no_grad():
            fn.mlp[4].weight.normal_(0, 0.003)       # temperature 0.01: keep the softmax away from one-hot
        x_rows = torch.rand(R, 2)
    fn.cuda()
    x_rows = x_rows.cuda()
    flat = ppo.FlatParams(fn)
    A = fn.action_dims
    g = torch.Generator(device="cuda").manual_seed(case)
    idx = torch.randint(0, K, (B, R, A), device="cuda", generator=g)
    with torch.no_grad():
        tables = fn.forward_({"x": x_rows})
    old = tables.unsqueeze(0).expand(B, R, A, K).gather(3, idx.unsqueeze(-1)).squeeze(-1)
    old = (old * (1 + 0.3 * torch.randn(old.shape, device="cuda", generator=g))).clamp(1e-4, 1.0)
    masks = (torch.rand(B, R, A, device="cuda", generator=g) > 0.2).float()
    rewards = torch.randn(B, 1, device="cuda", generator=g) if B > 1 else torch.zeros(1, 1, device="cuda")
    adv = ppo.advantages_from_rewards(rewards, masks) if B > 1 else torch.randn(B, R, A, device="cuda", generator=g) * masks
    clip, ent = rng.choice([0.2, 0.05, 0.5]), rng.choice([0.0, 0.01, 0.1])
    flat.zero_grad()
    loss, info = ppo.ppo_loss(fn, x_rows, idx, old, adv, clip, ent)
    loss.backward()
    g_ref = flat.grad.clone()
    flat.grad.zero_()
    st = ppo.ppo_loss_grad_cuda(fn, flat, x_rows, *(t_.transpose(0, 1).contiguous() for t_ in (idx, old, adv)), clip, ent)
    tag = f"ppo case {case} ({kind}, od={od}, sc={sc}, H={H}, K={K}, B={B}, R={R}, clip={clip}, ent={ent})"
    scale = float(g_ref.abs().max())
    torch.testing.assert_close(flat.grad, g_ref, rtol=2e-3, atol=max(scale, 1e-12) * 2e-4, msg=lambda m: tag + "\n" + m)
    torch.testing.assert_close(st[0], loss.detach(), rtol=1e-4, atol=1e-5, msg=lambda m: tag + "\n" + m)
    torch.testing.assert_close(st[2], info["entropy"], rtol=1e-4, atol=1e-6, msg=lambda m: tag + "\n" + m)
    torch.testing.assert_close(st[3], info["ratio_mean"], rtol=1e-4, atol=1e-6, msg=lambda m: tag + "\n" + m)


@pytest.mark.parametrize("case", range(max(CASES // 2, 1)))
def test_fixed_coefficient_solvers_match_the_same_arithmetic_in_torch(case):
    """DDIM / Adams-Bashforth baselines are the fused step with fixed coefficients: compare with the
    reference's op sequence (scheduler_ppo.py:263-280,:306-332, CUDA-tensor rules) written out in torch on the GPU."""
    import consolver_b200 as cb
    from consolver_b200 import baselines

    rng = random.Random(13000 + case)
    kind = rng.choice(["ddim", "ab2", "ab3", "ab4"])
    pred = rng.choice(["epsilon", "v_prediction"])
    n, B = rng.choice([1, 3, 8, 20]), rng.choice([1, 3, 17])
    shape = rng.choice([(4, 8, 8), (3, 5, 7), (4, 32, 32)])
    kw = dict(prediction_type=pred, beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012,
              timestep_spacing="trailing", steps_offset=1)
    s = baselines.ddim_solver(**kw) if kind == "ddim" else baselines.multistep_solver(int(kind[2]), **kw)
    s.factor_net.cuda()
    s.set_timesteps(n, device="cuda")
    g = torch.Generator().manual_seed(case)
    x = torch.randn(B, *shape, generator=g).cuda()
    xr = x.clone()
    hist = []
    ac = s.alphas_cumprod
    for i in range(n):
        e = torch.randn(B, *shape, generator=g).cuda()
        t = int(s.timesteps[i])
        x = s.step(e, s.timesteps[i], x, return_dict=False)[0]
        hist = ([e] + hist)[: s.fixed_depth or 1]
        c = [float(v) for v in s.fixed_coefficients(len(hist))]
        if len(hist) == 1:
            eff = hist[0]
        else:
            eff = sum(torch.full((B, 1, 1, 1), ci, device="cuda") * hi for ci, hi in zip(c, hist))
        pt = t - 1000 // n
        a_t, a_p = ac[t], ac[pt] if pt >= 0 else ac[0]
        if pred == "v_prediction":
            eff = (a_t ** 0.5) * eff + ((1 - a_t) ** 0.5) * xr
        x0 = (xr - ((1 - a_t) ** 0.5) * eff) / (a_t ** 0.5)
        xr = (a_p ** 0.5) * x0 + ((1 - a_p) ** 0.5) * eff
        assert torch.equal(x, xr), f"baseline case {case} ({kind}, {pred}, n={n}, B={B}, {shape}) step {i}"
