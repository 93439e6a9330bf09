"""CPU: the oracle (oracle/consolver_oracle.py, CPU-tensor rules) next to the UNMODIFIED reference running on CPU tensors,
over randomly drawn configurations — the checker itself checked beyond the committed fixtures, in the same wide space the
GPU fuzz walks (tests/test_gpu_live_reference.py).  Both draw from the default CPU generator under the same seed, so the
oracle's own categorical draw must select the reference's indices.  Runs where a reference tree exists (this container's
/root/reference, or the staged oracle/_ref); CONSOLVER_FUZZ_CASES scales the count."""
import os
import random

import numpy as np
import pytest
import torch

import consolver_oracle as orc
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")

CASES = int(os.environ.get("CONSOLVER_FUZZ_CASES", "16"))


def _seed_policy(fn, seed, last_std):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in fn.named_parameters():
            if name.startswith("mlp.4"):
                p.copy_(torch.randn(p.shape, generator=g) * (last_std if name.endswith("weight") else 0.1))
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.3)


@pytest.mark.parametrize("case", range(CASES))
def test_sd_oracle_next_to_the_reference_on_cpu(case):
    ref = ref_shim.load_reference()
    rng = random.Random(21000 + case)
    od = rng.choice([2, 3, 4, 4, 5, 6, 8])
    cfg = dict(order_dim=od, scaler_dim=rng.choice([0, 0, 1, 2]),
               prediction_type=rng.choice(["epsilon", "epsilon", "v_prediction"]),
               timestep_spacing=rng.choice(["trailing", "leading", "linspace"]),
               beta_schedule=rng.choice(["scaled_linear", "linear", "squaredcos_cap_v2"]),
               beta_start=0.00085, beta_end=0.012, steps_offset=rng.choice([0, 1]), use_conv=rng.choice([False, False, True]))
    K, hidden = rng.choice([3, 11, 11, 161]), rng.choice([16, 64])
    n, B = rng.choice([1, 2, 4, 8, 15]), rng.choice([1, 2, 5])
    shape = rng.choice([(4, 8, 8), (3, 5, 7), (1, 1, 33), (4, 16, 16)])
    flow = rng.choice(["f32", "f32", "f32", "f16_out", "bf16_out", "f16_pipeline", "bf16_pipeline"])
    if cfg["use_conv"]:
        flow = "f32"          # 16-bit cosine features: covered by fixtures made on the GPU
    guidance = rng.choice([3.0, 7.5, 1.0])
    with ref_shim.quiet():
        r = ref.PPOScheduler(factor_net_kwargs=dict(embedding_dim=64, hidden_dim=hidden, num_actions=K), **cfg)
    _seed_policy(r.factor_net, case, rng.choice([0.5, 0.05, 2.0]))
    o = orc.OracleSDScheduler({k: v.clone() for k, v in r.factor_net.state_dict().items()}, **cfg)
    r.set_timesteps(n, device="cpu")
    o.set_timesteps(n)
    assert torch.equal(o.timesteps, r.timesteps)
    mdt = torch.float32 if flow == "f32" else torch.float16 if "f16" in flow and "bf16" not in flow else torch.bfloat16
    xdt = mdt if flow.endswith("_pipeline") else torch.float32
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).to(xdt)
    tag0 = f"cpu sd case {case} ({flow}, K={K}, H={hidden}, n={n}, B={B}, {shape}, g={guidance}, {cfg})"
    for i in range(n):
        pair = torch.randn(2 * B, *shape, generator=g).to(mdt)
        u, c = pair.chunk(2)
        e = u + guidance * (c - u)
        assert torch.equal(orc.cfg_combine(u, c, guidance), e), tag0 + f" step {i}: CFG combine"
        torch.manual_seed(300 + i)
        with ref_shim.quiet(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(e, r.timesteps[i], xr, return_dict=False)
        state = torch.get_rng_state()
        torch.manual_seed(300 + i)
        xo, ao, po, co, mo = o.step(e, o.timesteps[i], xo)
        tag = tag0 + f" step {i}"
        assert torch.equal(ao, ar), tag + ": actions (the oracle's own draw)"
        assert torch.equal(mo, mr) and torch.equal(co["x"], cr["x"]) and torch.equal(co["epsilon"], cr["epsilon"]), tag
        torch.testing.assert_close(po, pr, rtol=0, atol=2e-6)
        assert xo.dtype == xr.dtype and torch.equal(xo, xr), tag + ": latent"
        assert torch.equal(torch.get_rng_state(), state), tag + ": default generator consumed differently"


@pytest.mark.parametrize("case", range(max(CASES // 2, 1)))
def test_fm_oracle_next_to_the_reference_on_cpu(case):
    ref = ref_shim.load_reference()
    rng = random.Random(22000 + case)
    od = rng.choice([2, 2, 3, 4, 6])
    cfg = dict(shift=rng.choice([3.0, 1.0]), use_dynamic_shifting=rng.choice([True, True, False]), order_dim=od,
               scaler_dim=rng.choice([0, 0, 1, 2]), mu_dim=rng.choice([0, 0, 1]),
               time_shift_type=rng.choice(["exponential", "linear"]), invert_sigmas=rng.choice([False, False, True]))
    opt = rng.choice([None, None, "use_karras_sigmas", "use_exponential_sigmas"])
    if opt:
        cfg[opt] = True
    K, hidden = rng.choice([3, 11, 161]), rng.choice([16, 64])
    n, B = rng.choice([1, 2, 5, 8]), rng.choice([1, 2, 4])
    shape = rng.choice([(16, 8), (5, 7), (64, 16)])
    dt = rng.choice([torch.bfloat16, torch.bfloat16, torch.float32, torch.float16])
    begin = rng.choice([0, None])
    with ref_shim.quiet():
        r = ref.FMPPOScheduler(factor_net_kwargs=dict(hidden_dim=hidden, num_actions=K), **cfg)
    _seed_policy(r.factor_net, 50 + case, rng.choice([0.02, 0.002]))
    o = orc.OracleFMScheduler({k: v.clone() for k, v in r.factor_net.state_dict().items()}, **cfg)
    kw = dict(sigmas=np.linspace(1.0, 1 / n, n), mu=1.15) if cfg["use_dynamic_shifting"] else {}
    r.set_timesteps(n, device="cpu", **kw)
    o.set_timesteps(n, **kw)
    assert torch.equal(o.timesteps, r.timesteps) and torch.equal(o.sigmas, r.sigmas)
    if begin is not None or len(set(r.timesteps.tolist())) < n:
        r.set_begin_index(0), o.set_begin_index(0)
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).to(dt)
    for i in range(n):
        v = torch.randn(B, *shape, generator=g).to(dt)
        torch.manual_seed(40 + i)
        with ref_shim.quiet(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(v, r.timesteps[i], xr, return_dict=False)
        torch.manual_seed(40 + i)
        xo, ao, po, co, mo = o.step(v, o.timesteps[i], xo)
        tag = f"cpu fm case {case} (K={K}, H={hidden}, n={n}, B={B}, {shape}, {dt}, begin={begin}, {cfg}) step {i}"
        assert torch.equal(ao, ar), tag + ": actions (the oracle's own draw)"
        assert torch.equal(mo, mr) and torch.equal(co["x"], cr["x"]), tag
        torch.testing.assert_close(po, pr, rtol=2e-5, atol=1e-6)
        assert xo.dtype == xr.dtype and torch.equal(xo, xr), tag + ": latent"


@pytest.mark.parametrize("case", range(max(CASES // 2, 1)))
def test_update_side_loss_and_gradients_next_to_the_reference_modules(case):
    """PPO update side on CPU: the REFERENCE's FactorNetPPO modules evaluate `curr_probs, entropy =
    factor_net(conds, actions)` on the B*(n-1) replicated rows (factor_net_ppo.py:170-184) and the loss lines of
    train_ppo.py:376-427 are written out on top; consolver_b200.ppo.ppo_loss evaluates the n-1 DISTINCT rows and gathers.
    Loss and every parameter gradient must agree, and so must the oracle's restatement, over random widths, bin counts,
    action dims, rollout lengths, masks, clip ranges and entropy weights (SD and FM policies)."""
    import consolver_b200 as cb
    from consolver_b200 import ppo

    ref = ref_shim.load_reference()
    rng = random.Random(23000 + case)
    variant = rng.choice(["sd", "sd", "fm"])
    od, sc = rng.choice([2, 3, 4, 6]), rng.choice([0, 1, 2])
    H, K = rng.choice([16, 64]), rng.choice([3, 11, 161])
    B, n1 = rng.choice([2, 5, 12]), rng.choice([1, 2, 7, 14])
    clip, ent_coef = rng.choice([0.2, 0.05, 0.5]), rng.choice([0.0, 0.01, 0.1])
    with ref_shim.quiet():
        if variant == "sd":
            rfn = ref.FactorNetPPO_SD(hidden_dim=H, num_actions=K, order_dim=od, scaler_dim=sc)
            ofn = cb.FactorNetPPO(hidden_dim=H, num_actions=K, order_dim=od, scaler_dim=sc)
        else:
            mu = rng.choice([0, 1])
            rfn = ref.FactorNetPPO_FM(hidden_dim=H, num_actions=K, order_dim=od, scaler_dim=sc, mu_dim=mu)
            ofn = cb.FactorNetPPOFM(hidden_dim=H, num_actions=K, order_dim=od, scaler_dim=sc, mu_dim=mu)
    _seed_policy(rfn, case, 0.3 if variant == "sd" else 0.004)
    ofn.load_state_dict(rfn.state_dict())
    g = torch.Generator().manual_seed(case)
    A = ofn.action_dims
    if variant == "sd":
        t = torch.tensor(sorted(rng.sample(range(70, 1000), n1), reverse=True), dtype=torch.float32)
        rows = torch.stack([t, t - 66], 1)
    else:
        rows = torch.rand(n1, 2, generator=g)
    idx = torch.randint(0, K, (B, n1, A), generator=g)
    actions = ofn.action_values[torch.arange(A).view(1, 1, A).expand(B, n1, A), idx]
    old = torch.rand(B, n1, A, generator=g) * 0.5 + 0.01
    masks = (torch.rand(B, n1, A, generator=g) > 0.25).float()
    rewards = torch.randn(B, 1, generator=g)
    x = rows.unsqueeze(0).expand(B, n1, 2)
    # the reference, literally (train_ppo.py:376-390 advantages, :406-427 loss) on its own module
    adv = (rewards - rewards.mean()) / (rewards.std() + 1e-8) * 10
    adv = adv.repeat(1, n1).reshape(B * n1, -1) * masks.reshape(B * n1, A)
    cur, entropy = rfn({"x": x.reshape(B * n1, 2)}, actions.reshape(B * n1, A))
    logp = (cur + 1e-9).log().sum(dim=1).unsqueeze(1)
    old_logp = (old.reshape(B * n1, A) + 1e-9).log().sum(dim=1).unsqueeze(1)
    ratio = (logp - old_logp).exp()
    loss_ref = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip, 1 + clip)).mean() - ent_coef * entropy.mean()
    g_ref = torch.autograd.grad(loss_ref, list(rfn.parameters()))
    # the drop-in: distinct rows, recorded indices
    loss, _ = ppo.ppo_loss(ofn, rows, idx, old, ppo.advantages_from_rewards(rewards, masks), clip, ent_coef)
    g_own = torch.autograd.grad(loss, list(ofn.parameters()))
    tag = f"update case {case} ({variant}, od={od}, sc={sc}, H={H}, K={K}, B={B}, n'={n1}, clip={clip}, ent={ent_coef})"
    torch.testing.assert_close(loss, loss_ref.detach(), rtol=2e-5, atol=2e-6, msg=lambda m: tag + "\n" + m)
    for (name, _), a, b in zip(ofn.named_parameters(), g_own, g_ref):
        scale = float(b.abs().max())
        # fp32 summation order differs (B-fold replicated rows summed by autograd vs distinct rows then a gather), and the
        # FM policy's temperature 0.01 multiplies it by 100: the bar is relative to the gradient's own scale
        torch.testing.assert_close(a, b, rtol=2e-3, atol=max(scale, 1e-12) * 2e-4, msg=lambda m: f"{tag}: d/d{name}\n{m}")
    # and the oracle's restatement of the same lines
    sd = {k: v.detach().clone() for k, v in rfn.state_dict().items()}
    loss_orc = orc.ppo_loss_replicated(sd, x, actions, old, masks, rewards, variant, clip, ent_coef)
    torch.testing.assert_close(loss_orc, loss_ref.detach(), rtol=2e-5, atol=2e-6, msg=lambda m: tag + " (oracle)\n" + m)
