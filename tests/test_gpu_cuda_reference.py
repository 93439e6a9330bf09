"""GPU parity against the reference AS IT RUNS ON A GPU — the product's default arithmetic.

The `cuda_*` fixtures were written by the unmodified reference executing on a B200 (oracle/make_golden.py cuda, through
the byte-identical copies under oracle/_ref): fp32 flows, 16-bit model outputs, gen_ppo.py's shipped fp16-autocast
inference flow (policy and bins cast to fp16, everything under torch.autocast), the autocast training rollouts
(train_ppo.py:353, edit_ppo/train_ppo.py:289), FM/FLUX flows and use_conv.  What they pin and a CPU run cannot: ATen's
CUDA treatment of the host-resident schedule scalars (one rounding per product, `t / s` as `t * (1/s)` — this moves
fp32 latents by an ulp), autocast (fp16/bf16 Linear layers, fp32 torch.sum) and cuBLAS in the policy MLP.

Bars: latents bit-identical INCLUDING their dtype at every step; the scheduler's OWN categorical draw (its MLP + the
fixture's Exp(1) values) picks the reference's indices; actions / masks / condition rows bit-identical; probabilities
within tolerances derived from the measured spreads in profiles/parity_spread_r02.md (see PROB_TOL)."""
import contextlib

import numpy as np
import pytest
import torch

import abi_helpers as ah
import consolver_oracle as orc
from golden_io import Golden, names

pytestmark = pytest.mark.gpu

_DT = {None: None, "float16": torch.float16, "bfloat16": torch.bfloat16}

# (rtol, atol) on the softmax tables, = 2x the worst spread MEASURED on the B200 over all fixtures of the class
# (profiles/parity_spread_r02.md, tools/parity_spread.py), never a round guess:
PROB_TOL = {
    "sd_f32": (0.0, 1e-6),          # SD fp32 policy: the north-star bar; measured worst |dp| 2.1e-7
    "fm_f32": (8e-6, 1e-6),         # FM fp32 policy, softmax temperature 0.01: measured worst rel 3.8e-6, abs 4.8e-7
                                    #   (the reference's own CPU-vs-CUDA spread on the same weights is 7.7e-6 rel)
    "autocast": (1.2e-4, 6e-6),     # 16-bit Linear layers (cuBLAS vs the kernel's fp64 butterfly): measured worst rel
                                    #   5.8e-5, abs 3.0e-6 (cuda_rollout_f16_*); every other autocast fixture <= 6e-8 abs
    "conv16": (3.4e-3, 5e-4),       # use_conv on bf16 outputs: features agree to bf16 precision; measured 1.7e-3 / 2.3e-4
}


def _tol(m):
    if m.get("autocast") or m.get("policy_dtype"):
        return PROB_TOL["autocast"]
    return PROB_TOL["sd_f32" if m["kind"] == "sd" else "fm_f32"]


def _scheduler(g):
    import consolver_b200 as cb

    m = g.meta
    if m["kind"] == "sd":
        s = cb.PPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
    else:
        s = cb.FMPPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
    assert s.reference_device == "cuda"                     # the product default
    s.factor_net.load_state_dict(g.state_dict)
    if m.get("policy_dtype"):
        s.factor_net.to("cuda", dtype=_DT[m["policy_dtype"]])          # gen_ppo.py:194-195
    else:
        s.factor_net.cuda()
    if m["kind"] == "sd":
        s.set_timesteps(m["n"], device="cuda")
    else:
        s.set_timesteps(m["n"], device="cuda", sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
        if m["use_begin_index"]:
            s.set_begin_index(0)
    return s


def _ctx(m):
    return torch.autocast("cuda", _DT[m["autocast"]]) if m.get("autocast") else contextlib.nullcontext()


def _check(g, i, s, out, rtol, atol):
    x, actions, probs, conds, masks = out
    lp = s.last_policy()
    conv = g.meta["config"].get("use_conv", False)
    ref_tab = g[f"probs_full_{i}"] if conv else g[f"probs_full_{i}"][0]
    torch.testing.assert_close(lp["probs_table"].cpu(), ref_tab, rtol=rtol, atol=atol)
    assert torch.equal(lp["idx"].cpu(), g[f"idx_{i}"]), f"step {i}: the scheduler's own draw differs from the reference's"
    ref_a = g[f"actions_{i}"]
    assert actions.dtype == ref_a.dtype and torch.equal(actions.cpu(), ref_a), f"step {i}: actions"
    assert torch.equal(masks.cpu(), g[f"masks_{i}"])
    assert torch.equal(conds["x"].cpu(), g[f"condx_{i}"])
    torch.testing.assert_close(probs.cpu(), g[f"probs_{i}"], rtol=rtol, atol=atol)
    ref = g[f"prev_{i}"]
    assert x.dtype == ref.dtype, f"step {i}: latent dtype {x.dtype} vs the reference's {ref.dtype}"
    assert torch.equal(x.cpu(), ref), f"step {i}: latent not bit-identical to the reference run on the GPU"


SD_CUDA = [n for n in names("cuda_") if Golden(n).meta["kind"] == "sd"]
FM_CUDA = [n for n in names("cuda_") if Golden(n).meta["kind"] == "fm"]
# use_conv on 16-bit model outputs: the reference evaluates cosine_similarity in bf16 arithmetic; see the dedicated test
FM_CONV16 = [n for n in FM_CUDA if Golden(n).meta["config"].get("use_conv") and Golden(n).meta["dtype"] != "float32"]


@pytest.mark.parametrize("mode", ["plain", "cfg_fused"])
@pytest.mark.parametrize("name", SD_CUDA)
def test_sd_scheduler_matches_the_reference_run_on_a_gpu(name, mode):
    g = Golden(name)
    m = g.meta
    s = _scheduler(g)
    s.replay = {"q": [g[f"q_{i}"].cuda() for i in range(m["n"])]}
    x = g["x_T"].cuda()
    rtol, atol = _tol(m)
    if m["config"].get("use_conv"):
        rtol, atol = max(rtol, 1e-5), max(atol, 2e-6)       # features are reductions over the latent
    with _ctx(m):
        for i, t in enumerate(s.timesteps):
            if mode == "plain":
                out = s.step(g[f"eps_{i}"].cuda(), t, x, return_dict=False)
            else:
                out = s.step_cfg(g[f"pair_{i}"].cuda(), t, x, m["guidance"])
                assert torch.equal(s.ets[-1].cpu(), g[f"eps_{i}"])           # the ring slot == the caller-side combine
            _check(g, i, s, out, rtol, atol)
            x = out[0]


@pytest.mark.parametrize("name", [n for n in FM_CUDA if n not in FM_CONV16])
def test_fm_scheduler_matches_the_reference_run_on_a_gpu(name):
    g = Golden(name)
    m = g.meta
    s = _scheduler(g)
    s.replay = {"q": [g[f"q_{i}"].cuda() for i in range(m["n"])]}
    x = g["x_T"].cuda()
    rtol, atol = _tol(m)
    with _ctx(m):
        for i, t in enumerate(s.timesteps):
            out = s.step(g[f"v_{i}"].cuda(), t, x, return_dict=False)
            _check(g, i, s, out, rtol, atol)
            x = out[0]


@pytest.mark.parametrize("name", FM_CONV16)
def test_fm_use_conv_on_bf16_outputs_samples_its_own_actions(name):
    """use_conv with 16-bit model outputs.  The reference computes the cosine features in bf16 arithmetic (every op of
    F.cosine_similarity rounded to 8 bits); the feature kernel reduces in fp32/fp64, so the features agree to bf16
    precision only and the tables to ~2e-3 relative (softmax temperature 0.01).  The scheduler draws its OWN actions from
    the fixture's Exp(1) values — nothing is replayed from the reference.  Measured on the B200
    (profiles/parity_spread_r02.md): 0 of 36 sampled indices differ, so the whole trajectory is bit-identical."""
    g = Golden(name)
    m = g.meta
    s = _scheduler(g)
    s.replay = {"q": [g[f"q_{i}"].cuda() for i in range(m["n"])]}
    x = g["x_T"].cuda()
    for i, t in enumerate(s.timesteps):
        out = s.step(g[f"v_{i}"].cuda(), t, x, return_dict=False)
        _check(g, i, s, out, *PROB_TOL["conv16"])
        x = out[0]


# ---- kernel level, default (CUDA-tensor) rules against the oracle with sem=CUDA ----------------------------------------
def _rand_coef(B, od, g):
    c = torch.randn(B, od + 2, generator=g)
    c[:, od:] = 1 + 0.05 * torch.randn(B, 2, generator=g)
    return c


def _oracle(eps, hist, x, c, od, scalars, vpred, sdim):
    n_hist = len(hist) + 1
    coef = None if n_hist == 1 else [c[:, j] for j in range(n_hist)]
    scale = [c[:, od + j] for j in range(sdim)]
    eff, xs = orc.combine_history([eps] + hist, coef, scale, x)
    sc = [torch.tensor(v, dtype=torch.float32) for v in scalars]
    return orc.ddim_update(xs, eff, sc, "v_prediction" if vpred else "epsilon", sem=orc.CUDA)


@pytest.mark.parametrize("B,shape", [(3, (4, 8, 8)), (2, (3, 5, 7)), (1, (4, 64, 64)), (160, (4, 64, 64))])
@pytest.mark.parametrize("n_hist", [1, 2, 4, 6])
@pytest.mark.parametrize("vpred,sdim", [(False, 0), (True, 2), (False, 1)])
def test_step_sd_f32_bit_exact_with_cuda_rules(B, shape, n_hist, vpred, sdim):
    """fp32: `(x - sb*e) / sa` is `(x - sb*e) * (1/sa)` on CUDA tensors.  B=160 is a multi-wave grid and is launched with
    the legacy `unroll = 2` knob (accepted, runs the one-vector form — the only one compiled); (3,5,7) the scalar path."""
    from consolver_b200 import _lib

    lib = _lib.load()
    od = max(n_hist, 4)
    g = torch.Generator().manual_seed(n_hist * 10 + B)
    rn = lambda: torch.randn(B, *shape, generator=g)  # noqa: E731
    e0, cond, x = rn(), rn(), rn()
    hist = [rn() for _ in range(n_hist - 1)]
    c = _rand_coef(B, od, g)
    scalars = (0.8378, 0.5460, 0.9151, 0.4033)
    flags = (1 if vpred else 0) | (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0)
    eps = orc.cfg_combine(e0, cond, 3.0)
    ref = _oracle(eps, hist, x, c, od, scalars, vpred, sdim)
    assert lib.consolver_set_step_launch(0, 2 if B == 160 else 0) == 0
    try:
        out, slot = ah.step_sd(e0.cuda(), cond.cuda(), 3.0, [h.cuda() for h in hist], x.cuda(), c.cuda(), od, scalars,
                               flags, slot=True, host=False)
    finally:
        lib.consolver_set_step_launch(0, 0)
    assert torch.equal(slot.cpu(), eps)
    assert torch.equal(out.cpu(), ref)
    host_out, _ = ah.step_sd(e0.cuda(), cond.cuda(), 3.0, [h.cuda() for h in hist], x.cuda(), c.cuda(), od, scalars,
                             flags, host=True)
    assert (host_out.cpu() - ref).abs().max() <= 1e-5 * ref.abs().max()      # the two rule sets agree to ulps


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("x32", [False, True])
@pytest.mark.parametrize("n_hist,vpred,sdim", [(1, False, 0), (1, True, 0), (3, False, 0), (4, True, 2), (1, False, 1)])
def test_step_sd_16bit_outputs_with_cuda_rules(dtype, x32, n_hist, vpred, sdim):
    """16-bit model outputs, fp32 policy.  All-16-bit latents are only requested while the estimate is the raw output
    (n_hist == 1, no scalers); every other combination has fp32 latents (torch promotion)."""
    if not x32 and (n_hist > 1 or sdim):
        pytest.skip("the reference's latent is fp32 there")
    from consolver_b200 import _lib

    B, shape, od = 3, (4, 16, 16), 4
    g = torch.Generator().manual_seed(11 + n_hist)
    rn = lambda: torch.randn(B, *shape, generator=g).to(dtype)  # noqa: E731
    e0, cond, x = rn(), rn(), rn()
    hist = [rn() for _ in range(n_hist - 1)]
    c = _rand_coef(B, od, g)
    scalars = (0.8378, 0.5460, 0.9151, 0.4033)
    flags = (1 if vpred else 0) | (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0)
    eps = orc.cfg_combine(e0, cond, 3.0)
    xin = x.float() if x32 else x
    ref = _oracle(eps, hist, xin, c, od, scalars, vpred, sdim)
    assert ref.dtype == (torch.float32 if x32 else dtype)
    out, _ = ah.step_sd(e0.cuda(), cond.cuda(), 3.0, [h.cuda() for h in hist], xin.cuda(), c.cuda(), od, scalars, flags,
                        host=False)
    assert torch.equal(out.cpu(), ref)
    if x32 and sdim < 2:
        # the promotion step of a 16-bit pipeline: the sample is still a 16-bit tensor in the reference (X_WAS_LOWP)
        ref2 = _oracle(eps, hist, x, c, od, scalars, vpred, sdim)
        if ref2.dtype == torch.float32:
            out2, _ = ah.step_sd(e0.cuda(), cond.cuda(), 3.0, [h.cuda() for h in hist], xin.cuda(), c.cuda(), od,
                                 scalars, flags | _lib.FLAG_X_WAS_LOWP, host=False)
            assert torch.equal(out2.cpu(), ref2)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n_hist,vpred,sdim,x_state", [(1, False, 0, "lowp"), (1, True, 2, "lowp"), (2, False, 0, "was_lowp"),
                                                       (2, True, 2, "was_lowp"), (3, False, 1, "f32"), (4, True, 2, "f32")])
def test_step_sd_16bit_coefficient_tensors(dtype, n_hist, vpred, sdim, x_state):
    """CONSOLVER_FLAG_LOWP_COEF — the arithmetic of gen_ppo.py's flow: coefficients and scalers are 16-bit tensors, the
    closing coefficient is fp32.  Checked against torch's own evaluation (oracle) with tensors of those dtypes."""
    from consolver_b200 import _lib

    B, shape, od = 2, (4, 16, 16), 4
    g = torch.Generator().manual_seed(5 + n_hist)
    rn = lambda: torch.randn(B, *shape, generator=g).to(dtype)  # noqa: E731
    eps, x = rn(), rn()
    hist = [rn() for _ in range(n_hist - 1)]
    act = (torch.randn(B, od + 1, generator=g) * 0.5).to(dtype)               # bin values as 16-bit tensors
    coef, scale = orc.coefficients(act, n_hist, od, sdim, sem=orc.TorchSemantics("cuda", dtype))
    if coef is not None:
        assert coef[-1].dtype == torch.float32 and all(cj.dtype == dtype for cj in coef[:-1])
    c = torch.zeros(B, od + 2)
    c[:, od:] = 1
    for j, cj in enumerate(coef or []):
        c[:, j] = cj.float()
    for j, sj in enumerate(scale):
        c[:, od + j] = sj.float()
    xin = x if x_state != "f32" else torch.randn(B, *shape, generator=g)
    eff, xs = orc.combine_history([eps] + hist, coef, scale, xin)
    sc = [torch.tensor(v, dtype=torch.float32) for v in (0.8378, 0.5460, 0.9151, 0.4033)]
    ref = orc.ddim_update(xs, eff, sc, "v_prediction" if vpred else "epsilon", sem=orc.CUDA)
    flags = (1 if vpred else 0) | (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0) | _lib.FLAG_LOWP_COEF
    if x_state == "lowp":
        assert ref.dtype == dtype
        x_dev = x.cuda()
    else:
        assert ref.dtype == torch.float32
        x_dev = xin.float().cuda()
        flags |= _lib.FLAG_X_WAS_LOWP if x_state == "was_lowp" else 0
    out, _ = ah.step_sd(eps.cuda(), None, 0.0, [h.cuda() for h in hist], x_dev, c.cuda(), od,
                        (0.8378, 0.5460, 0.9151, 0.4033), flags, host=False)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("variant", ["sd", "fm"])
@pytest.mark.parametrize("act", [None, torch.float16, torch.bfloat16])
def test_policy_table_with_cuda_rules_and_autocast(variant, act):
    """x/999 and logits/0.01 as reciprocal multiplies; under autocast the Linear layers run in 16 bit."""
    from consolver_b200 import _lib
    from test_gpu_kernels import make_sd

    sd = make_sd(variant, 256, 11, 4, 2 if variant == "sd" else 0, 0, seed=3, last_std=0.5 if variant == "sd" else 0.02)
    rows = torch.tensor([[999.0, 874.0], [499.0, 374.0], [124.0, -1.0]]) if variant == "sd" else \
        torch.tensor([[1.0, 0.9567], [0.7595, 0.6546], [0.3109, 0.0]])
    sem = orc.TorchSemantics("cuda", act)
    ref = orc.policy_probs(sd, rows, variant, sem=sem)
    pf = 0 if act is None else (_lib.POLICY_ACT_F16 if act == torch.float16 else _lib.POLICY_ACT_BF16)
    dsd = ah.sd_to_dev({k: (v.to(act).float() if act is not None and k != "action_values" else v) for k, v in sd.items()})
    got = ah.policy_table(dsd, rows, 999.0 if variant == "sd" else 1.0, 1.0 if variant == "sd" else 0.01, host=False,
                          policy_flags=pf)
    rtol, atol = PROB_TOL["autocast"] if act is not None else PROB_TOL["sd_f32" if variant == "sd" else "fm_f32"]
    torch.testing.assert_close(got.cpu(), ref, rtol=rtol, atol=atol)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n_hist,sdim", [(1, 0), (2, 2), (3, 1), (4, 2)])
def test_policy_sample_writes_16bit_coefficient_records(dtype, n_hist, sdim):
    """CONSOLVER_POLICY_COEF_*: a0+1 and 1+s rounded to the bins' dtype, closing coefficient = fp32 1 - sum"""
    from consolver_b200 import _lib

    B, od, K = 37, 4, 11
    A = od + sdim - 1
    g = torch.Generator().manual_seed(n_hist)
    av = orc.action_value_table("sd", K, od, sdim).to(dtype)
    table = torch.softmax(torch.randn(A, K, generator=g), -1)
    q = torch.empty(B * A, K).exponential_(1, generator=g)
    dsd = {"action_values": av.float().cuda()}
    pf = _lib.POLICY_COEF_F16 if dtype == torch.float16 else _lib.POLICY_COEF_BF16
    out = ah.policy_sample(dsd, table.cuda(), B, od, sdim, n_hist, q=q.cuda(), policy_flags=pf)
    idx = orc.sample_indices(table.unsqueeze(0).expand(B, A, K), q)
    assert torch.equal(out["idx"].cpu(), idx)
    actions = av[torch.arange(A), idx]                                           # 16-bit tensor, like the reference's
    assert torch.equal(out["actions"].cpu(), actions.float())
    coef, scale = orc.coefficients(actions, n_hist, od, sdim, sem=orc.TorchSemantics("cuda", dtype))
    c = out["coef"].cpu()
    for j, cj in enumerate(coef or []):
        assert torch.equal(c[:, j], cj.float()), f"coef {j}"
    for j, sj in enumerate(scale):
        assert torch.equal(c[:, od + j], sj.float())


def test_default_and_cpu_reference_devices_differ_only_by_ulps():
    """The same trajectory under both rule sets: different bits (that is why both are pinned), same numbers to 1e-5."""
    import consolver_b200 as cb

    g = Golden("cuda_sd_eps_s0_n8_B3")
    m = g.meta
    outs = {}
    for dev in ("cuda", "cpu"):
        s = cb.PPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
        s.reference_device = dev
        s.factor_net.load_state_dict(g.state_dict)
        s.factor_net.cuda()
        s.set_timesteps(m["n"], device="cuda")
        s.replay = {"idx": [g[f"idx_{i}"] for i in range(m["n"])]}
        x = g["x_T"].cuda()
        for i, t in enumerate(s.timesteps):
            x = s.step_cfg(g[f"pair_{i}"].cuda(), t, x, m["guidance"])[0]
        outs[dev] = x.cpu()
    ref = g[f"prev_{m['n'] - 1}"]
    assert torch.equal(outs["cuda"], ref)
    assert not torch.equal(outs["cpu"], ref)
    assert (outs["cpu"] - ref).abs().max() <= 1e-4 * ref.abs().max()      # north-star bar on the final latent


@pytest.mark.parametrize("B", [1, 2, 37])
@pytest.mark.parametrize("od", [3, 4, 6, 8])
@pytest.mark.parametrize("host", [False, True])
def test_closing_coefficient_is_summed_in_atens_order(B, od, host):
    """`1 - torch.sum(torch.stack(terms), dim=0)` (scheduler_ppo.py:172): left to right on CPU tensors; on CUDA tensors
    ATen's reduce order — four accumulators per sample for B >= 2, a two-level tree when B == 1 (where the reduced
    dimension is the fastest one).  The restated order is checked against torch.sum ON THIS DEVICE and against the
    kernel's coefficient records, for every history depth up to order_dim."""
    from consolver_b200 import _lib

    if host and od > 5:
        pytest.skip("CPU-tensor rules are restated (left to right) for up to 4 terms, i.e. order_dim <= 5: torch's CPU "
                    "reduction of more terms uses several accumulators and no CPU-made fixture has order_dim > 4")
    K, sdim = 11, 1
    A = od + sdim - 1
    g = torch.Generator().manual_seed(B * 10 + od)
    av = (torch.randn(A, K, generator=g) * 0.9)                       # generic values: rounding differences are common
    table = torch.softmax(torch.randn(A, K, generator=g), -1)
    dsd = {"action_values": av.cuda()}
    sem = orc.HOST if host else orc.CUDA
    for n_hist in range(1, od + 1):
        q = torch.empty(B * A, K).exponential_(1, generator=g)
        out = ah.policy_sample(dsd, table.cuda(), B, od, sdim, n_hist, q=q.cuda(),
                               policy_flags=_lib.POLICY_HOST_DIV if host else 0)
        idx = orc.sample_indices(table.unsqueeze(0).expand(B, A, K), q)
        actions = av[torch.arange(A), idx]
        coef, scale = orc.coefficients(actions, n_hist, od, sdim, sem=sem)
        c = out["coef"].cpu()
        for j, cj in enumerate(coef or []):
            assert torch.equal(c[:, j], cj), f"n_hist {n_hist} coef {j}"
        if coef is not None and not host:
            # ... and the restated order IS what torch.sum does on this GPU with the reference's [B,1,1,1] tensors
            terms = [(actions[:, 0] + 1)] + [actions[:, i] for i in range(1, n_hist - 1)]
            stacked = torch.stack([t.view(B, 1, 1, 1) for t in terms]).cuda()
            ref_last = (1 - torch.sum(stacked, dim=0)).flatten().cpu()
            assert torch.equal(coef[-1], ref_last), f"n_hist {n_hist}: oracle order != torch.sum on cuda"
