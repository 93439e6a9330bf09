"""GPU parity, kernel level: every entry point of the C ABI against the CPU oracle on seeded inputs, plus the
edge cases (ragged / unaligned sizes, every history depth up to CONSOLVER_MAX_ORDER, all scaler / prediction
flags, 16-bit I/O, forced indices, multi-CTA policy grids)."""
import itertools

import pytest
import torch

import abi_helpers as ah
import consolver_oracle as orc

# cpu_reference: these tests check against CPU-made fixtures / the oracle's default (CPU-torch) rules; the product
# default — the reference as executed on CUDA tensors — is covered by tests/test_gpu_cuda_reference.py
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("cpu_reference")]


def make_sd(variant, H, K, order_dim, scaler_dim, mu_dim, seed, last_std):
    g = torch.Generator().manual_seed(seed)
    A = orc.action_dims(variant, order_dim, scaler_dim, mu_dim)
    r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1)  # noqa: E731
    return {
        "mlp.0.weight": r(H, 2) / 2 ** 0.5, "mlp.0.bias": r(H) / 2 ** 0.5,
        "mlp.2.weight": r(H, H) / H ** 0.5, "mlp.2.bias": r(H) / H ** 0.5,
        "mlp.4.weight": torch.randn(A * K, H, generator=g) * last_std, "mlp.4.bias": torch.randn(A * K, generator=g) * 0.1,
        "action_values": orc.action_value_table(variant, K, order_dim, scaler_dim, mu_dim),
    }


POLICY_CASES = [
    # variant, H, K, order_dim, scaler_dim, mu_dim, B
    ("sd", 256, 11, 4, 0, 0, 64), ("sd", 256, 11, 4, 2, 0, 3), ("sd", 256, 161, 4, 2, 0, 5),
    ("sd", 64, 11, 2, 0, 0, 1), ("sd", 100, 7, 3, 1, 0, 9), ("sd", 30, 5, 8, 2, 0, 4), ("sd", 1024, 11, 4, 0, 0, 2),
    ("fm", 256, 11, 2, 0, 0, 10), ("fm", 256, 11, 4, 2, 1, 7), ("sd", 256, 11, 4, 0, 0, 4096), ("sd", 256, 11, 4, 0, 0, 1500),
]


@pytest.mark.parametrize("variant,H,K,od,sdim,mu,B", POLICY_CASES)
def test_policy_kernel_vs_oracle(variant, H, K, od, sdim, mu, B):
    sd = make_sd(variant, H, K, od, sdim, mu, seed=H + K + B, last_std=0.5 if variant == "sd" else 0.02)
    dsd = ah.sd_to_dev(sd)
    A = sd["action_values"].shape[0]
    x = torch.tensor([[874.0, 749.0]]) if variant == "sd" else torch.tensor([[0.9567, 0.9045]])
    g = torch.Generator().manual_seed(B)
    for n_hist in sorted({1, 2, od}):
        q = torch.empty(B * A, K).exponential_(1, generator=g)
        out = ah.policy(dsd, x[0, 0], x[0, 1], 999.0 if variant == "sd" else 1.0, 1.0 if variant == "sd" else 0.01,
                        B, od, sdim, n_hist, q=q.cuda())
        probs = orc.policy_probs(sd, x, variant)                       # [1, A, K]
        rtol = 0 if variant == "sd" else 2e-5      # FM (temperature 0.01): see the measured spreads in test_gpu_golden.py
        torch.testing.assert_close(out["probs_table"].cpu(), probs[0], rtol=rtol, atol=1e-6)
        # the draw itself is checked against the kernel's own table (bit-exact argmax of p/q) ...
        tab = out["probs_table"].cpu().unsqueeze(0).expand(B, A, K)
        idx = orc.sample_indices(tab, q)
        assert torch.equal(out["idx"].cpu(), idx)
        # ... and against the oracle's table wherever the two tables do not disagree on a near-tie
        idx_o = orc.sample_indices(probs.expand(B, A, K), q)
        assert (out["idx"].cpu() != idx_o).float().mean() <= 1e-3
        actions, act_probs = orc.gather_actions(sd, tab, idx)
        assert torch.equal(out["actions"].cpu(), actions)
        assert torch.equal(out["probs"].cpu(), act_probs)
        torch.testing.assert_close(out["logp"].cpu(), torch.log(act_probs + 1e-9), rtol=0, atol=1e-6)
        assert torch.equal(out["masks"].cpu(), orc.step_masks(B, A, n_hist, od))
        coef, scale = orc.coefficients(actions, n_hist, od, sdim)
        c = out["coef"].cpu()
        if coef is not None:
            for j, cj in enumerate(coef):
                assert torch.equal(c[:, j], cj), f"coef {j} n_hist {n_hist}"
        for j in range(sdim):
            assert torch.equal(c[:, od + j], scale[j])
        for j in range(sdim, 2):
            assert torch.all(c[:, od + j] == 1)


def test_policy_forced_indices():
    sd = make_sd("sd", 256, 11, 4, 2, 0, seed=5, last_std=0.5)
    dsd = ah.sd_to_dev(sd)
    B, A, K = 33, 5, 11
    idx = torch.randint(0, K, (B, A))
    out = ah.policy(dsd, 499.0, 374.0, 999.0, 1.0, B, 4, 2, 4, idx_in=idx.cuda())
    assert torch.equal(out["idx"].cpu(), idx)
    tab = out["probs_table"].cpu().unsqueeze(0).expand(B, A, K)
    actions, act_probs = orc.gather_actions(sd, tab, idx)
    assert torch.equal(out["actions"].cpu(), actions) and torch.equal(out["probs"].cpu(), act_probs)


def _rand_coef(B, od, g, sdim):
    c = torch.randn(B, od + 2, generator=g)
    c[:, od:] = 1 + 0.05 * torch.randn(B, 2, generator=g)
    return c


def _oracle_sd(e0, cond, guidance, hist, x, c, od, scalars, vpred, sdim):
    eps = orc.cfg_combine(e0, cond, guidance) if cond is not None else e0
    n_hist = len(hist) + 1
    coef = None if n_hist == 1 else [c[:, j] for j in range(n_hist)]
    scale = [c[:, od + j] for j in range(sdim)]
    eff, xs = orc.combine_history([eps] + hist, coef, scale, x)
    sc = [torch.tensor(v, dtype=torch.float32) for v in scalars]
    return orc.ddim_update(xs, eff, sc, "v_prediction" if vpred else "epsilon"), eps


SHAPES = [(3, (4, 8, 8)), (2, (3, 5, 7)), (1, (4, 64, 64)), (5, (1, 1, 1)), (2, (1030,))]


@pytest.mark.parametrize("B,shape", SHAPES)
@pytest.mark.parametrize("n_hist", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("pair,vpred,sdim", [(True, False, 0), (False, False, 0), (True, True, 2), (False, True, 1),
                                             (True, False, 2)])
def test_step_sd_f32_bit_exact(B, shape, n_hist, pair, vpred, sdim):
    od = max(n_hist, 4)
    g = torch.Generator().manual_seed(n_hist * 100 + B)
    rn = lambda: torch.randn(B, *shape, generator=g)  # noqa: E731
    e0, cond, x = rn(), (rn() if pair else None), rn()
    hist = [rn() for _ in range(n_hist - 1)]
    c = _rand_coef(B, od, g, sdim)
    scalars = (0.8378, 0.5460, 0.9151, 0.4033)
    flags = (1 if vpred else 0) | (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0)
    ref, eps = _oracle_sd(e0, cond, 3.0, hist, x, c, od, scalars, vpred, sdim)
    out, slot = ah.step_sd(e0.cuda(), cond.cuda() if pair else None, 3.0, [h.cuda() for h in hist], x.cuda(),
                           c.cuda(), od, scalars, flags, slot=True)
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(slot.cpu(), eps)


def test_step_sd_unaligned_pointers_use_scalar_path():
    B, N, od = 3, 257, 4
    g = torch.Generator().manual_seed(0)
    big = lambda: torch.randn(B * N + 1, generator=g).cuda()[1:].view(B, N)  # 4-byte aligned only  # noqa: E731
    e0, cond, x, h1 = big(), big(), big(), big()
    c = _rand_coef(B, od, g, 0)
    scalars = (0.3, 0.95, 0.5, 0.86)
    ref, _ = _oracle_sd(e0.cpu(), cond.cpu(), 2.0, [h1.cpu()], x.cpu(), c, od, scalars, False, 0)
    out, _ = ah.step_sd(e0, cond, 2.0, [h1], x, c.cuda(), od, scalars, 0)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n_hist", [1, 4])
def test_step_sd_16bit_io(dtype, n_hist):
    """All-16-bit I/O at the C ABI.  The CFG combine rounds after every op like the caller's torch expression on
    16-bit tensors.  n_hist == 1 (the estimate is the raw 16-bit output): torch's own 16-bit evaluation of
    scheduler_ppo.py:316-330, bit for bit.  n_hist > 1 with a 16-bit latent is a layout the schedulers never request
    (torch promotion makes the latent fp32 there, see CONSOLVER_FLAG_X_F32): fp32 math, one rounding."""
    B, shape, od = 3, (4, 16, 16), 4
    g = torch.Generator().manual_seed(7)
    rn = lambda: torch.randn(B, *shape, generator=g).to(dtype)  # noqa: E731
    e0, cond, x = rn(), rn(), rn()
    hist = [rn() for _ in range(n_hist - 1)]
    c = _rand_coef(B, od, g, 0)
    scalars = (0.8378, 0.5460, 0.9151, 0.4033)
    eps = orc.cfg_combine(e0, cond, 3.0)                      # torch's 16-bit evaluation: a rounding per op
    assert eps.dtype == dtype
    if n_hist == 1:
        ref, _ = _oracle_sd(eps, None, 0.0, [], x, c, od, scalars, False, 0)        # 16-bit tensors, 0-d fp32 scalars
        assert ref.dtype == dtype
    else:
        ref, _ = _oracle_sd(eps.float(), None, 0.0, [h.float() for h in hist], x.float(), c, od, scalars, False, 0)
    out, slot = ah.step_sd(e0.cuda(), cond.cuda(), 3.0, [h.cuda() for h in hist], x.cuda(), c.cuda(), od, scalars, 0,
                           slot=True)
    assert torch.equal(slot.cpu(), eps)
    assert torch.equal(out.cpu(), ref.to(dtype))


@pytest.mark.parametrize("dtype,x_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                           (torch.bfloat16, torch.float32), (torch.float16, torch.float16)])
@pytest.mark.parametrize("n_hist,sdim", [(1, 0), (1, 2), (2, 0), (4, 1), (6, 2)])
@pytest.mark.parametrize("B,shape", [(2, (16, 8)), (3, (5, 3)), (1, (4096, 64))])
def test_step_fm_bit_exact(dtype, x_dtype, n_hist, sdim, B, shape):
    od = max(n_hist, 2)
    g = torch.Generator().manual_seed(n_hist + B)
    v = torch.randn(B, *shape, generator=g).to(dtype)
    x = torch.randn(B, *shape, generator=g).to(x_dtype)
    hist = [torch.randn(B, *shape, generator=g).to(dtype) for _ in range(n_hist - 1)]
    c = _rand_coef(B, od, g, sdim)
    dt = torch.tensor(0.8403, dtype=torch.float32) - torch.tensor(0.9045, dtype=torch.float32)
    coef = None if n_hist == 1 else [c[:, j] for j in range(n_hist)]
    scale = [c[:, od + j] for j in range(sdim)]
    eff, xs = orc.combine_history([v] + hist, coef, scale, x.to(torch.float32))
    ref = orc.fm_update(xs, eff, dt, dtype)
    flags = (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0)
    out = ah.step_fm(v.cuda(), [h.cuda() for h in hist], x.cuda(), c.cuda(), od, float(dt), flags)
    assert out.dtype == dtype
    assert torch.equal(out.cpu(), ref)


def test_launch_config_knobs_do_not_change_results():
    from consolver_b200 import _lib
    lib = _lib.load()
    B, shape, od = 4, (4, 64, 64), 4
    g = torch.Generator().manual_seed(3)
    rn = lambda: torch.randn(B, *shape, generator=g).cuda()  # noqa: E731
    e0, cond, x, h = rn(), rn(), rn(), [rn(), rn(), rn()]
    c = _rand_coef(B, od, g, 0).cuda()
    scalars = (0.8378, 0.5460, 0.9151, 0.4033)
    base, _ = ah.step_sd(e0, cond, 3.0, h, x, c, od, scalars, 0)
    try:
        for threads, unroll in itertools.product((64, 128, 512), (1, 2)):
            assert lib.consolver_set_step_launch(threads, unroll) == 0
            out, _ = ah.step_sd(e0, cond, 3.0, h, x, c, od, scalars, 0)
            assert torch.equal(out, base)
        out, _ = ah.step_sd(e0, cond, 3.0, h, x, c, od, scalars, 8)   # PDL launch attribute
        assert torch.equal(out, base)
        out, _ = ah.step_sd(e0, cond, 3.0, h, x, c, od, scalars, 16)  # CHAIN: late loads of x and the newest slot
        assert torch.equal(out, base)
    finally:
        lib.consolver_set_step_launch(0, 0)


@pytest.mark.parametrize("variant,H,K,od,sdim,mu,B", POLICY_CASES)
def test_table_and_sample_kernels_match_the_fused_policy_kernel(variant, H, K, od, sdim, mu, B):
    """The per-trajectory table launch + per-step sample launch must give exactly what the fused launch gives."""
    sd = make_sd(variant, H, K, od, sdim, mu, seed=H + K + B, last_std=0.5 if variant == "sd" else 0.02)
    dsd = ah.sd_to_dev(sd)
    A = sd["action_values"].shape[0]
    rows = torch.tensor([[999.0, 874.0], [874.0, 749.0], [124.0, -1.0]]) if variant == "sd" else \
        torch.tensor([[1.0, 0.9567], [0.9567, 0.9045], [0.3109, 0.0]])
    x_div, temp = (999.0, 1.0) if variant == "sd" else (1.0, 0.01)
    tables = ah.policy_table(dsd, rows, x_div, temp)
    ref = orc.policy_probs(sd, rows, variant)
    torch.testing.assert_close(tables.cpu(), ref, rtol=0 if variant == "sd" else 2e-5, atol=1e-6)
    g = torch.Generator().manual_seed(B + 1)
    q = torch.empty(B * A, K).exponential_(1, generator=g).cuda()
    for r, n_hist in ((1, od), (2, 1)):
        fused = ah.policy(dsd, rows[r, 0], rows[r, 1], x_div, temp, B, od, sdim, n_hist, q=q)
        assert torch.equal(fused["probs_table"], tables[r])
        split = ah.policy_sample(dsd, tables[r], B, od, sdim, n_hist, q=q)
        for k in ("idx", "actions", "probs", "logp", "masks", "coef"):
            assert torch.equal(split[k], fused[k]), k
    forced = torch.randint(0, K, (B, A)).cuda()
    a = ah.policy_sample(dsd, tables[0], B, od, sdim, 2, idx_in=forced)
    assert torch.equal(a["idx"], forced)


def test_policy_q_slab_unaligned_base():
    """q handed over at a 4-byte-aligned (not 16) address: the cp.async staging falls back to 4-byte copies."""
    sd = make_sd("sd", 64, 11, 4, 0, 0, seed=1, last_std=0.5)
    dsd = ah.sd_to_dev(sd)
    B, A, K = 37, 3, 11
    qbig = torch.empty(B * A * K + 1).exponential_(1).cuda()
    q = qbig[1:].view(B * A, K)
    out = ah.policy(dsd, 624.0, 499.0, 999.0, 1.0, B, 4, 0, 4, q=q)
    tab = out["probs_table"].cpu().unsqueeze(0).expand(B, A, K)
    assert torch.equal(out["idx"].cpu(), orc.sample_indices(tab, q.cpu()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,shape", [(3, (4, 8, 8)), (2, (3, 5, 7)), (64, (4, 64, 64)), (1, (1,))])
@pytest.mark.parametrize("n_hist,od,pair", [(1, 4, False), (2, 4, True), (4, 4, True), (3, 3, False), (6, 8, True)])
def test_cosine_features_vs_oracle(dtype, B, shape, n_hist, od, pair):
    g = torch.Generator().manual_seed(B + n_hist)
    rn = lambda: torch.randn(B, *shape, generator=g).to(dtype)  # noqa: E731
    e0, cond = rn(), (rn() if pair else None)
    hist = [rn() * (j + 1) for j in range(n_hist - 1)]
    eps = orc.cfg_combine(e0.float(), cond.float(), 2.5).to(dtype) if pair else e0
    stack = torch.stack([eps] + hist + [torch.zeros_like(eps)] * (od - n_hist), dim=1).float()
    ref = orc.cosine_features(stack, od)
    got = ah.cosine_features(e0.cuda(), cond.cuda() if pair else None, 2.5, [h.cuda() for h in hist], od)
    torch.testing.assert_close(got.cpu(), ref, rtol=0, atol=2e-6)
    assert torch.all(got[:, n_hist - 1:] == 0)


def test_policy_per_sample_features_vs_oracle():
    """use_conv: MLP input [t, t_prev, cos_1..cos_{od-1}] per sample -> per-sample tables."""
    od, K, H, B = 4, 11, 64, 9
    g = torch.Generator().manual_seed(0)
    sd = make_sd("sd", H, K, od, 0, 0, seed=3, last_std=0.5)
    sd["mlp.0.weight"] = (torch.rand(H, 2 + od - 1, generator=g) * 2 - 1) / 5 ** 0.5
    dsd = ah.sd_to_dev(sd)
    feat = torch.rand(B, od - 1, generator=g) * 2 - 1
    q = torch.empty(B * 3, K).exponential_(1, generator=g)
    out = ah.policy(dsd, 749.0, 624.0, 999.0, 1.0, B, od, 0, 4, q=q.cuda(), feat=feat.cuda().contiguous())
    xn = torch.cat([(torch.tensor([[749.0, 624.0]]) / 999.0).expand(B, 2), feat], dim=1)
    F = torch.nn.functional
    h = torch.relu(F.linear(xn, sd["mlp.0.weight"], sd["mlp.0.bias"]))
    h = torch.relu(F.linear(h, sd["mlp.2.weight"], sd["mlp.2.bias"]))
    ref = torch.softmax(F.linear(h, sd["mlp.4.weight"], sd["mlp.4.bias"]).view(B, 3, K), dim=-1)
    torch.testing.assert_close(out["probs_table"].cpu(), ref, rtol=0, atol=1e-6)
    idx = orc.sample_indices(out["probs_table"].cpu(), q)
    assert torch.equal(out["idx"].cpu(), idx)
    actions, act_probs = orc.gather_actions(sd, out["probs_table"].cpu(), idx)
    assert torch.equal(out["actions"].cpu(), actions) and torch.equal(out["probs"].cpu(), act_probs)


def test_second_destination_out2():
    """x' also lands in a second buffer (other half of the next CFG-doubled input / a wider packed sequence)."""
    B, shape, od = 3, (4, 16, 16), 4
    g = torch.Generator().manual_seed(11)
    rn = lambda: torch.randn(B, *shape, generator=g).cuda()  # noqa: E731
    e0, cond, x, h = rn(), rn(), rn(), [rn()]
    c = _rand_coef(B, od, g, 0).cuda()
    scalars = (0.8378, 0.5460, 0.9151, 0.4033)
    nxt = torch.zeros(2 * B, *shape, device="cuda")
    out, _ = ah.step_sd(e0, cond, 3.0, h, x, c, od, scalars, 0, out2=nxt[B:])
    assert torch.equal(nxt[B:], out) and torch.count_nonzero(nxt[:B]) == 0
    # FM: strided destination inside a [B, 2L, D] packed sequence
    v, xf = torch.randn(B, 64, 8, device="cuda").bfloat16(), torch.randn(B, 64, 8, device="cuda").bfloat16()
    wide = torch.zeros(B, 128, 8, device="cuda", dtype=torch.bfloat16)
    cf = _rand_coef(B, 2, g, 0).cuda()
    o = ah.step_fm(v, [], xf, cf, 2, -0.05, 0, out2=wide[:, :64])
    assert torch.equal(wide[:, :64], o) and torch.count_nonzero(wide[:, 64:]) == 0


def test_denoise_loop_without_cat_matches_manual_loop():
    import consolver_b200 as cb
    from consolver_b200.denoise import denoise_loop

    kw = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, steps_offset=1,
              timestep_spacing="trailing", order_dim=4, scaler_dim=0,
              factor_net_kwargs=dict(hidden_dim=64, num_actions=11))
    torch.manual_seed(0)
    a, b = cb.PPOScheduler(**kw), cb.PPOScheduler(**kw)
    with torch.no_grad():
        a.factor_net.mlp[4].weight.normal_(0, 0.3)
    b.factor_net.load_state_dict(a.factor_net.state_dict())
    a.factor_net.cuda(), b.factor_net.cuda()
    w = torch.randn(4, 4, device="cuda") * 0.3
    den = lambda xin, t, i: torch.einsum("oc,bchw->bohw", w * (1 + 0.1 * i), xin)  # noqa: E731
    noise = torch.randn(5, 4, 16, 16, device="cuda")
    torch.manual_seed(5)
    lat, rec = denoise_loop(a, den, noise, cfg=3.0, num_inference_steps=6)
    torch.manual_seed(5)
    b.set_timesteps(6, device="cuda")
    x = noise.clone()
    for i, t in enumerate(b.timesteps):
        pred = den(torch.cat([x] * 2), t, i)
        u, c = pred.chunk(2)
        x = b.step(u + 3.0 * (c - u), t, x, return_dict=False)[0]     # the reference's caller-side sequence
    assert torch.equal(lat, x)
    assert torch.equal(rec["idx"], b.trajectory()["idx"])


@pytest.mark.parametrize("B,A,K", [(64, 3, 11), (5, 5, 161), (1, 1, 11), (4096, 3, 11), (12000, 3, 11)])
def test_in_kernel_exponential_draw_is_torchs(B, A, K):
    """The sample kernel's own Exp(1) draw == torch.empty(B*A,K).exponential_(1) for the same generator state,
    bit for bit (incl. numel beyond one grid of the torch launch), and the generator advance matches."""
    from consolver_b200 import _lib
    sd = make_sd("sd", 64, K, A + 1, 0, 0, seed=B, last_std=0.5)
    dsd = ah.sd_to_dev(sd)
    table = torch.softmax(torch.randn(A, K), -1).cuda()
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(1234 + B)
    torch.empty(37, device="cuda").normal_()                      # move the offset off zero
    seed, off = gen.initial_seed(), gen.get_offset()
    ref = torch.empty(B * A, K, device="cuda").exponential_(1)
    nthreads, inc = _lib.philox_plan(B * A * K)
    assert gen.get_offset() - off == inc
    q_out = torch.zeros(B * A, K, device="cuda")
    out = ah.policy_sample(dsd, table, B, A + 1, 0, A + 1, rng=_lib.Rng(seed, off, None, nthreads), q_out=q_out)
    assert torch.equal(q_out, ref)
    via_q = ah.policy_sample(dsd, table, B, A + 1, 0, A + 1, q=ref)
    for k in ("idx", "actions", "probs", "coef"):
        assert torch.equal(out[k], via_q[k])
    # device-resident state: {seed, base} + per-launch offset
    state = torch.tensor([seed, off - 8], dtype=torch.int64, device="cuda")
    out2 = ah.policy_sample(dsd, table, B, A + 1, 0, A + 1, rng=_lib.Rng(0, 8, state.data_ptr(), nthreads))
    assert torch.equal(out2["idx"], out["idx"])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n_hist", [1, 2, 4])
@pytest.mark.parametrize("pair,vpred,sdim", [(False, False, 0), (True, False, 0), (False, True, 2)])
def test_step_sd_fp32_latents_with_16bit_model_outputs(dtype, n_hist, pair, vpred, sdim):
    """CONSOLVER_FLAG_X_F32 — the mixed-precision layout of the reference's training loop (train_ppo.py:353):
    16-bit model outputs / history, fp32 latent in and out.  From the second step on torch promotion makes the
    reference's arithmetic fp32 on the upcast values; at the first step (no scalers) the two products with the raw
    16-bit output are 16-bit products.  Either way: the oracle's torch expressions on these dtypes, bit for bit;
    ragged size included."""
    B, shape, od = 3, (4, 9, 7), 4
    g = torch.Generator().manual_seed(17)
    r16 = lambda: torch.randn(B, *shape, generator=g).to(dtype)  # noqa: E731
    e0, cond = r16(), (r16() if pair else None)
    hist = [r16() for _ in range(n_hist - 1)]
    x = torch.randn(B, *shape, generator=g)
    c = _rand_coef(B, od, g, sdim)
    scalars = (0.8378, 0.5460, 0.9151, 0.4033)
    eps = orc.cfg_combine(e0, cond, 3.0) if pair else e0      # the caller's 16-bit combine (a rounding per op)
    # the reference's torch expressions on exactly these dtypes: 16-bit estimate / history, fp32 latent and coefficients
    ref, _ = _oracle_sd(eps, None, 0.0, hist, x, c, od, scalars, vpred, sdim)
    assert ref.dtype == torch.float32
    flags = (1 if vpred else 0) | (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0)
    out, slot = ah.step_sd(e0.cuda(), cond.cuda() if pair else None, 3.0, [h.cuda() for h in hist], x.cuda(), c.cuda(),
                           od, scalars, flags, slot=pair)
    assert out.dtype == torch.float32
    assert torch.equal(out.cpu(), ref)
    if pair:
        assert slot.dtype == dtype and torch.equal(slot.cpu(), eps)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_step_fm_strided_model_outputs_equal_contiguous(dtype):
    """consolver_step_fm_strided: e0 and the history as `[:, :L]` views of wider tensors (sample stride (L+Li)*D) give
    exactly what contiguous copies give."""
    from consolver_b200 import _lib
    lib = _lib.load()
    B, L, Li, D, od = 3, 64, 32, 8, 2
    g = torch.Generator(device="cuda").manual_seed(3)
    for extra in (Li, 3):                                     # sample strides of 768 and 536 elements
        wide_v = torch.randn(B, L + extra, D, device="cuda", generator=g).to(dtype)
        wide_h = torch.randn(B, L + extra, D, device="cuda", generator=g).to(dtype)
        x = torch.randn(B, L, D, device="cuda", generator=g).to(dtype)
        coef = torch.tensor([[1.3, -0.3, 1.0, 1.0]], device="cuda").repeat(B, 1).contiguous()
        v, h = wide_v[:, :L], wide_h[:, :L]
        assert not v.is_contiguous()
        ref = ah.step_fm(v.contiguous(), [h.contiguous()], x, coef, od, -0.0433)
        out = torch.empty_like(x)
        rc = lib.consolver_step_fm_strided(
            _lib.dtype_code(dtype), _lib.dtype_code(dtype), v.data_ptr(), v.stride(0), None,
            _lib.ptr_array([h.data_ptr()]), 2, x.data_ptr(), out.data_ptr(), None, 0, coef.data_ptr(), od + 2, od,
            -0.0433, 0, B, L * D, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "consolver_step_fm_strided")
        assert torch.equal(out, ref)
    # a stride smaller than one sample is rejected
    rc = lib.consolver_step_fm_strided(_lib.dtype_code(dtype), _lib.dtype_code(dtype), v.data_ptr(), L * D - 8, None,
                                       _lib.ptr_array([h.data_ptr()]), 2, x.data_ptr(), out.data_ptr(), None, 0,
                                       coef.data_ptr(), od + 2, od, -0.0433, 0, B, L * D, None)
    assert rc == -2
