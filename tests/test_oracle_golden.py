"""Pins the CPU oracle (oracle/consolver_oracle.py) against the reference's own outputs (tests/golden)."""
import pytest
import torch

import consolver_oracle as orc
from golden_io import Golden, names

SD = names("sd_")
FM = names("fm_")


def _sd_sched(g):
    return orc.OracleSDScheduler(g.state_dict, **g.meta["config"])


@pytest.mark.parametrize("name", SD)
def test_sd_oracle_matches_reference(name):
    g = Golden(name)
    m = g.meta
    s = _sd_sched(g)
    s.set_timesteps(m["n"])
    assert torch.equal(s.timesteps, g["timesteps"])
    x = g["x_T"]
    for i, t in enumerate(s.timesteps):
        u, c = g[f"pair_{i}"].chunk(2)
        eps = orc.cfg_combine(u, c, m["guidance"])
        assert torch.equal(eps, g[f"eps_{i}"])
        x, actions, probs, conds, masks = s.step(eps, t, x, q=g[f"q_{i}"])
        assert torch.equal(s.last_idx, g[f"idx_{i}"]), f"step {i} indices"
        assert torch.equal(actions, g[f"actions_{i}"])
        assert torch.equal(masks, g[f"masks_{i}"])
        assert torch.equal(conds["x"], g[f"condx_{i}"])
        torch.testing.assert_close(s.last_probs_full, g[f"probs_full_{i}"], rtol=0, atol=1e-7)
        torch.testing.assert_close(probs, g[f"probs_{i}"], rtol=0, atol=1e-7)
        assert torch.equal(x, g[f"prev_{i}"]), f"step {i} latent not bit-exact"


@pytest.mark.parametrize("name", FM)
def test_fm_oracle_matches_reference(name):
    g = Golden(name)
    m = g.meta
    cfg = dict(m["config"])
    for k in ("base_shift", "max_shift", "base_image_seq_len", "max_image_seq_len"):
        cfg.pop(k, None)
    s = orc.OracleFMScheduler(g.state_dict, **cfg)
    import numpy as np
    s.set_timesteps(m["n"], sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
    if m["use_begin_index"]:
        s.set_begin_index(0)
    assert torch.equal(s.timesteps, g["timesteps"])
    assert torch.equal(s.sigmas, g["sigmas"])
    x = g["x_T"]
    for i, t in enumerate(s.timesteps):
        x, actions, probs, conds, masks = s.step(g[f"v_{i}"], t, x, q=g[f"q_{i}"])
        assert torch.equal(s.last_idx, g[f"idx_{i}"]), f"step {i} indices"
        assert torch.equal(actions, g[f"actions_{i}"])
        assert torch.equal(masks, g[f"masks_{i}"])
        assert torch.equal(conds["x"], g[f"condx_{i}"])
        torch.testing.assert_close(s.last_probs_full, g[f"probs_full_{i}"], rtol=1e-6, atol=1e-7)
        assert x.dtype == g[f"prev_{i}"].dtype
        assert torch.equal(x, g[f"prev_{i}"]), f"step {i} latent not bit-exact"


# ---- fixtures made by the reference running ON A B200 (oracle/make_golden.py cuda): ATen's CUDA scalar rules, the
# ---- shipped fp16-autocast inference flow (gen_ppo.py:193-195,:309), the autocast training rollouts ---------------------
_DT = {None: None, "float16": torch.float16, "bfloat16": torch.bfloat16}
# measured on the 23 fixtures (profiles/parity_spread_r02.md): the oracle's torch-CPU MLP vs the reference's cuBLAS MLP
CUDA_PROB_ATOL, CUDA_PROB_RTOL = 4e-6, 1e-4


def cuda_oracle(g):
    """the oracle configured the way fixture `g` was produced (device / autocast / policy dtype from its meta)"""
    import numpy as np
    m = g.meta
    sem = orc.TorchSemantics(m.get("device", "cpu"), _DT[m.get("autocast")])
    if m["kind"] == "sd":
        s = orc.OracleSDScheduler(g.state_dict, sem=sem, policy_dtype=_DT[m.get("policy_dtype")], **m["config"])
        s.set_timesteps(m["n"])
        return s
    cfg = dict(m["config"])
    for k in ("base_shift", "max_shift", "base_image_seq_len", "max_image_seq_len"):
        cfg.pop(k, None)
    s = orc.OracleFMScheduler(g.state_dict, sem=sem, **cfg)
    s.set_timesteps(m["n"], sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
    if m["use_begin_index"]:
        s.set_begin_index(0)
    return s


@pytest.mark.parametrize("name", names("cuda_"))
def test_oracle_with_cuda_semantics_matches_the_reference_run_on_a_gpu(name):
    """Every latent (value AND dtype), the CFG-combined outputs, actions, masks and condition rows bit-identical; the
    oracle's own categorical draw (its MLP + the fixture's Exp(1) values) picks the reference's indices; softmax
    tables within the measured MKL-vs-cuBLAS spread."""
    g = Golden(name)
    m = g.meta
    assert m["device"] == "cuda"
    s = cuda_oracle(g)
    assert torch.equal(s.timesteps, g["timesteps"])
    x = g["x_T"]
    for i, t in enumerate(s.timesteps):
        if m["kind"] == "sd":
            u, c = g[f"pair_{i}"].chunk(2)
            mo = orc.cfg_combine(u, c, m["guidance"])
            assert torch.equal(mo, g[f"eps_{i}"])
        else:
            mo = g[f"v_{i}"]
        x, actions, probs, conds, masks = s.step(mo, t, x, q=g[f"q_{i}"])
        assert torch.equal(s.last_idx, g[f"idx_{i}"]), f"step {i} indices"
        assert actions.dtype == g[f"actions_{i}"].dtype and torch.equal(actions, g[f"actions_{i}"])
        assert torch.equal(masks, g[f"masks_{i}"])
        assert torch.equal(conds["x"], g[f"condx_{i}"])
        torch.testing.assert_close(s.last_probs_full, g[f"probs_full_{i}"], rtol=CUDA_PROB_RTOL, atol=CUDA_PROB_ATOL)
        ref = g[f"prev_{i}"]
        assert x.dtype == ref.dtype, f"step {i}: latent dtype {x.dtype} vs {ref.dtype}"
        assert torch.equal(x, ref), f"step {i} latent not bit-exact"


def test_host_and_cuda_semantics_differ_where_aten_differs():
    """the two rule sets are not interchangeable: true division vs reciprocal multiply (fp32), scalar rounding (fp16)"""
    torch.manual_seed(0)
    x, e = torch.randn(4096), torch.randn(4096)
    sc = orc.ddim_scalars(orc.sd_alphas_cumprod(orc.sd_betas(beta_schedule="scaled_linear", beta_start=0.00085,
                                                             beta_end=0.012)), 624, 499)
    a, b = orc.ddim_update(x, e, sc, sem=orc.HOST), orc.ddim_update(x, e, sc, sem=orc.CUDA)
    assert not torch.equal(a, b) and (a - b).abs().max() <= 4e-6
    a, b = orc.ddim_update(x.half(), e.half(), sc, sem=orc.HOST), orc.ddim_update(x.half(), e.half(), sc, sem=orc.CUDA)
    assert a.dtype == b.dtype == torch.float16 and not torch.equal(a, b)


def _fmgen_set_timesteps(s, m):
    import numpy as np
    if m["config"]["use_dynamic_shifting"]:
        s.set_timesteps(m["n"], sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
    else:
        s.set_timesteps(m["n"])
    if m["use_begin_index"]:
        s.set_begin_index(0)


@pytest.mark.parametrize("name", names("fmgen_"))
def test_fm_baseline_solvers_oracle_matches_reference(name):
    """euler / heun / dpm-solver / dpm-solver-multistep of edit_ppo/scheduler_fm.py:384-488, bit-exact."""
    g = Golden(name)
    m = g.meta
    cfg = dict(m["config"])
    s = orc.OracleFMGeneralScheduler(kind=cfg.pop("type"), **cfg)
    _fmgen_set_timesteps(s, m)
    assert torch.equal(s.timesteps, g["timesteps"])
    assert torch.equal(s.sigmas, g["sigmas"])
    x = g["x_T"]
    for i, t in enumerate(s.timesteps):
        x = s.step(g[f"v_{i}"], t, x)
        assert x.dtype == g[f"prev_{i}"].dtype
        assert torch.equal(x, g[f"prev_{i}"]), f"step {i} latent not bit-exact"


@pytest.mark.parametrize("name", names("amed_"))
def test_amed_dpm_solver_oracle_matches_plugin(name):
    """diffusers_amed_plugin_dpmpp.py run unmodified (over a stand-in of its diffusers base, see
    oracle/ref_shim.py::_dpm_base) vs the oracle restatement: grids and every latent bit-exact."""
    g = Golden(name)
    m = g.meta
    s = orc.OracleDPMSolverAMED(scale_dirs=m["scale_dirs"], scale_times=m["scale_times"], **m["config"])
    s.set_timesteps(m["n"], timesteps=m["schedule"])
    assert torch.equal(s.timesteps, g["timesteps"])
    assert torch.equal(s.sigmas, g["sigmas"])
    x = g["x_T"]
    for i, t in enumerate(s.timesteps):
        x = s.step(g[f"eps_{i}"], t, x)
        assert torch.equal(x, g[f"prev_{i}"]), f"step {i} latent not bit-exact"


@pytest.mark.parametrize("name", names("update_"))
def test_update_side_oracle(name):
    g = Golden(name)
    p, e = orc.action_probs_entropy(g.state_dict, g["x"], g["actions"], g.meta["variant"])
    torch.testing.assert_close(p, g["probs"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(e, g["entropy"], rtol=1e-5, atol=1e-6)


def test_action_value_tables_match_reference_buffers():
    for name in SD + FM:
        g = Golden(name)
        c, f = g.meta["config"], g.meta["factor_net_kwargs"]
        tab = orc.action_value_table(g.meta["kind"], f["num_actions"], c["order_dim"], c["scaler_dim"],
                                     c.get("mu_dim", 0))
        assert torch.equal(tab, g.state_dict["action_values"]), name


@pytest.mark.parametrize("name", names("sd16_"))
def test_sd_oracle_matches_reference_on_16bit_model_outputs(name):
    """fp16 / bf16 denoiser outputs with 16-bit or fp32 latents: the oracle is the reference's torch ops, so it must
    follow torch's promotion (latent dtype trajectory) and 16-bit scalar-product roundings exactly."""
    g = Golden(name)
    m = g.meta
    s = orc.OracleSDScheduler(g.state_dict, **m["config"])
    s.set_timesteps(m["n"])
    x = g["x_T"]
    for i, t in enumerate(s.timesteps):
        u, c = g[f"pair_{i}"].chunk(2)
        eps = orc.cfg_combine(u, c, m["guidance"])
        assert torch.equal(eps, g[f"eps_{i}"])
        x, actions, probs, conds, masks = s.step(eps, t, x, q=g[f"q_{i}"])
        assert torch.equal(s.last_idx, g[f"idx_{i}"]) and torch.equal(actions, g[f"actions_{i}"])
        assert torch.equal(conds["x"], g[f"condx_{i}"])
        assert x.dtype == g[f"prev_{i}"].dtype and torch.equal(x, g[f"prev_{i}"]), f"step {i}"
