"""Host-side logic of the drop-in schedulers (no GPU): constructor/config surface, schedules against the golden
vectors, state_dict interchange, error behaviour, lazy conds."""
import os

import numpy as np
import pytest
import torch

import consolver_b200 as cb
from consolver_b200.config_utils import LazyConds
from golden_io import Golden, names

PROD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000,
            steps_offset=1, timestep_spacing="trailing", order_dim=4, scaler_dim=0, use_conv=False,
            factor_net_kwargs=dict(embedding_dim=64, hidden_dim=256, num_actions=11))


@pytest.mark.parametrize("name", names("sd_"))
def test_sd_schedule_and_state_dict_match_reference(name):
    g = Golden(name)
    m = g.meta
    s = cb.PPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
    ref_sd = g.state_dict
    own = s.factor_net.state_dict()
    assert list(own.keys()) == list(ref_sd.keys())
    assert all(own[k].shape == ref_sd[k].shape for k in own)
    assert torch.equal(own["action_values"], ref_sd["action_values"])      # incl. the -1.49e-08 bin
    s.factor_net.load_state_dict(ref_sd)
    s.set_timesteps(m["n"])
    assert torch.equal(s.timesteps, g["timesteps"])
    assert len(s) == 1000 and s.init_noise_sigma == 1.0 and s.order == 1
    x = torch.randn(2, 3)
    assert s.scale_model_input(x, 5) is x


@pytest.mark.parametrize("name", names("fm_"))
def test_fm_schedule_and_state_dict_match_reference(name):
    g = Golden(name)
    m = g.meta
    s = cb.FMPPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
    assert list(s.factor_net.state_dict().keys()) == list(g.state_dict.keys())
    assert torch.equal(s.factor_net.action_values, g.state_dict["action_values"])
    s.factor_net.load_state_dict(g.state_dict)
    s.set_timesteps(m["n"], sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
    assert torch.equal(s.timesteps, g["timesteps"]) and torch.equal(s.sigmas, g["sigmas"])
    assert s.step_index is None and s.begin_index is None
    s.set_begin_index(0)
    assert s.begin_index == 0


def test_sd_config_surface_and_defaults():
    s = cb.PPOScheduler(**PROD)
    assert s.config.order_dim == 4 and s.config["scaler_dim"] == 0 and s.config.get("prediction_type") == "epsilon"
    assert s.config.ppo_type == "discrete" and s.config.trained_betas is None
    d = cb.PPOScheduler()      # reference defaults (scheduler_ppo.py:82-97, :132-136)
    fn = d.factor_net
    assert (fn.hidden_dim, fn.num_actions, fn.action_dims) == (256, 161, 5)
    assert fn.mlp[0].in_features == 2 and fn.mlp[4].out_features == 5 * 161
    assert torch.count_nonzero(fn.mlp[4].weight) == 0 and torch.count_nonzero(fn.mlp[4].bias) == 0
    assert sum(p.numel() for p in s.factor_net.parameters()) == 75041     # production policy
    f = cb.FMPPOScheduler()
    assert f.factor_net.action_dims == 4 + 2 + 1 - 1 and f.config.mu_dim == 1
    assert torch.count_nonzero(f.factor_net.mlp[4].weight) > 0            # FM variant keeps default init


def test_sd_timestep_grids():
    s = cb.PPOScheduler(**PROD)
    s.set_timesteps(8)
    assert s.timesteps.tolist() == [999, 874, 749, 624, 499, 374, 249, 124]
    s.set_timesteps(7)
    assert s._stride == 142 and s.timesteps[1].item() == 856           # prev_t = t - 1000//n quirk
    lead = cb.PPOScheduler(timestep_spacing="leading", steps_offset=1)
    lead.set_timesteps(4)
    assert lead.timesteps.tolist() == [751, 501, 251, 1]
    lin = cb.PPOScheduler(timestep_spacing="linspace")
    lin.set_timesteps(3)
    assert lin.timesteps.tolist() == [999, 500, 0]


def test_error_behaviour_matches_reference():
    s = cb.PPOScheduler(**PROD)
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(ValueError, match="set_timesteps"):
        s.step(x, 999, x)
    with pytest.raises(ValueError):
        s.set_timesteps(1001)
    with pytest.raises(ValueError, match="timestep_spacing"):
        cb.PPOScheduler(timestep_spacing="bogus").set_timesteps(4)
    with pytest.raises(NotImplementedError):
        cb.PPOScheduler(beta_schedule="bogus")
    with pytest.raises(ValueError):
        cb.PPOScheduler(prediction_type="sample")
    cont = cb.PPOScheduler(ppo_type="continuous")        # extension (parity unpinned): builds FactorNetPPOContinous
    assert isinstance(cont.factor_net, cb.FactorNetPPOContinous) and cont.factor_net.action_dims == 5
    with pytest.raises(NotImplementedError):
        cb.FMPPOScheduler(ppo_type="continuous")          # edit_ppo/scheduler_fmppo.py:169-170 is `assert 0`
    s.set_timesteps(4)
    with pytest.raises(RuntimeError, match="no CPU path"):     # the product never falls back to the CPU
        s.step(x, 999, x)
    f = cb.FMPPOScheduler(use_dynamic_shifting=True)
    with pytest.raises(ValueError, match="mu"):
        f.set_timesteps(4)
    f = cb.FMPPOScheduler(order_dim=2, scaler_dim=0, mu_dim=0)
    v = torch.zeros(1, 4, 4)
    with pytest.raises(ValueError, match="set_timesteps"):
        f.step(v, 1.0, v)
    f.set_timesteps(4)
    with pytest.raises(ValueError, match="integer"):
        f.step(v, 3, v)
    with pytest.raises(ValueError, match="integer"):
        f.step(v, torch.tensor(3), v)
    with pytest.raises(RuntimeError, match="no CPU path"):
        f.step(v, f.timesteps[0], v)
    with pytest.raises(ValueError):
        cb.FMPPOScheduler(use_karras_sigmas=True, use_exponential_sigmas=True)
    with pytest.raises(ValueError):
        cb.FMPPOScheduler(time_shift_type="cubic")
    with pytest.raises(ValueError, match="same length"):
        f.set_timesteps(3, sigmas=[1.0, 0.5])


def test_fm_sigma_variants_run():
    for kw in (dict(use_karras_sigmas=True), dict(use_exponential_sigmas=True), dict(shift=3.0),
               dict(shift_terminal=0.02), dict(invert_sigmas=True), dict(time_shift_type="linear", use_dynamic_shifting=True)):
        f = cb.FMPPOScheduler(order_dim=2, scaler_dim=0, mu_dim=0, **kw)
        f.set_timesteps(6, mu=0.8 if kw.get("use_dynamic_shifting") else None)
        assert f.sigmas.shape == (7,) and f.timesteps.shape == (6,)
        assert f.index_for_timestep(f.timesteps[2]) == 2
    f = cb.FMPPOScheduler(shift=3.0, order_dim=2, scaler_dim=0, mu_dim=0)
    f.set_timesteps(8)
    f.set_begin_index(3)
    lat, noise = torch.ones(2, 4), torch.zeros(2, 4)
    out = f.scale_noise(lat, f.timesteps[:2], noise)
    torch.testing.assert_close(out, (1 - f.sigmas[3]) * lat)


def test_add_noise_matches_closed_form():
    s = cb.PPOScheduler(**PROD)
    x0, n = torch.randn(3, 4, 2, 2), torch.randn(3, 4, 2, 2)
    t = torch.tensor([0, 500, 999])
    out = s.add_noise(x0, n, t)
    a = s.alphas_cumprod[t].view(3, 1, 1, 1)
    torch.testing.assert_close(out, a.sqrt() * x0 + (1 - a).sqrt() * n)


def test_lazy_conds_materialises_on_access_only():
    calls = []
    c = LazyConds(torch.zeros(2, 2), lambda: calls.append(1) or torch.ones(2, 4, 3))
    assert c["x"].shape == (2, 2) and not calls
    assert "epsilon" in c and not calls
    assert c["epsilon"].shape == (2, 4, 3) and calls == [1]
    assert c.get("epsilon") is c["epsilon"] and calls == [1]
    assert sorted(c.keys()) == ["epsilon", "x"]


def test_fp16_cast_of_the_policy_keeps_a_loadable_state_dict():
    """gen_ppo.py:188-195 loads model.ckpt then casts the module (and its buffer) to fp16."""
    s = cb.PPOScheduler(**PROD)
    sd = {k: v.clone() for k, v in s.factor_net.state_dict().items()}
    s.factor_net.load_state_dict(sd)
    s.factor_net.to(dtype=torch.float16)
    assert s.factor_net.action_values.dtype == torch.float16


def test_config_round_trip_through_scheduler_config_json(tmp_path):
    """edit_ppo/train_ppo.py:87: FMPPOScheduler.from_pretrained(path, subfolder="scheduler", order_dim=..., ...)."""
    import json
    import os
    d = tmp_path / "flux" / "scheduler"
    os.makedirs(d)
    json.dump({"_class_name": "FlowMatchEulerDiscreteScheduler", "_diffusers_version": "0.30.0", "shift": 3.0,
               "use_dynamic_shifting": True, "base_shift": 0.5, "max_shift": 1.15, "num_train_timesteps": 1000,
               "unknown_key": 1}, open(d / "scheduler_config.json", "w"))
    if not hasattr(cb.FMPPOScheduler, "from_pretrained"):
        pytest.skip("diffusers mixin without local loader")
    f = cb.FMPPOScheduler.from_pretrained(str(tmp_path / "flux"), subfolder="scheduler", order_dim=2, scaler_dim=0, mu_dim=0,
                                          factor_net_kwargs=dict(hidden_dim=256, num_actions=11))
    assert f.config.shift == 3.0 and f.config.use_dynamic_shifting and f.config.order_dim == 2
    assert f.factor_net.action_dims == 1
    f.save_pretrained(str(tmp_path / "out"))
    again = cb.FMPPOScheduler.from_pretrained(str(tmp_path / "out"))
    assert dict(again.config) == dict(f.config)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU / PyTorch fallback: if libconsolver.so is absent and cannot be built, loading raises."""
    from consolver_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setenv("CONSOLVER_NO_AUTOBUILD", "1")
    with pytest.raises(_lib.ConsolverError, match="not found"):
        _lib.load()


@pytest.mark.parametrize("name", names("fmgen_"))
def test_fm_baseline_scheduler_schedule_matches_reference(name):
    """FlowMatchGeneralDiscreteScheduler (edit_ppo/scheduler_fm.py): constructor surface and sigma schedule."""
    g = Golden(name)
    m = g.meta
    s = cb.FlowMatchGeneralDiscreteScheduler(**m["config"])
    assert s.config.type == m["solver"] == s.type and s.order == 1 and len(s) == 1000
    if m["config"]["use_dynamic_shifting"]:
        s.set_timesteps(m["n"], sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
    else:
        s.set_timesteps(m["n"])
    assert torch.equal(s.timesteps, g["timesteps"]) and torch.equal(s.sigmas, g["sigmas"])
    assert s.step_index is None and s.begin_index is None
    with pytest.raises(RuntimeError, match="no CPU path"):
        s.step(g["v_0"], s.timesteps[0], g["x_T"])


def test_fm_baseline_scheduler_errors():
    with pytest.raises(ValueError):
        cb.FlowMatchGeneralDiscreteScheduler(use_karras_sigmas=True, use_exponential_sigmas=True)
    with pytest.raises(ValueError):
        cb.FlowMatchGeneralDiscreteScheduler(time_shift_type="cubic")
    s = cb.FlowMatchGeneralDiscreteScheduler(use_dynamic_shifting=True)
    with pytest.raises(ValueError, match="mu"):
        s.set_timesteps(4)
    d = cb.FlowMatchGeneralDiscreteScheduler()                # reference defaults (edit_ppo/scheduler_fm.py:92-109)
    assert d.config.type == "euler" and d.config.shift == 1.0 and d.shift == 1.0
    assert d.sigma_max == 1.0 and abs(d.sigma_min - 1e-3) < 1e-9


# ---- DPMSolverMultistepScheduler with AMED scaling (diffusers_amed_plugin_dpmpp.py) -----------------------------
def _amed(g):
    m = g.meta
    s = cb.DPMSolverMultistepScheduler(**m["config"])
    if m["amed"]:
        s.scale_dirs, s.scale_times = m["scale_dirs"], m["scale_times"]
        s.set_timesteps(m["n"], timesteps=m["schedule"])
    else:
        s.set_timesteps(m["n"])
    return s


def _dpm_kernel_arithmetic(p, order, e, x, hist):
    """include/consolver.h's statement of consolver_step_dpm, in torch-CPU fp32 ops (same roundings).
    `hist`: converted outputs of the earlier steps, newest first.  Returns (x', [m0] + hist)."""
    from consolver_b200 import _lib
    if p.convert == _lib.DPM_CONVERT_DIV:
        m0 = (x - p.ck0 * e) / p.ck1
    elif p.convert == _lib.DPM_CONVERT_LIN:
        m0 = p.ck1 * x + p.ck0 * e
    else:
        m0 = e
    out = p.cx * x - p.a0 * m0
    if order == 2:
        out = out - p.a1 * (p.rinv * (m0 - hist[0]))
    elif order == 3:
        a1, rinv, a2, rinv1, w, rs = p.third
        d10, d11 = rinv * (m0 - hist[0]), rinv1 * (hist[0] - hist[1])
        dd = d10 - d11
        out = (out - a1 * (d10 + w * dd)) - a2 * (rs * dd)
    return out, ([m0] + hist)[:2]


@pytest.mark.parametrize("name", names("amed_"))
def test_amed_scheduler_grid_and_host_scalars_reproduce_the_plugin(name):
    """The host half of the AMED / DPM-Solver scheduler without a GPU: grids equal the plugin's, and the per-step
    scalars + order selection, pushed through the kernel's documented arithmetic, give the plugin's latents
    bit for bit (the GPU tests then only have to show that the kernel does that arithmetic)."""
    g = Golden(name)
    s = _amed(g)
    assert torch.equal(s.timesteps, g["timesteps"]) and torch.equal(s.sigmas, g["sigmas"])
    x, hist = g["x_T"], []
    for i, t in enumerate(s.timesteps):
        idx, plan, order = s._plan_for_step(t)
        assert idx == i
        x, hist = _dpm_kernel_arithmetic(plan, order, g[f"eps_{i}"], x, hist)
        s._advance()
        assert torch.equal(x, g[f"prev_{i}"]), f"step {i}"
    with pytest.raises(RuntimeError, match="no CPU path"):
        s.step(g["eps_0"], 0, x)


def test_amed_scheduler_surface_and_errors():
    D = cb.DPMSolverMultistepScheduler
    s = D(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1)
    assert s.config.solver_order == 2 and s.config.algorithm_type == "dpmsolver++" and s.order == 1
    assert s.init_noise_sigma == 1.0 and len(s) == 1000
    with pytest.raises(ValueError, match="set_timesteps"):
        s._plan_for_step(999)
    with pytest.raises(AssertionError):                      # plugin :48-49
        s.set_timesteps(4, timesteps=[999, 694, 500, 110, 0])
    for bad in (dict(algorithm_type="sde-dpmsolver++"), dict(solver_order=4), dict(thresholding=True),
                dict(solver_type="bh1"), dict(beta_schedule="squaredcos_cap_v2")):
        with pytest.raises(NotImplementedError):
            D(**bad)
    with pytest.raises(ValueError):
        D(prediction_type="flow")
    # SD1.5's scheduler_config.json carries PNDM-only keys; from_config must ignore them (gen_ppo.py:160-163)
    cfg = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
               set_alpha_to_one=False, skip_prk_steps=True, steps_offset=1, clip_sample=False, trained_betas=None)
    s2 = D.from_config(cfg)
    s2.scale_dirs, s2.scale_times = [1.0, 0.991, 1.0, 0.9912, 1.0], [1.0, 1.0333, 1.0, 0.9861, 1.0]
    s2.set_timesteps(4, timesteps=[999, 694, 500, 110, 0])
    assert s2.timesteps.tolist()[0] == 999 and s2.timesteps.tolist()[2] == 500 and len(s2.timesteps) == 4
    assert s2.num_inference_steps == 5                        # plugin :60 counts the trailing 0


# ---- round-2 host logic --------------------------------------------------------------------------------------------------
def test_lazy_conds_is_a_complete_mapping_and_guards_stale_ring_reads():
    from consolver_b200.config_utils import LazyConds

    calls = []
    x = torch.zeros(2, 2)

    def thunk():
        calls.append(1)
        return torch.ones(2, 4, 3)

    c = LazyConds(x, thunk)
    assert len(c) == 2 and "epsilon" in c and not calls            # nothing materialised by len / contains
    assert set(iter(c)) == {"x", "epsilon"} and len(calls) == 1     # iteration materialises (dict(conds), {**conds})
    assert dict(c)["epsilon"].shape == (2, 4, 3) and len(calls) == 1
    c2 = LazyConds(x, thunk)
    assert set({**c2}) == {"x", "epsilon"} and c2.copy()["epsilon"].sum() == 24
    alive = [True]
    stale = LazyConds(x, thunk, still_valid=lambda: alive[0])
    alive[0] = False
    with pytest.raises(RuntimeError, match="history ring had been overwritten"):
        stale["epsilon"]
    with pytest.raises(RuntimeError, match="history ring had been overwritten"):
        dict(stale)
    ok = LazyConds(x, thunk, still_valid=lambda: True)
    assert ok.get("epsilon").shape == (2, 4, 3)


def test_reference_semantics_flags_follow_the_policy_dtype_and_the_device_setting(monkeypatch):
    from consolver_b200 import _lib, _sched_common

    s = cb.PPOScheduler(**PROD)
    assert s.reference_device == "cuda"                                               # product default
    assert s._semantics(torch.float32) == (0, 0, None)
    s.reference_device = "cpu"
    assert s._semantics(torch.float32) == (_lib.FLAG_HOST_SCALARS, _lib.POLICY_HOST_DIV, None)
    s.reference_device = "tpu"
    with pytest.raises(ValueError):
        s._semantics(torch.float32)
    s.reference_device = "cuda"
    s.factor_net.to(torch.float16)                                                    # gen_ppo.py:194-195: bins too
    sf, pf, act = s._semantics(torch.float16)
    assert sf == _lib.FLAG_LOWP_COEF and pf == _lib.POLICY_ACT_F16 | _lib.POLICY_COEF_F16 and act == torch.float16
    assert s._semantics(torch.float32)[0] == 0                 # fp32 model outputs: the 16-bit coefficients just promote
    assert s._estimate_stays_lowp(torch.float16, 1) and not s._estimate_stays_lowp(torch.float16, 2)
    monkeypatch.setattr(_sched_common, "DEFAULT_REFERENCE_DEVICE", "cpu")
    assert cb.PPOScheduler(**PROD).reference_device == "cpu"
    assert cb.DPMSolverMultistepScheduler().reference_device == "cpu"
    f = cb.FMPPOScheduler(order_dim=2, scaler_dim=0, mu_dim=0)
    assert f._semantics(torch.bfloat16)[1] == _lib.POLICY_HOST_DIV


def test_stage_ref_copies_byte_identical_files_with_a_manifest(tmp_path):
    import stage_ref

    if not os.path.isdir(stage_ref.SOURCE):
        pytest.skip("reference tree not present")
    dest = tmp_path / "_ref"
    assert stage_ref.stage(dest=str(dest), quiet=True)
    assert stage_ref.verify(str(dest))
    for rel in stage_ref.FILES:
        assert open(os.path.join(stage_ref.SOURCE, rel), "rb").read() == open(dest / rel, "rb").read()
    (dest / "scheduler_ppo.py").write_text("# edited\n")
    assert not stage_ref.verify(str(dest))                     # an edited copy no longer matches its recorded hash
    assert not stage_ref.stage(source=str(tmp_path / "nowhere"), dest=str(dest), quiet=True)


def test_bench_arms_print_the_same_config():
    import importlib

    bench = importlib.import_module("bench")
    a, b = bench.workload_config(64, 8), bench.workload_config(64, 8)
    assert a == b and a["workload"] == bench.WORKLOAD and "l2_policy" in a and a["batch_per_gpu"] == 64
