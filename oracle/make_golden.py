"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python oracle/make_golden.py            # regenerates every fixture

Each fixture holds the inputs and the reference's own outputs for one scheduler configuration:
weights (`sd.*`), initial latent, per-step model outputs, the Exp(1) draw `q` that
`torch.multinomial` consumed (recovered by re-seeding the default generator with the same seed;
the script asserts argmax(p/q) reproduces the reference's indices), sampled actions / probs / masks,
the full softmax table and the next latent.  bf16 tensors are stored as their uint16 bit patterns
(`__bf16__` lists the keys).  The tests never read /root/reference; they read these files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.environ.get("CONSOLVER_GOLDEN_OUT") or os.path.join(os.path.dirname(HERE), "tests", "golden")

SD_PROD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000,
               steps_offset=1, timestep_spacing="trailing", order_dim=4, scaler_dim=0, use_conv=False)
FM_PROD = dict(shift=3.0, use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15, base_image_seq_len=256,
               max_image_seq_len=4096, order_dim=2, scaler_dim=0, mu_dim=0)


def _np(t: torch.Tensor, bf16_keys, key):
    t = t.detach().cpu()
    if t.dtype == torch.bfloat16:
        bf16_keys.append(key)
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def _seed_policy(fn, seed, last_std):
    """Random-init stand-in for the unreachable published checkpoint: default nn.Linear init for
    layers 0/2 under `seed`, last layer N(0, last_std^2) weight, small bias (the reference zero-inits
    the SD last layer -> uniform policy, which would leave the sampling path untested)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in fn.named_parameters():
            if name.startswith("mlp.4"):
                p.copy_(torch.randn(p.shape, generator=g) * (last_std if name.endswith("weight") else 0.1))
            else:
                bound = 1.0 / (p.shape[-1] if p.dim() > 1 else fn.mlp[int(name.split(".")[1])].in_features) ** 0.5
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)


def _hook_probs(fn, store):
    orig = fn.forward_

    def wrapped(x_dict):
        p = orig(x_dict)
        store.append(p.detach().clone())
        return p

    fn.forward_ = wrapped


def _autocast(device, dtype):
    import contextlib

    return torch.autocast(torch.device(device).type, dtype) if dtype is not None else contextlib.nullcontext()


def gen_sd(name, *, B, shape, n, guidance, seed, last_std=0.5, dtype=torch.float32, latent_dtype=None, device="cpu",
           policy_dtype=None, autocast=None, **cfg_over):
    """`dtype`: dtype of the denoiser output (and of the caller-side CFG combine); `latent_dtype`: dtype of the
    initial latent (default = dtype).  16-bit `dtype` with fp32 latents is the autocast layout of train_ppo.py:353;
    16-bit both is gen_ppo.py's fp16 pipeline, whose latents torch promotion turns fp32 after the second step.
    `device="cuda"`: the reference runs on the GPU exactly as its drivers run it — policy moved with `.to(device)`,
    `set_timesteps(device=...)`, CUDA 0-d timesteps, schedule tables left on the host (scheduler_ppo.py:110-114).
    `policy_dtype`: gen_ppo.py:193-195 casts the policy, bin buffer included, to the pipeline dtype.
    `autocast`: dtype of the `torch.autocast` region the CFG combine and `step` run in (gen_ppo.py:309,
    train_ppo.py:353)."""
    ref = ref_shim.load_reference()
    cfg = dict(SD_PROD, **cfg_over)
    fkw = dict(embedding_dim=64, hidden_dim=cfg.pop("hidden_dim", 256), num_actions=cfg.pop("num_actions", 11))
    with ref_shim.quiet():
        s = ref.PPOScheduler(factor_net_kwargs=dict(fkw), **cfg)
    _seed_policy(s.factor_net, seed, last_std)
    bf, d = [], {}
    for k, v in s.factor_net.state_dict().items():
        d[f"sd.{k}"] = v.numpy().copy()                    # fp32 master weights, before any cast
    if policy_dtype is not None:
        s.factor_net.to(device, dtype=policy_dtype)        # gen_ppo.py:194-195
    else:
        s.factor_net.to(device)
    full = []
    _hook_probs(s.factor_net, full)
    s.set_timesteps(n, device=device)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, *shape, generator=g).to(latent_dtype or dtype).to(device)
    d["x_T"] = _np(x, bf, "x_T")
    d["timesteps"] = s.timesteps.cpu().numpy()
    A, K = s.factor_net.action_dims, s.factor_net.num_actions
    for i, t in enumerate(s.timesteps):
        pair = torch.randn(2 * B, *shape, generator=g).to(dtype).to(device)
        u, c = pair.chunk(2)
        with _autocast(device, autocast):
            eps = u + guidance * (c - u)                      # denoise_ppo.py:97-100, gen_pretrain/pipeline.py:1069-1071
            torch.manual_seed(seed * 1000 + i)
            with ref_shim.quiet():
                x, actions, probs, conds, masks = s.step(eps, t, x, return_dict=False)
        torch.manual_seed(seed * 1000 + i)
        q = torch.empty(B * A, K, device=device).exponential_(1)
        idx = torch.argmax(full[i].view(-1, K) / q, dim=-1).view(B, A)
        assert torch.equal(s.factor_net.action_values[torch.arange(A, device=device), idx], actions), "q recovery failed"
        d[f"pair_{i}"] = _np(pair, bf, f"pair_{i}")
        d[f"eps_{i}"] = _np(eps, bf, f"eps_{i}")
        d[f"q_{i}"] = q.cpu().numpy()
        d[f"idx_{i}"] = idx.cpu().numpy()
        d[f"actions_{i}"] = _np(actions, bf, f"actions_{i}")
        d[f"probs_{i}"] = _np(probs, bf, f"probs_{i}")
        d[f"probs_full_{i}"] = _np(full[i], bf, f"probs_full_{i}")
        d[f"masks_{i}"] = _np(masks, bf, f"masks_{i}")
        d[f"condx_{i}"] = _np(conds["x"], bf, f"condx_{i}")
        d[f"prev_{i}"] = _np(x, bf, f"prev_{i}")
        assert conds["epsilon"].shape == (B, cfg["order_dim"], *shape)
        assert torch.equal(conds["epsilon"][:, 0], eps)
    short = lambda t: None if t is None else str(t).split(".")[-1]  # noqa: E731
    meta = dict(kind="sd", B=B, shape=list(shape), n=n, guidance=guidance, seed=seed, config=cfg,
                factor_net_kwargs=fkw, dtype=str(dtype).split(".")[-1],
                latent_dtype=str(latent_dtype or dtype).split(".")[-1], torch=torch.__version__,
                device=torch.device(device).type, policy_dtype=short(policy_dtype), autocast=short(autocast),
                gpu=torch.cuda.get_device_name(0) if torch.device(device).type == "cuda" else None)
    d["__meta__"] = np.array(json.dumps(meta))
    d["__bf16__"] = np.array(json.dumps(bf))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, "A=%d K=%d" % (A, K))


def main_sd16():
    """`python oracle/make_golden.py sd16`: 16-bit denoiser outputs (prefix sd16_ keeps them out of the fp32 sweeps)."""
    small = (4, 8, 8)
    f16, b16, f32 = torch.float16, torch.bfloat16, torch.float32
    gen_sd("sd16_f16_pipeline_eps_s0_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=90, dtype=f16)
    gen_sd("sd16_bf16_pipeline_v_s0_n6_B2", hidden_dim=64, B=2, shape=small, n=6, guidance=3.0, seed=91, dtype=b16,
           prediction_type="v_prediction")
    gen_sd("sd16_f16_autocast_eps_s0_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=92, dtype=f16,
           latent_dtype=f32)                                                               # train_ppo.py:353
    gen_sd("sd16_bf16_autocast_v_s0_n5_B2_ragged", hidden_dim=64, B=2, shape=(3, 5, 7), n=5, guidance=3.0, seed=93,
           dtype=b16, latent_dtype=f32, prediction_type="v_prediction")
    gen_sd("sd16_f16_pipeline_eps_s2_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=3.0, seed=94, dtype=f16,
           scaler_dim=2)
    gen_sd("sd16_f16_pipeline_v_s1_n4_B2", hidden_dim=64, B=2, shape=small, n=4, guidance=3.0, seed=96, dtype=f16,
           scaler_dim=1, prediction_type="v_prediction")
    gen_sd("sd16_bf16_autocast_eps_s1_o3_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=3.0, seed=95, dtype=b16,
           latent_dtype=f32, scaler_dim=1, order_dim=3)


def gen_fm(name, *, B, shape, n, seed, dtype, last_std=0.02, use_begin_index=True, device="cpu", autocast=None,
           **cfg_over):
    """`device` / `autocast`: as in gen_sd (edit_ppo/generate_ours.py:135-141 runs the fp32 policy on the GPU with bf16
    latents and no autocast; the training rollout runs under accelerator.autocast, edit_ppo/train_ppo.py:289)."""
    ref = ref_shim.load_reference()
    cfg = dict(FM_PROD, **cfg_over)
    fkw = dict(hidden_dim=cfg.pop("hidden_dim", 256), num_actions=cfg.pop("num_actions", 11))
    with ref_shim.quiet():
        s = ref.FMPPOScheduler(factor_net_kwargs=dict(fkw), **cfg)
    _seed_policy(s.factor_net, seed, last_std)
    bf, d = [], {}
    for k, v in s.factor_net.state_dict().items():
        d[f"sd.{k}"] = v.numpy().copy()
    s.factor_net.to(device)
    full = []
    _hook_probs(s.factor_net, full)
    s.set_timesteps(n, device=device, sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
    if use_begin_index:
        s.set_begin_index(0)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, *shape, generator=g).to(dtype).to(device)
    d["x_T"] = _np(x, bf, "x_T")
    d["timesteps"] = s.timesteps.cpu().numpy()
    d["sigmas"] = s.sigmas.cpu().numpy()
    A, K = s.factor_net.action_dims, s.factor_net.num_actions
    for i, t in enumerate(s.timesteps):
        v = torch.randn(B, *shape, generator=g).to(dtype).to(device)
        torch.manual_seed(seed * 1000 + i)
        with ref_shim.quiet(), _autocast(device, autocast):
            x, actions, probs, conds, masks = s.step(v, t, x, return_dict=False)
        torch.manual_seed(seed * 1000 + i)
        q = torch.empty(B * A, K, device=device).exponential_(1)
        idx = torch.argmax(full[i].view(-1, K) / q, dim=-1).view(B, A)
        assert torch.equal(s.factor_net.action_values[torch.arange(A, device=device), idx], actions), "q recovery failed"
        d[f"v_{i}"] = _np(v, bf, f"v_{i}")
        d[f"q_{i}"] = q.cpu().numpy()
        d[f"idx_{i}"] = idx.cpu().numpy()
        d[f"actions_{i}"] = _np(actions, bf, f"actions_{i}")
        d[f"probs_{i}"] = _np(probs, bf, f"probs_{i}")
        d[f"probs_full_{i}"] = _np(full[i], bf, f"probs_full_{i}")
        d[f"masks_{i}"] = _np(masks, bf, f"masks_{i}")
        d[f"condx_{i}"] = _np(conds["x"], bf, f"condx_{i}")
        d[f"prev_{i}"] = _np(x, bf, f"prev_{i}")
    meta = dict(kind="fm", B=B, shape=list(shape), n=n, seed=seed, config=cfg, factor_net_kwargs=fkw,
                dtype=str(dtype).split(".")[-1], mu=1.15, use_begin_index=use_begin_index,
                torch=torch.__version__, device=torch.device(device).type,
                autocast=None if autocast is None else str(autocast).split(".")[-1],
                gpu=torch.cuda.get_device_name(0) if torch.device(device).type == "cuda" else None)
    d["__meta__"] = np.array(json.dumps(meta))
    d["__bf16__"] = np.array(json.dumps(bf))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, "A=%d K=%d" % (A, K))


def gen_fm_general(name, *, kind, B, shape, n, seed, dtype, use_begin_index=True, **cfg_over):
    """Baseline flow-matching solvers (edit_ppo/scheduler_fm.py:384-488): no policy, no RNG."""
    ref = ref_shim.load_reference()
    cfg = dict(dict(shift=3.0, use_dynamic_shifting=True, type=kind), **cfg_over)
    s = ref.FlowMatchGeneralDiscreteScheduler(**cfg)
    if cfg["use_dynamic_shifting"]:
        s.set_timesteps(n, sigmas=np.linspace(1.0, 1 / n, n), mu=1.15)
    else:
        s.set_timesteps(n)
    if use_begin_index:
        s.set_begin_index(0)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, *shape, generator=g).to(dtype)
    bf, d = [], {}
    d["x_T"] = _np(x, bf, "x_T")
    d["timesteps"] = s.timesteps.numpy()
    d["sigmas"] = s.sigmas.numpy()
    for i, t in enumerate(s.timesteps):
        v = torch.randn(B, *shape, generator=g).to(dtype)
        x = s.step(v, t, x, return_dict=False)[0]
        d[f"v_{i}"] = _np(v, bf, f"v_{i}")
        d[f"prev_{i}"] = _np(x, bf, f"prev_{i}")
    meta = dict(kind="fm_general", solver=kind, B=B, shape=list(shape), n=n, seed=seed, config=cfg,
                dtype=str(dtype).split(".")[-1], mu=1.15, use_begin_index=use_begin_index, torch=torch.__version__)
    d["__meta__"] = np.array(json.dumps(meta))
    d["__bf16__"] = np.array(json.dumps(bf))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name)


# gen_ppo.py:24-55: the AMED schedules, gradient ("direction") scales and time scales the reference ships
AMED_SCHEDULES = {
    4: ([999, 694, 500, 110, 0], [1.0, 0.991, 1.0, 0.9912, 1.0], [1.0, 1.0333, 1.0, 0.9861, 1.0]),
    6: ([999, 758, 666, 495, 333, 107, 0], [1.0, 0.9924, 1.0, 0.9916, 1.0, 0.9906, 1.0],
        [1.0, 1.052, 1.0, 0.9998, 1.0, 0.9781, 1.0]),
    8: ([999, 831, 749, 623, 500, 394, 250, 88, 0], [1.0, 0.9976, 1.0, 0.991, 1.0, 0.9907, 1.0, 0.9905, 1.0],
        [1.0, 1.0257, 1.0, 0.9989, 1.0, 1.0022, 1.0, 0.9747, 1.0]),
}
SD15_SCHED = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1)


def gen_amed(name, *, n, B, shape, seed, amed=True, **cfg_over):
    """diffusers_amed_plugin_dpmpp.py (UNMODIFIED) over the stand-in of its diffusers base class — see
    ref_shim._dpm_base for what that pins and what it does not."""
    ref = ref_shim.load_reference()
    cfg = dict(SD15_SCHED, **cfg_over)
    s = ref.AMEDDPMSolverMultistepScheduler(**cfg)
    if amed:
        ts, dirs, times = AMED_SCHEDULES[n]
        s.scale_dirs, s.scale_times = dirs, times
        s.set_timesteps(n, timesteps=ts)
    else:
        ts, dirs, times = None, [1.0] * (n + 1), None
        s.scale_dirs = dirs                          # the plugin's step() reads it unconditionally (:417)
        s.set_timesteps(n)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, *shape, generator=g)
    d = dict(x_T=x.numpy(), timesteps=s.timesteps.numpy(), sigmas=s.sigmas.numpy())
    for i, t in enumerate(s.timesteps):
        e = torch.randn(B, *shape, generator=g)
        x = s.step(e, t, x, return_dict=False)[0]
        d[f"eps_{i}"] = e.numpy()
        d[f"prev_{i}"] = x.numpy()
    meta = dict(kind="amed", B=B, shape=list(shape), n=n, seed=seed, config=cfg, amed=amed, schedule=ts,
                scale_dirs=dirs, scale_times=times, dtype="float32", torch=torch.__version__)
    d["__meta__"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name)


def main_amed():
    """`python oracle/make_golden.py amed` writes only these."""
    small = (4, 8, 8)
    gen_amed("amed_n8_B3", n=8, B=3, shape=small, seed=80)                               # gen_ppo.py `type == "amed"`
    gen_amed("amed_n4_B2", n=4, B=2, shape=small, seed=81)
    gen_amed("amed_n6_heun_B2_ragged", n=6, B=2, shape=(3, 5, 7), seed=82, solver_type="heun")
    gen_amed("amed_n8_dpmsolver_B2", n=8, B=2, shape=small, seed=83, algorithm_type="dpmsolver")
    gen_amed("amed_n6_dpmsolver_heun_v_B2", n=6, B=2, shape=small, seed=84, algorithm_type="dpmsolver",
             solver_type="heun", prediction_type="v_prediction")
    gen_amed("amed_n8_v_B2", n=8, B=2, shape=small, seed=85, prediction_type="v_prediction")
    gen_amed("amed_n4_order1_sample_B2", n=4, B=2, shape=small, seed=86, solver_order=1, prediction_type="sample")
    gen_amed("amed_n8_order3_B2", n=8, B=2, shape=small, seed=97, solver_order=3)       # 1st, 2nd, 3rd x4, 2nd, 1st
    gen_amed("amed_stock_n16_order3_dpmsolver_v_B2", n=16, B=2, shape=small, seed=98, amed=False, solver_order=3,
             algorithm_type="dpmsolver", final_sigmas_type="sigma_min", prediction_type="v_prediction")
    # stock grids (no AMED schedule): gen_ppo.py `type == "dpm"` uses dpmsolver + final sigma_min
    gen_amed("amed_stock_n8_dpmsolver_sigmamin_B2", n=8, B=2, shape=small, seed=87, amed=False,
             algorithm_type="dpmsolver", final_sigmas_type="sigma_min")
    gen_amed("amed_stock_n5_trailing_zero_B2", n=5, B=2, shape=small, seed=88, amed=False, timestep_spacing="trailing")
    gen_amed("amed_stock_n20_leading_B1", n=20, B=1, shape=small, seed=89, amed=False, timestep_spacing="leading")


def gen_update_side(name, variant, seed, **kw):
    """FactorNetPPO.get_action_probs (factor_net_ppo.py:170-184): the PPO-update-side evaluation."""
    ref = ref_shim.load_reference()
    with ref_shim.quiet():
        fn = (ref.FactorNetPPO_SD if variant == "sd" else ref.FactorNetPPO_FM)(**kw)
    _seed_policy(fn, seed, 0.5 if variant == "sd" else 0.02)
    g = torch.Generator().manual_seed(seed)
    R = 14
    if variant == "sd":
        t = torch.randint(0, 1000, (R, 1), generator=g).float()
        x = torch.cat([t, t - 66], dim=1)
    else:
        x = torch.rand(R, 2, generator=g)
    with ref_shim.quiet():
        torch.manual_seed(seed)
        actions, _ = fn.sample_action({"x": x})
        probs, ent = fn(dict(x=x), actions)
    d = {f"sd.{k}": v.numpy() for k, v in fn.state_dict().items()}
    d.update(x=x.numpy(), actions=actions.numpy(), probs=probs.detach().numpy(), entropy=ent.detach().numpy())
    d["__meta__"] = np.array(json.dumps(dict(kind="update", variant=variant, kwargs=kw, torch=torch.__version__)))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name)


def main_fm_general():
    """`python oracle/make_golden.py fm_general` writes only these."""
    tok = (16, 8)
    for j, kind in enumerate(("euler", "heun", "dpm-solver", "dpm-solver-multistep")):
        tag = kind.replace("-", "")
        gen_fm_general(f"fmgen_{tag}_bf16_n8_B3", kind=kind, B=3, shape=tok, n=8, seed=50 + j, dtype=torch.bfloat16)
        gen_fm_general(f"fmgen_{tag}_f32_n7_B2", kind=kind, B=2, shape=(5, 7), n=7, seed=60 + j, dtype=torch.float32,
                       use_begin_index=False)
    gen_fm_general("fmgen_heun_f16_static_n6_B2", kind="heun", B=2, shape=tok, n=6, seed=70, dtype=torch.float16,
                   use_dynamic_shifting=False)


def main_cuda():
    """`python oracle/make_golden.py cuda` — run ON A GPU BOX (gpurun), where the reference is the byte-identical copy
    under oracle/_ref (oracle/stage_ref.py).  The reference's schedulers execute on cuda:0 the way its drivers run them;
    fixtures are prefixed `cuda_`.  They pin what a CPU run cannot:
      * ATen's CUDA treatment of `0-d CPU scalar (op) 16-bit CUDA tensor` (the schedule scalars of
        scheduler_ppo.py:309-330 stay on the host) — the sd16 flows,
      * gen_ppo.py:193-195,:309 — policy cast to fp16 (bins included) and everything under torch.autocast("cuda", fp16),
      * train_ppo.py:353 / edit_ppo/train_ppo.py:289 — fp32 policy evaluated under accelerator.autocast,
      * cuBLAS instead of MKL in the policy MLP (the FM softmax at temperature 0.01 amplifies the difference),
      * use_conv on 16-bit outputs, with the reference's own sampling.
    Same seeds as the CPU fixtures of the same configuration, so CPU-vs-CUDA spreads of the reference itself can be
    read off the pairs."""
    assert torch.cuda.is_available(), "main_cuda needs a GPU"
    dev = "cuda"
    small, tok = (4, 8, 8), (16, 8)
    f16, b16, f32 = torch.float16, torch.bfloat16, torch.float32
    # fp32 production flows
    gen_sd("cuda_sd_eps_s0_n8_B3", B=3, shape=small, n=8, guidance=3.0, seed=18, device=dev)
    gen_sd("cuda_sd_v_s2_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=7.5, seed=24, device=dev,
           prediction_type="v_prediction", scaler_dim=2)
    gen_sd("cuda_sd_eps_conv_s0_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=29, use_conv=True,
           device=dev)
    # 16-bit model outputs, fp32 policy, no autocast (the CPU sd16_* configurations, now with CUDA scalar semantics)
    gen_sd("cuda_sd16_f16_pipeline_eps_s0_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=90, dtype=f16,
           device=dev)
    gen_sd("cuda_sd16_bf16_pipeline_v_s0_n6_B2", hidden_dim=64, B=2, shape=small, n=6, guidance=3.0, seed=91, dtype=b16,
           prediction_type="v_prediction", device=dev)
    gen_sd("cuda_sd16_f16_autocastlayout_eps_s0_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=92,
           dtype=f16, latent_dtype=f32, device=dev)
    gen_sd("cuda_sd16_bf16_autocastlayout_v_s0_n5_B2_ragged", hidden_dim=64, B=2, shape=(3, 5, 7), n=5, guidance=3.0,
           seed=93, dtype=b16, latent_dtype=f32, prediction_type="v_prediction", device=dev)
    gen_sd("cuda_sd16_f16_pipeline_eps_s2_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=3.0, seed=94, dtype=f16,
           scaler_dim=2, device=dev)
    gen_sd("cuda_sd16_f16_pipeline_v_s1_n4_B2", hidden_dim=64, B=2, shape=small, n=4, guidance=3.0, seed=96, dtype=f16,
           scaler_dim=1, prediction_type="v_prediction", device=dev)
    gen_sd("cuda_sd16_bf16_autocastlayout_eps_s1_o3_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=3.0, seed=95,
           dtype=b16, latent_dtype=f32, scaler_dim=1, order_dim=3, device=dev)
    # gen_ppo.py's shipped inference flow: fp16 pipeline, policy (bins included) cast to fp16, all under autocast
    gen_sd("cuda_genppo_f16_eps_s0_n8_B3", B=3, shape=small, n=8, guidance=3.0, seed=101, dtype=f16, policy_dtype=f16,
           autocast=f16, device=dev)
    gen_sd("cuda_genppo_f16_eps_s0_n8_B2_full", B=2, shape=(4, 64, 64), n=8, guidance=7.5, seed=102, dtype=f16,
           policy_dtype=f16, autocast=f16, device=dev)
    gen_sd("cuda_genppo_f16_v_s2_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=3.0, seed=103, dtype=f16,
           policy_dtype=f16, autocast=f16, prediction_type="v_prediction", scaler_dim=2, device=dev)
    gen_sd("cuda_genppo_bf16_eps_s1_n6_B2", hidden_dim=64, B=2, shape=small, n=6, guidance=3.0, seed=104, dtype=b16,
           policy_dtype=b16, autocast=b16, scaler_dim=1, device=dev)
    # train_ppo.py:353 rollout: fp32 policy and latents, 16-bit U-Net output, step() inside accelerator.autocast
    gen_sd("cuda_rollout_f16_eps_s0_n8_B3", B=3, shape=small, n=8, guidance=3.0, seed=105, dtype=f16, latent_dtype=f32,
           autocast=f16, device=dev)
    gen_sd("cuda_rollout_bf16_v_s2_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=3.0, seed=106, dtype=b16,
           latent_dtype=f32, autocast=b16, prediction_type="v_prediction", scaler_dim=2, device=dev)
    # FM / FLUX flows on the GPU: same seeds as fm_* (CPU) for the reference-vs-reference spread
    gen_fm("cuda_fm_o2_s0_m0_bf16_n8_B3", B=3, shape=tok, n=8, seed=31, dtype=b16, device=dev)
    gen_fm("cuda_fm_o2_s0_m0_f32_n8_B3", hidden_dim=64, B=3, shape=tok, n=8, seed=32, dtype=f32, device=dev)
    gen_fm("cuda_fm_o4_s2_m1_bf16_n8_B3", B=3, shape=tok, n=8, seed=33, dtype=b16, order_dim=4, scaler_dim=2, mu_dim=1,
           device=dev)
    gen_fm("cuda_fm_o4_s0_m0_f32_n5_B2", hidden_dim=64, B=2, shape=tok, n=5, seed=34, dtype=f32, order_dim=4, device=dev)
    gen_fm("cuda_fm_o4_s1_m0_bf16_n6_B2", hidden_dim=64, B=2, shape=tok, n=6, seed=35, dtype=b16, order_dim=4,
           scaler_dim=1, device=dev)
    gen_fm("cuda_fm_conv_o4_s0_m0_bf16_n6_B2", hidden_dim=64, B=2, shape=tok, n=6, seed=38, dtype=b16, order_dim=4,
           use_conv=True, device=dev)
    gen_fm("cuda_fm_rollout_bf16_autocast_n8_B3", B=3, shape=tok, n=8, seed=39, dtype=b16, autocast=b16, device=dev)


def main_cuda2():
    """`python oracle/make_golden.py cuda2` (GPU box): fixtures that pin the ORDER in which ATen's CUDA reduce kernel
    adds the terms of `1 - torch.sum(torch.stack(...), dim=0)` (scheduler_ppo.py:172): batch size 1 (the reduced
    dimension is the fastest one: a shuffle tree) and more than four terms (order_dim 6 / 8: four accumulators)."""
    assert torch.cuda.is_available(), "main_cuda2 needs a GPU"
    dev, small, tok = "cuda", (4, 8, 8), (16, 8)
    gen_sd("cuda_sd_eps_s0_n8_B1", hidden_dim=64, B=1, shape=small, n=8, guidance=3.0, seed=111, device=dev)
    gen_sd("cuda_sd_v_o8_s1_n12_B2", hidden_dim=64, B=2, shape=small, n=12, guidance=3.0, seed=112, device=dev,
           order_dim=8, scaler_dim=1, prediction_type="v_prediction")
    gen_sd("cuda_sd_eps_o6_s0_n10_B1", hidden_dim=64, B=1, shape=small, n=10, guidance=3.0, seed=113, device=dev,
           order_dim=6)
    gen_sd("cuda_sd_eps_o8_s2_n12_B1", hidden_dim=64, B=1, shape=small, n=12, guidance=3.0, seed=114, device=dev,
           order_dim=8, scaler_dim=2)
    gen_fm("cuda_fm_o4_s0_m0_f32_n6_B1", hidden_dim=64, B=1, shape=tok, n=6, seed=115, dtype=torch.float32, order_dim=4,
           device=dev)
    gen_fm("cuda_fm_o6_s1_m0_bf16_n8_B2", hidden_dim=64, B=2, shape=tok, n=8, seed=116, dtype=torch.bfloat16, order_dim=6,
           scaler_dim=1, device=dev)


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["cuda"]:
        return main_cuda()
    if sys.argv[1:] == ["cuda2"]:
        return main_cuda2()
    if sys.argv[1:] == ["fm_general"]:
        return main_fm_general()
    if sys.argv[1:] == ["amed"]:
        return main_amed()
    if sys.argv[1:] == ["sd16"]:
        return main_sd16()
    small = (4, 8, 8)
    # --- SD / PPOScheduler: production config at several step counts (warm-up depths, n=7 quirk) -------
    for n in (2, 5, 7, 8, 12):
        gen_sd(f"sd_eps_s0_n{n}_B3", B=3, shape=small, n=n, guidance=3.0, seed=10 + n,
               hidden_dim=256 if n == 8 else 64)
    gen_sd("sd_eps_s0_n8_B1_full", B=1, shape=(4, 64, 64), n=8, guidance=3.0, seed=3)       # BASELINE config 0
    gen_sd("sd_eps_s0_n8_B64", hidden_dim=64, B=64, shape=(4, 4, 4), n=8, guidance=3.0, seed=4)
    # --- scaler dims, v-prediction, other spacings / schedules / orders -------------------------------
    gen_sd("sd_eps_s1_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=21, scaler_dim=1)
    gen_sd("sd_eps_s2_n8_B3", B=3, shape=small, n=8, guidance=3.0, seed=22, scaler_dim=2)
    gen_sd("sd_v_s0_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=23, prediction_type="v_prediction")
    gen_sd("sd_v_s2_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=7.5, seed=24, prediction_type="v_prediction",
           scaler_dim=2)
    gen_sd("sd_eps_leading_linear_o3_n6_B2", hidden_dim=64, B=2, shape=small, n=6, guidance=1.5, seed=25,
           timestep_spacing="leading", beta_schedule="linear", beta_start=1e-4, beta_end=0.02, order_dim=3,
           scaler_dim=1)
    gen_sd("sd_eps_linspace_cos_o2_n9_B2", hidden_dim=64, B=2, shape=small, n=9, guidance=3.0, seed=26,
           timestep_spacing="linspace", beta_schedule="squaredcos_cap_v2", order_dim=2, scaler_dim=0)
    gen_sd("sd_eps_k161_o4_s2_n4_B2", hidden_dim=64, B=2, shape=small, n=4, guidance=3.0, seed=27, scaler_dim=2,
           num_actions=161, last_std=0.2)                                                     # ctor defaults
    gen_sd("sd_eps_s0_n8_B2_ragged", hidden_dim=64, B=2, shape=(3, 5, 7), n=8, guidance=3.0, seed=28)        # N_s % 4 != 0
    gen_sd("sd_eps_conv_s0_n8_B3", hidden_dim=64, B=3, shape=small, n=8, guidance=3.0, seed=29, use_conv=True)
    gen_sd("sd_v_conv_s2_o3_n5_B2", hidden_dim=64, B=2, shape=small, n=5, guidance=2.0, seed=30, use_conv=True,
           order_dim=3, scaler_dim=2, prediction_type="v_prediction")
    # --- FM / FMPPOScheduler --------------------------------------------------------------------------
    tok = (16, 8)
    gen_fm("fm_o2_s0_m0_bf16_n8_B3", B=3, shape=tok, n=8, seed=31, dtype=torch.bfloat16)     # FLUX prod
    gen_fm("fm_o2_s0_m0_f32_n8_B3", hidden_dim=64, B=3, shape=tok, n=8, seed=32, dtype=torch.float32)
    gen_fm("fm_o4_s2_m1_bf16_n8_B3", B=3, shape=tok, n=8, seed=33, dtype=torch.bfloat16, order_dim=4,
           scaler_dim=2, mu_dim=1)                                                            # ctor defaults
    gen_fm("fm_o4_s0_m0_f32_n5_B2", hidden_dim=64, B=2, shape=tok, n=5, seed=34, dtype=torch.float32, order_dim=4)
    gen_fm("fm_o4_s1_m0_bf16_n6_B2", hidden_dim=64, B=2, shape=tok, n=6, seed=35, dtype=torch.bfloat16, order_dim=4,
           scaler_dim=1)
    gen_fm("fm_o2_s0_m0_bf16_n4_B2_search", hidden_dim=64, B=2, shape=tok, n=4, seed=36, dtype=torch.bfloat16,
           use_begin_index=False)
    gen_fm("fm_conv_o4_s0_m0_bf16_n6_B2", hidden_dim=64, B=2, shape=tok, n=6, seed=38, dtype=torch.bfloat16, order_dim=4,
           use_conv=True)
    gen_fm("fm_o2_s0_m0_bf16_n3_B1_full", B=1, shape=(4096, 64), n=3, seed=37, dtype=torch.bfloat16)
    # --- PPO-update side ------------------------------------------------------------------------------
    gen_update_side("update_sd_o4_s0", "sd", 41, hidden_dim=256, num_actions=11, order_dim=4, scaler_dim=0)
    gen_update_side("update_fm_o2_s0_m0", "fm", 42, hidden_dim=256, num_actions=11, order_dim=2, scaler_dim=0,
                    mu_dim=0)
    # --- baseline flow-matching solvers -------------------------------------------------
    main_fm_general()
    main_amed()
    main_sd16()


if __name__ == "__main__":
    main()
