"""Re-run chosen cases of tests/test_gpu_live_reference.py::test_wide_random_sd_configurations_step_for_step and print what
differs (per-sample error, the coefficient record, the reference's actions).  python tools/exp/fuzz_debug.py 277 289"""
import contextlib
import os
import random
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import ref_shim  # noqa: E402
import test_gpu_live_reference as T  # noqa: E402


def run(case):
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
    print(f"== case {case}: {flow} K={K} H={hidden} n={n} B={B} {shape} style={style} fused={fused_cfg} g={guidance} "
          f"last_std={last_std} {cfg}")
    r, o = T._wide_pair("sd", case, hidden, K, last_std, **cfg)
    mdt = torch.float32 if flow == "f32" else (torch.float16 if "f16" in flow and "bf16" not in flow else torch.bfloat16)
    xdt = mdt if flow.endswith("_pipeline") or flow.startswith("genppo") else torch.float32
    ac = mdt if flow.startswith(("autocast", "genppo")) else None
    if flow.startswith("genppo"):
        r.factor_net.to("cuda", dtype=mdt), o.factor_net.to("cuda", dtype=mdt)
    r.set_timesteps(n, device="cuda"), o.set_timesteps(n, device="cuda")
    if os.environ.get("DBG_NOPDL") == "1":
        o.use_pdl = False
    if os.environ.get("DBG_NOFUSEDRNG") == "1":
        o.use_fused_rng = False
    if os.environ.get("DBG_SYNC") == "1":
        torch.cuda.synchronize()
    g = torch.Generator().manual_seed(case)
    xr = xo = torch.randn(B, *shape, generator=g).to(xdt).cuda()
    ctx = (lambda: torch.autocast("cuda", ac)) if ac is not None else contextlib.nullcontext
    for i in range(n):
        pair = torch.randn(2 * B, *shape, generator=g).to(mdt).cuda()
        u, c = pair.chunk(2)
        e = u + guidance * (c - u)
        x_in = xr
        torch.manual_seed(77 + i)
        with ref_shim.quiet(), ctx(), torch.no_grad():
            xr, ar, pr, cr, mr = r.step(e, T._hand_over(style, r, i), xr, return_dict=False)
        torch.manual_seed(77 + i)
        with ctx(), torch.no_grad():
            if fused_cfg:
                xo, ao, po, co, mo = o.step_cfg(pair, T._hand_over(style, o, i), xo, guidance)
            else:
                xo, ao, po, co, mo = o.step(e, T._hand_over(style, o, i), xo, return_dict=False)
        same_a = torch.equal(ao, ar)
        dp = (po.float() - pr.float()).abs().max().item()
        same_x = xo.dtype == xr.dtype and torch.equal(xo, xr)
        print(f" step {i}: actions {'==' if same_a else '!='}  max|dp| {dp:.3e}  latent {'==' if same_x else '!='} "
              f"({xo.dtype} vs {xr.dtype})  t={int(r.timesteps[i])}")
        if not same_x:
            d = (xo.float() - xr.float()).flatten(1)
            print("   per-sample max|dx|:", [f"{v:.3e}" for v in d.abs().max(dim=1).values.tolist()][:8])
            print("   finite ref/ours:", torch.isfinite(xr).all().item(), torch.isfinite(xo).all().item(),
                  " nan-equal:", torch.equal(torch.nan_to_num(xo.float(), 7.0, 8.0, -8.0),
                                             torch.nan_to_num(xr.float(), 7.0, 8.0, -8.0)))
            lp = o.last_policy()
            print("   coef[0]:", lp["coef"][0].tolist(), " ref actions[0]:", ar[0].tolist(), ar.dtype)
            print("   ours[0,:6]:", xo.flatten(1)[0, :6].tolist())
            print("   ref [0,:6]:", xr.flatten(1)[0, :6].tolist())
            print("   x_in[0,:6]:", x_in.flatten(1)[0, :6].tolist(), " e[0,:6]:", e.flatten(1)[0, :6].tolist())
            sa, sb = float(o._sqrt_abar[int(r.timesteps[i])]), float(o._sqrt_1m_abar[int(r.timesteps[i])])
            print(f"   sqrt_abar_t {sa:.6e} sqrt_1m {sb:.6e}")
            break


if __name__ == "__main__":
    for c in sys.argv[1:]:
        run(int(c))
