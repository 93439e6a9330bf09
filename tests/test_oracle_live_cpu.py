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
