"""The plain-C restatement (oracle/consolver_oracle.c) against the reference's golden vectors and against the
torch-CPU oracle: two independent CPU statements of the path must agree bit-for-bit on everything that is not a
BLAS dot product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import consolver_oracle as orc
from golden_io import Golden, names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))


def _p(t):
    return C.c_void_p(t.data_ptr())


def _ptrs(ts):
    arr = (C.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


@pytest.mark.parametrize("name", [n for n in names("sd_")] + [n for n in names("cuda_sd_")])
def test_c_oracle_sd_trajectory_matches_reference(lib, name):
    g = Golden(name)
    m = g.meta
    cfg = m["config"]
    od, sdim = cfg["order_dim"], cfg["scaler_dim"]
    s = orc.OracleSDScheduler(g.state_dict, **cfg)      # only for the schedule scalars
    s.set_timesteps(m["n"])
    B = m["B"]
    N = int(np.prod(m["shape"]))
    av = g.state_dict["action_values"].contiguous()
    A, K = av.shape
    x = g["x_T"].contiguous()
    hist = []
    flags = (1 if cfg.get("prediction_type", "epsilon") == "v_prediction" else 0) | (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0)
    on_gpu = m.get("device") == "cuda"          # fixture made by the reference running on a B200: ATen's CUDA rules
    flags |= 0 if on_gpu else 256               # 256 = CONSOLVER_FLAG_HOST_SCALARS: true division
    sum_mode = 0 if not on_gpu else (2 if B == 1 else 1)      # the order ATen adds the closing coefficient's terms in
    for i, t in enumerate(s.timesteps):
        pair = g[f"pair_{i}"].contiguous()
        eps = torch.empty_like(x)
        lib.oracle_cfg_f32(_p(pair[:B]), _p(pair[B:]), C.c_float(m["guidance"]), _p(eps), C.c_int64(B * N))
        assert torch.equal(eps, g[f"eps_{i}"])
        hist = ([eps] + hist)[:od]
        n_hist = len(hist)
        q = g[f"q_{i}"].contiguous()
        idx = torch.empty(B, A, dtype=torch.int64)
        actions, probs, masks = torch.empty(B, A), torch.empty(B, A), torch.empty(B, A)
        coef = torch.empty(B, od + 2)
        if not cfg.get("use_conv"):
            table = g[f"probs_full_{i}"][0].contiguous()     # the reference's own softmax table (rows identical)
            lib.oracle_policy_sample(_p(table), _p(av), _p(q), B, A, K, od, sdim, n_hist, sum_mode, _p(idx), _p(actions),
                                     _p(probs), _p(masks), _p(coef))
        else:                                                # use_conv: one table per sample
            for b in range(B):
                table = g[f"probs_full_{i}"][b].contiguous()
                lib.oracle_policy_sample(_p(table), _p(av), _p(q[b * A:(b + 1) * A]), 1, A, K, od, sdim, n_hist,
                                         sum_mode, _p(idx[b]), _p(actions[b]), _p(probs[b]), _p(masks[b]), _p(coef[b]))
        assert torch.equal(idx, g[f"idx_{i}"])
        assert torch.equal(actions, g[f"actions_{i}"]) and torch.equal(probs, g[f"probs_{i}"])
        assert torch.equal(masks, g[f"masks_{i}"])
        tt = int(t)
        sc = [float(v) for v in orc.ddim_scalars(s.alphas_cumprod, tt, orc.sd_prev_timestep(tt, m["n"]))]
        out = torch.empty_like(x)
        lib.oracle_sd_step_f32(_ptrs(hist), n_hist, _p(x), _p(out), _p(coef), od, *[C.c_float(v) for v in sc], flags,
                               B, C.c_int64(N))
        assert torch.equal(out, g[f"prev_{i}"]), f"step {i}"
        x = out


@pytest.mark.parametrize("name", [n for n in names("fm_") + names("cuda_fm_") if "f32" in n and "conv" not in n])
def test_c_oracle_fm_trajectory_matches_reference(lib, name):
    g = Golden(name)
    m = g.meta
    cfg = m["config"]
    od, sdim = cfg["order_dim"], cfg["scaler_dim"]
    B, N = m["B"], int(np.prod(m["shape"]))
    av = g.state_dict["action_values"].contiguous()
    A, K = av.shape
    sig = g["sigmas"]
    x = g["x_T"].contiguous()
    hist = []
    flags = (2 if sdim >= 1 else 0) | (4 if sdim >= 2 else 0)
    for i in range(m["n"]):
        hist = ([g[f"v_{i}"].contiguous()] + hist)[:od]
        n_hist = len(hist)
        idx = torch.empty(B, A, dtype=torch.int64)
        actions, probs, masks = torch.empty(B, A), torch.empty(B, A), torch.empty(B, A)
        coef = torch.empty(B, od + 2)
        sum_mode = 0 if m.get("device") != "cuda" else (2 if B == 1 else 1)
        lib.oracle_policy_sample(_p(g[f"probs_full_{i}"][0].contiguous()), _p(av), _p(g[f"q_{i}"].contiguous()), B, A, K,
                                 od, sdim, n_hist, sum_mode, _p(idx), _p(actions), _p(probs), _p(masks), _p(coef))
        assert torch.equal(idx, g[f"idx_{i}"]) and torch.equal(masks, g[f"masks_{i}"])
        dt = float(sig[i + 1] - sig[i])
        out = torch.empty_like(x)
        lib.oracle_fm_step_f32(_ptrs(hist), n_hist, _p(x), _p(out), _p(coef), od, C.c_float(dt), flags, B, C.c_int64(N))
        assert torch.equal(out, g[f"prev_{i}"]), f"step {i}"
        x = out


def test_c_oracle_policy_table_close_to_torch_oracle(lib):
    g = Golden("sd_eps_s0_n8_B3")
    sd = {k: v.contiguous() for k, v in g.state_dict.items()}
    A, K = sd["action_values"].shape
    H = sd["mlp.0.weight"].shape[0]
    h1, h2, probs = torch.empty(H), torch.empty(H), torch.empty(A, K)
    lib.oracle_policy_table(*[_p(sd[k]) for k in ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias",
                                                  "mlp.4.weight", "mlp.4.bias")],
                            C.c_float(874.0), C.c_float(749.0), C.c_float(999.0), C.c_float(1.0), H, A, K,
                            _p(h1), _p(h2), _p(probs))
    torch.testing.assert_close(probs, g["probs_full_1"][0], rtol=0, atol=1e-6)
