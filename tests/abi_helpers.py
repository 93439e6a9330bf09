"""Thin torch<->C-ABI helpers for the tests: every call goes through libconsolver.so exactly as a foreign
host (cgo/JNI/ctypes) would — raw device pointers, sizes, a stream handle."""
import ctypes

import torch

from consolver_b200 import _lib


def sd_to_dev(sd, device="cuda"):
    return {k: v.to(device=device, dtype=torch.float32).contiguous() for k, v in sd.items()}


def weights(sd):
    return [sd[k].data_ptr() for k in ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight",
                                      "mlp.4.bias", "action_values")]


# `host=True` (default of these helpers): torch's CPU rules (CONSOLVER_POLICY_HOST_DIV / CONSOLVER_FLAG_HOST_SCALARS),
# the rules the oracle's default functions follow; host=False: ATen's CUDA rules (oracle: sem=orc.CUDA).
def policy(sd, x0, x1, x_div, temp, B, order_dim, scaler_dim, n_hist, q=None, idx_in=None, feat=None, host=True,
           policy_flags=0):
    lib = _lib.load()
    policy_flags |= _lib.POLICY_HOST_DIV if host else 0
    A, K = sd["action_values"].shape
    H = sd["mlp.0.weight"].shape[0]
    dev = sd["action_values"].device
    f = dict(device=dev, dtype=torch.float32)
    out = dict(probs_table=torch.full((B, A, K) if feat is not None else (A, K), -1.0, **f), idx=torch.full((B, A), -1, device=dev, dtype=torch.int64),
               actions=torch.zeros(B, A, **f), probs=torch.zeros(B, A, **f), logp=torch.zeros(B, A, **f),
               masks=torch.zeros(B, A, **f), coef=torch.zeros(B, order_dim + 2, **f))
    rc = lib.consolver_policy_f32(
        *weights(sd), float(x0), float(x1), float(x_div), float(temp),
        feat.data_ptr() if feat is not None else None, feat.shape[1] if feat is not None else 0,
        q.data_ptr() if q is not None else None, idx_in.data_ptr() if idx_in is not None else None,
        B, H, A, K, order_dim, scaler_dim, n_hist, policy_flags,
        out["probs_table"].data_ptr(), out["idx"].data_ptr(), out["actions"].data_ptr(), out["probs"].data_ptr(),
        out["logp"].data_ptr(), out["masks"].data_ptr(), out["coef"].data_ptr(),
        torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "consolver_policy_f32")
    return out


def step_sd(e0, cond, guidance, hist, x, coef, order_dim, scalars, flags=0, slot=False, out2=None, host=True):
    lib = _lib.load()
    flags |= _lib.FLAG_HOST_SCALARS if host else 0
    B = x.shape[0]
    N = x.numel() // B
    x_out = torch.empty_like(x)
    slot_t = torch.empty_like(e0) if slot else None
    if x.dtype != e0.dtype:                      # fp32 latents, 16-bit model outputs (autocast pipelines)
        assert x.dtype == torch.float32
        flags |= _lib.FLAG_X_F32
    rc = lib.consolver_step_sd(
        _lib.dtype_code(e0.dtype), e0.data_ptr(), cond.data_ptr() if cond is not None else None, float(guidance),
        slot_t.data_ptr() if slot else None, _lib.ptr_array([h.data_ptr() for h in hist]), len(hist) + 1,
        x.data_ptr(), x_out.data_ptr(), out2.data_ptr() if out2 is not None else None,
        out2.stride(0) if out2 is not None else 0, coef.data_ptr(), coef.shape[1], order_dim,
        *[float(s) for s in scalars], flags, B, N, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "consolver_step_sd")
    return x_out, slot_t


def step_fm(e0, hist, x, coef, order_dim, dt, flags=0, out2=None):
    lib = _lib.load()
    B = x.shape[0]
    N = x.numel() // B
    x_out = torch.empty(x.shape, device=x.device, dtype=e0.dtype)
    rc = lib.consolver_step_fm(
        _lib.dtype_code(e0.dtype), _lib.dtype_code(x.dtype), e0.data_ptr(), None,
        _lib.ptr_array([h.data_ptr() for h in hist]), len(hist) + 1, x.data_ptr(), x_out.data_ptr(),
        out2.data_ptr() if out2 is not None else None, out2.stride(0) if out2 is not None else 0,
        coef.data_ptr(), coef.shape[1], order_dim, float(dt), flags, B, N, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "consolver_step_fm")
    return x_out


def policy_table(sd, x_rows, x_div, temp, host=True, policy_flags=0):
    lib = _lib.load()
    policy_flags |= _lib.POLICY_HOST_DIV if host else 0
    A, K = sd["action_values"].shape
    H = sd["mlp.0.weight"].shape[0]
    x_rows = x_rows.to(device=sd["action_values"].device, dtype=torch.float32).contiguous()
    out = torch.full((x_rows.shape[0], A, K), -1.0, device=x_rows.device)
    rc = lib.consolver_policy_table_f32(*weights(sd)[:6], x_rows.data_ptr(), x_rows.shape[0], float(x_div), float(temp),
                                        H, A, K, policy_flags, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "consolver_policy_table_f32")
    return out


def policy_sample(sd, table, B, order_dim, scaler_dim, n_hist, q=None, idx_in=None, rng=None, q_out=None,
                  policy_flags=0):
    lib = _lib.load()
    A, K = sd["action_values"].shape
    dev = table.device
    f = dict(device=dev, dtype=torch.float32)
    out = dict(idx=torch.full((B, A), -1, device=dev, dtype=torch.int64), actions=torch.zeros(B, A, **f),
               probs=torch.zeros(B, A, **f), logp=torch.zeros(B, A, **f), masks=torch.zeros(B, A, **f),
               coef=torch.zeros(B, order_dim + 2, **f))
    rc = lib.consolver_policy_sample_f32(
        table.data_ptr(), sd["action_values"].data_ptr(), q.data_ptr() if q is not None else None,
        idx_in.data_ptr() if idx_in is not None else None,
        ctypes.byref(rng) if rng is not None else None, q_out.data_ptr() if q_out is not None else None,
        B, A, K, order_dim, scaler_dim, n_hist, policy_flags,
        out["idx"].data_ptr(), out["actions"].data_ptr(), out["probs"].data_ptr(), out["logp"].data_ptr(),
        out["masks"].data_ptr(), out["coef"].data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "consolver_policy_sample_f32")
    return out


def cosine_features(e0, cond, guidance, hist, order_dim):
    lib = _lib.load()
    B = e0.shape[0]
    N = e0.numel() // B
    nbytes = lib.consolver_cosine_features_workspace(B, order_dim)
    ws = torch.empty(nbytes // 8 + 1, device=e0.device, dtype=torch.float64)
    feat = torch.full((B, order_dim - 1), -7.0, device=e0.device)
    rc = lib.consolver_cosine_features(_lib.dtype_code(e0.dtype), e0.data_ptr(),
                                       cond.data_ptr() if cond is not None else None, float(guidance),
                                       _lib.ptr_array([h.data_ptr() for h in hist]), len(hist) + 1, order_dim, B, N,
                                       ws.data_ptr(), feat.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "consolver_cosine_features")
    return feat
