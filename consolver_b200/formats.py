"""On-disk formats either side of the hot path — host-side I/O only, no kernels.

  * policy checkpoint `model.ckpt`: `torch.save(factor_net.state_dict())` (train_ppo.py:174-178), loaded with
    `factor_net.load_state_dict(torch.load(path))` and optionally cast to fp16 afterwards (gen_ppo.py:188-195).
    Keys: `action_values`, `mlp.{0,2,4}.{weight,bias}`.
  * SD teacher pairs (gen_pretrain/generate_data.py:192-213, read by data_processing.py:38-61):
    `{id}.txt` (prompt), `{id}.png`, `noise_{id}.pth` (initial noise [4,64,64]) and `latent_{id}.pth` (teacher latent),
    where id = `{device_id}_{index:08d}`.
  * FLUX teacher pairs (edit_ppo/edit_pretrain/generate.py:81,133): `initial_noises/{i}.pt`, `obtained_noises/{i}.pt`
    (packed latents [1,4096,64]).
"""
from __future__ import annotations

import os
import random
from typing import Dict, List, Optional, Tuple

import torch

CKPT_KEYS = ("action_values", "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")


def save_policy(factor_net, path: str) -> None:
    """Writes exactly what the reference's save hook writes: the bare state_dict (train_ppo.py:177)."""
    fn = factor_net.module if hasattr(factor_net, "module") else factor_net
    sd = {k: v.detach().cpu() for k, v in fn.state_dict().items()}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(sd, path)


def load_policy(factor_net, path: str, dtype: Optional[torch.dtype] = None, device=None, strict: bool = True) -> Dict:
    """`factor_net.load_state_dict(torch.load(path))` (+ the optional `.to(device, dtype)` of gen_ppo.py:193-195).
    Validates the key set / shapes first so a checkpoint for another (order_dim, scaler_dim, num_actions,
    hidden_dim) fails with a readable message."""
    fn = factor_net.module if hasattr(factor_net, "module") else factor_net
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(sd, dict) or set(sd.keys()) != set(CKPT_KEYS):
        raise ValueError(f"{path}: not a ConsistencySolver policy checkpoint (keys {sorted(sd)[:8]})")
    own = fn.state_dict()
    bad = [f"{k}: file {tuple(sd[k].shape)} vs policy {tuple(own[k].shape)}" for k in CKPT_KEYS
           if tuple(sd[k].shape) != tuple(own[k].shape)]
    if bad:
        raise ValueError(f"{path}: shape mismatch — " + "; ".join(bad))
    fn.load_state_dict(sd, strict=strict)
    if dtype is not None or device is not None:
        fn.to(device=device, dtype=dtype)
    return sd


def policy_hparams_from_ckpt(path: str, variant: str = "sd", mu_dim: int = 0) -> Dict[str, int]:
    """Recover (hidden_dim, num_actions, action_dims) from a checkpoint's shapes; order_dim/scaler_dim follow from
    the bin table: rows holding linspace(-0.05, 0.05) are scaler dims."""
    sd = torch.load(path, map_location="cpu", weights_only=True)
    av = sd["action_values"].float()
    A, K = av.shape
    scaler = sum(1 for i in range(A) if abs(av[i, 0] + 0.05) < 1e-6 and abs(av[i, -1] - 0.05) < 1e-6)
    mu = mu_dim if variant == "fm" else 0
    return dict(hidden_dim=sd["mlp.0.weight"].shape[0], num_actions=K, action_dims=A, scaler_dim=scaler,
                order_dim=A - scaler - mu + 1, use_conv=sd["mlp.0.weight"].shape[1] > 2)


# ---- teacher pairs ---------------------------------------------------------------------------------------------
def sd_pair_ids(data_dir: str) -> List[str]:
    """ids of complete SD teacher pairs: every `{id}.txt` with both tensors present (data_processing.py:18-23)."""
    ids = []
    for f in sorted(os.listdir(data_dir)):
        if f.endswith(".txt"):
            i = f[:-4]
            if os.path.exists(os.path.join(data_dir, f"noise_{i}.pth")) and \
                    os.path.exists(os.path.join(data_dir, f"latent_{i}.pth")):
                ids.append(i)
    return ids


def load_sd_pair(data_dir: str, pair_id: str) -> Tuple[str, torch.Tensor, torch.Tensor]:
    """(prompt, noise, teacher latent) of one pair; NaN latents are rejected like data_processing.py:52-53."""
    with open(os.path.join(data_dir, pair_id + ".txt")) as f:
        text = f.read().strip()
    noise = torch.load(os.path.join(data_dir, f"noise_{pair_id}.pth"), map_location="cpu", weights_only=True)
    latent = torch.load(os.path.join(data_dir, f"latent_{pair_id}.pth"), map_location="cpu", weights_only=True)
    if torch.isnan(latent).any():
        raise ValueError(f"latent_{pair_id}.pth contains NaN")
    return text, noise, latent


def save_sd_pair(data_dir: str, pair_id: str, prompt: str, noise: torch.Tensor, latent: torch.Tensor) -> None:
    os.makedirs(data_dir, exist_ok=True)
    with open(os.path.join(data_dir, pair_id + ".txt"), "w") as f:
        f.write(prompt)
    torch.save(noise.clone(), os.path.join(data_dir, f"noise_{pair_id}.pth"))      # .clone(): generate_data.py:210
    torch.save(latent.clone(), os.path.join(data_dir, f"latent_{pair_id}.pth"))


def load_flux_pair(root: str, index: int, initial="initial_noises", obtained="obtained_noises"):
    a = torch.load(os.path.join(root, initial, f"{index}.pt"), map_location="cpu", weights_only=True)
    b = torch.load(os.path.join(root, obtained, f"{index}.pt"), map_location="cpu", weights_only=True)
    return a, b


def repeat_random_sample(noise: torch.Tensor, target: torch.Tensor, texts: List[str], rng: random.Random = random):
    """data_processing.py:65-80: pick ONE element of the batch and replicate it batch-size times, so a PPO batch
    differs only by the sampled actions.  Returns expanded VIEWS (no copies; `rollout_sd` makes them contiguous)."""
    B = noise.shape[0]
    i = rng.randint(0, B - 1)
    return (noise[i:i + 1].expand(B, *noise.shape[1:]), target[i:i + 1].expand(B, *target.shape[1:]), [texts[i]] * B, i)
