"""BASELINE config 5: PPO training rollout — batched multi-seed previews + synthetic reward + policy update with a
flat-buffer gradient all-reduce over NVLink.  Launch with torchrun for N > 1.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import consolver_b200 as cb  # noqa: E402
from consolver_b200 import ppo, sharding  # noqa: E402

PROD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000,
            steps_offset=1, timestep_spacing="trailing", order_dim=4, scaler_dim=0, use_conv=False,
            factor_net_kwargs=dict(embedding_dim=64, hidden_dim=256, num_actions=11))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=80)          # run_ppo.sh batch
    ap.add_argument("--ppo-epochs", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = sharding.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(1234 + rank)                            # per-rank seeds (edit_ppo/train_ppo.py:76)
    s = cb.PPOScheduler(**PROD)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    s.factor_net.to(dev)
    flat = ppo.FlatParams(s.factor_net)
    ppo.broadcast_parameters(flat, 0)
    opt = torch.optim.AdamW(s.factor_net.parameters(), lr=1e-4)
    g = torch.Generator(device=dev).manual_seed(rank)
    w = torch.randn(4, 4, device=dev, generator=g) * 0.3
    den = lambda x, t, i: torch.einsum("oc,bchw->bohw", w, x)  # noqa: E731  stand-in denoiser (not the product)
    noise = torch.randn(4, 64, 64, device=dev, generator=g)
    target = torch.randn(4, 64, 64, device=dev, generator=g)

    def one(it):
        n = ppo.shared_step_count(it, seed=0)
        lat, rec = ppo.rollout_sd(s, den, noise, a.batch, 3.0, n)
        r = ppo.latent_mse_reward(lat, target.unsqueeze(0).expand_as(lat))
        return ppo.ppo_update(s.factor_net, flat, opt, rec, r, ppo_epochs=a.ppo_epochs, entropy_coef=0.01), n

    for it in range(3):
        one(it)
    # all-reduce latency on the flat gradient buffer alone
    ar_us = None
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ppo.allreduce_gradients(flat)
        e1.record()
        torch.cuda.synchronize()
        ar_us = e0.elapsed_time(e1) * 1e3 / 50
        flat.grad.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 0
    for it in range(3, 3 + a.iters):
        st, n = one(it)
        steps += n
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = sharding.gather_job_stats(a.iters * a.batch, dt, flat.checksum(), device=dev)
    if rank == 0:
        print(json.dumps({"metric": "ppo_rollout_previews_per_s", "value": round(stats["total"] / stats["max_elapsed_s"], 1),
                          "n_gpus": world, "iters": a.iters, "batch_per_gpu": a.batch, "ppo_epochs": a.ppo_epochs,
                          "allreduce_us_300KB": None if ar_us is None else round(ar_us, 2),
                          "param_checksum_identical_across_ranks": abs(stats["checksum"] / world - flat.checksum()) < 1e-6,
                          "grad_buffer_floats": flat.numel, "last_loss": st["loss"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
