"""GPU, N > 1 (skipped on a single-GPU box): one process per GPU over NCCL — sharded sampling needs no collective,
the PPO update keeps replicas bit-identical through one flat-buffer all-reduce per epoch."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

PROD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, num_train_timesteps=1000,
            steps_offset=1, timestep_spacing="trailing", order_dim=4, scaler_dim=0, use_conv=False,
            factor_net_kwargs=dict(embedding_dim=64, hidden_dim=256, num_actions=11))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import consolver_b200 as cb
    from consolver_b200 import ppo, sharding

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sharding.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    torch.manual_seed(100 + rank)                      # different init and different rollouts per rank
    s = cb.PPOScheduler(**PROD)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    s.factor_net.to(dev)
    flat = ppo.FlatParams(s.factor_net)
    ppo.broadcast_parameters(flat, 0)
    opt = torch.optim.SGD(s.factor_net.parameters(), lr=0.05)
    w = torch.randn(4, 4, device=dev) * 0.3
    den = lambda x, t, i: torch.einsum("oc,bchw->bohw", w, x)  # noqa: E731
    lo, hi = sharding.shard_bounds_sd(10, rank, world)         # prompt/seed shard of this rank
    done = 0
    for it, seed in enumerate(range(lo, hi)):
        torch.manual_seed(1000 + seed)
        noise = torch.randn(4, 16, 16, device=dev)
        lat, rec = ppo.rollout_sd(s, den, noise, 8, 3.0, ppo.shared_step_count(it, 0, 3, 6))
        ppo.ppo_update(s.factor_net, flat, opt, rec, ppo.latent_mse_reward(lat, torch.zeros_like(lat)), ppo_epochs=1,
                       entropy_coef=0.01)
        done += lat.shape[0]
    stats = sharding.gather_job_stats(done, 1.0, flat.checksum(), device=dev)
    q.put((rank, flat.checksum(), stats["total"], stats["per_rank"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_nccl_rollouts_and_update_keep_replicas_identical():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1], "parameter checksums differ across ranks"
    assert res[0][2] == 80 and res[0][3] == [40, 40]


def _exchange_worker(rank, world, port, q):
    """the fused one-shot exchange (csrc/ppo.cu::ppo_reduce_allreduce_kernel) against NCCL's all-reduce(AVG) on the same
    per-rank gradients, several epochs in a row (both buffer parities, monotonic signals)"""
    import torch.distributed as dist

    import consolver_b200 as cb
    from consolver_b200 import ppo, sharding

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sharding.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    torch.manual_seed(7)                                # same policy on every rank ...
    s = cb.PPOScheduler(**PROD)
    with torch.no_grad():
        s.factor_net.mlp[4].weight.normal_(0, 0.05)
    s.factor_net.to(dev)
    flat = ppo.FlatParams(s.factor_net)
    ex = ppo.PeerGradExchange(flat)
    R, B, A = 7, 24, 3
    worst, same = 0.0, True
    for it in range(5):
        g = torch.Generator(device=dev).manual_seed(1000 * it + rank)          # ... different rollouts per rank
        x_rows = torch.tensor([[999.0 - 125 * r, 874.0 - 125 * r] for r in range(R)], device=dev)
        idx = torch.randint(0, 11, (R, B, A), device=dev, generator=g)
        old = torch.rand(R, B, A, device=dev, generator=g) * 0.5 + 0.05
        adv = torch.randn(R, B, A, device=dev, generator=g)
        ppo.ppo_loss_grad_cuda(s.factor_net, flat, x_rows, idx, old, adv, 0.2, 0.01)             # local gradient
        want = flat.grad.clone()
        dist.all_reduce(want, op=dist.ReduceOp.AVG)
        ppo.ppo_loss_grad_cuda(s.factor_net, flat, x_rows, idx, old, adv, 0.2, 0.01, exchange=ex)  # fused exchange
        got = flat.grad.clone()
        worst = max(worst, float((got - want).abs().max() / want.abs().max()))
        both = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(both, got)
        same = same and all(torch.equal(both[0], b) for b in both[1:])
    q.put((rank, worst, same))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_peer_memory_gradient_exchange_matches_nccl_average():
    world, port = min(torch.cuda.device_count(), 8), _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, worst, same in res:
        assert worst <= 2e-6, f"rank {rank}: fused exchange vs ncclAllReduce(AVG): {worst}"     # summation order only
        assert same, "ranks hold different averages (the rank-ordered sum must be bit-identical everywhere)"
