"""GPU: baseline solvers (DDIM, Adams-Bashforth multistep, FM Euler) through the same fused kernels vs closed forms."""
import numpy as np
import pytest
import torch

import consolver_oracle as orc

pytestmark = pytest.mark.gpu

SD = dict(beta_end=0.012, beta_schedule="scaled_linear", beta_start=0.00085, steps_offset=1, timestep_spacing="trailing")


def test_ddim_and_multistep_match_closed_form():
    from consolver_b200 import baselines

    g = torch.Generator().manual_seed(0)
    B, shape, n = 3, (4, 16, 16), 6
    x0 = torch.randn(B, *shape, generator=g)
    pairs = [torch.randn(2 * B, *shape, generator=g) for _ in range(n)]
    ac = orc.sd_alphas_cumprod(orc.sd_betas(1000, 0.00085, 0.012, "scaled_linear"))
    for order, solver in ((1, baselines.ddim_solver(**SD)), (4, baselines.multistep_solver(4, **SD)),
                          (2, baselines.multistep_solver(2, **SD))):
        solver.set_timesteps(n, device="cuda")
        x_gpu, x_cpu, hist = x0.cuda(), x0, []
        for i, t in enumerate(solver.timesteps.tolist()):
            out = solver.step_cfg(pairs[i].cuda(), t, x_gpu, 3.0)
            assert out[1] is None and out[2] is None
            x_gpu = out[0]
            u, c = pairs[i].chunk(2)
            hist = ([orc.cfg_combine(u, c, 3.0)] + hist)[:order]
            w = baselines.ADAMS_BASHFORTH[min(len(hist), order)]
            coef = None if len(hist) == 1 else [torch.full((B,), float(np.float32(v))) for v in w] + \
                [torch.zeros(B)] * (len(hist) - len(w))
            eff, _ = orc.combine_history(hist, coef, [], x_cpu)
            x_cpu = orc.ddim_update(x_cpu, eff, orc.ddim_scalars(ac, t, orc.sd_prev_timestep(t, n)))
            assert torch.equal(x_gpu.cpu(), x_cpu), f"order {order} step {i}"


def test_flow_euler_matches_reference_formula():
    from consolver_b200 import baselines

    g = torch.Generator().manual_seed(1)
    B, shape, n = 2, (64, 16), 5
    s = baselines.flow_euler_solver(shift=3.0)
    s.set_timesteps(n, device="cuda")
    s.set_begin_index(0)
    x = torch.randn(B, *shape, generator=g).bfloat16()
    x_gpu, x_cpu = x.cuda(), x
    sig = s.sigmas.cpu()
    for i, t in enumerate(s.timesteps):
        v = torch.randn(B, *shape, generator=g).bfloat16()
        x_gpu = s.step(v.cuda(), t, x_gpu, return_dict=False)[0]
        x_cpu = (x_cpu.float() + (sig[i + 1] - sig[i]) * v).to(torch.bfloat16)      # edit_ppo/scheduler_fm.py:405-410
        assert torch.equal(x_gpu.cpu(), x_cpu), f"step {i}"
