"""Host-time profile of one PPO iteration of bench.ppo_rollout_leg (rollout + reward + update), single GPU."""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import consolver_b200 as cb  # noqa: E402
from consolver_b200 import ppo  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(1234)
s = cb.PPOScheduler(factor_net_kwargs=dict(bench.FN_KW), **bench.SD_CFG)
with torch.no_grad():
    s.factor_net.mlp[4].weight.normal_(0, 0.05)
s.factor_net.to(dev)
flat = ppo.FlatParams(s.factor_net)
opt = torch.optim.AdamW(s.factor_net.parameters(), lr=1e-4)
g = torch.Generator(device=dev).manual_seed(0)
w = torch.randn(4, 4, device=dev, generator=g) * 0.3
den = lambda x, t, i: torch.einsum("oc,bchw->bohw", w, x)  # noqa: E731
noise = torch.randn(*bench.SHAPE, device=dev, generator=g)
target = torch.randn(*bench.SHAPE, device=dev, generator=g)
batch = 80


def rollout(it):
    n = ppo.shared_step_count(it, seed=0)
    return ppo.rollout_sd(s, den, noise, batch, bench.GUIDANCE, n), n


def update(lat, rec):
    r = ppo.latent_mse_reward(lat, target.unsqueeze(0).expand_as(lat))
    return ppo.ppo_update(s.factor_net, flat, opt, rec, r, ppo_epochs=2, entropy_coef=0.01)


for it in range(5):
    (lat, rec), n = rollout(it)
    update(lat, rec)
torch.cuda.synchronize()
tr = tu = 0.0
steps = 0
for it in range(5, 105):
    t0 = time.perf_counter()
    (lat, rec), n = rollout(it)
    t1 = time.perf_counter()
    update(lat, rec)
    t2 = time.perf_counter()
    tr += t1 - t0
    tu += t2 - t1
    steps += n
print(f"per iteration: rollout {tr * 10:.3f} ms host ({steps / 100:.1f} steps, {tr / steps * 1e6:.1f} us/step), "
      f"update {tu * 10:.3f} ms (incl. its sync)")
pr = cProfile.Profile()
pr.enable()
for it in range(105, 155):
    (lat, rec), n = rollout(it)
    update(lat, rec)
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue()[:5000])
