import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from consolver_b200.denoise import preview_from_pairs
dev = torch.device("cuda", 0)
sd = bench.policy_state_dict(0)
s = bench.make_scheduler(dev, sd)
x, pairs = bench.synth_batch(64, 1, dev)
pl = list(pairs.unbind(0))
def run(n):
    for _ in range(n):
        s.set_timesteps(8, device=dev)
        preview_from_pairs(s, x, pl, 3.0)
run(20); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); run(300); torch.cuda.synchronize(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(22); print(st.getvalue()[:4500])
