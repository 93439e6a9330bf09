"""Which order does torch.sum(torch.stack([c0, a1, a2]), dim=0) add in on CUDA?  (set_default_coefficients,
scheduler_ppo.py:172.)  Enumerates all 11^3 bin combinations of the production policy and compares with the candidate
fp32 evaluation orders; prints the ones that disagree with the left-to-right order the kernels use."""
import itertools
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from consolver_b200.factor_net import _action_values

av = _action_values("sd", 11, 4, 0, 0)
for dev in ("cpu", "cuda"):
    a0, a1, a2 = (av[i].to(dev) for i in range(3))
    I, J, K = torch.meshgrid(torch.arange(11), torch.arange(11), torch.arange(11), indexing="ij")
    c0 = (a0[I.flatten()] + 1).view(-1, 1, 1, 1)
    v1 = a1[J.flatten()].view(-1, 1, 1, 1)
    v2 = a2[K.flatten()].view(-1, 1, 1, 1)
    ref = (1 - torch.sum(torch.stack([c0, v1, v2]), dim=0)).flatten().cpu().numpy()
    c0n, v1n, v2n = (t.flatten().cpu().numpy().astype(np.float32) for t in (c0, v1, v2))
    one = np.float32(1)
    cands = {"(c0+a1)+a2": one - ((c0n + v1n) + v2n), "c0+(a1+a2)": one - (c0n + (v1n + v2n)),
             "(c0+a2)+a1": one - ((c0n + v2n) + v1n)}
    f64 = (1 - (c0n.astype(np.float64) + v1n + v2n)).astype(np.float32)
    cands["fp64 sum rounded"] = f64
    print(dev, {k: int((v != ref).sum()) for k, v in cands.items()}, "of", ref.size)
    # B = 1 (the stack is [3,1,1,1,1]) may take another reduction path than B = 1331
    mism = 0
    for i, j, k in itertools.islice(itertools.product(range(11), repeat=3), 0, 1331, 7):
        s = torch.stack([(a0[i] + 1).view(1, 1, 1, 1), a1[j].view(1, 1, 1, 1), a2[k].view(1, 1, 1, 1)])
        r = float(1 - torch.sum(s, dim=0))
        want = float(one - ((np.float32(a0[i].item()) + one) + np.float32(a1[j].item()) + np.float32(a2[k].item())))
        mism += r != want
    print(dev, "B=1 mismatches vs left-to-right:", mism)


# ---- the restated ATen order (oracle TorchSemantics.sum0) against torch.sum on this device, m = 2..7 terms ---------------
sys.path.insert(0, "oracle")
import consolver_oracle as orc  # noqa: E402

if torch.cuda.is_available():
    g = torch.Generator().manual_seed(0)
    for m in range(2, 8):
        for B in (1, 2, 5, 64):
            bad = 0
            for trial in range(200):
                vals = (torch.randn(m, B, 1, 1, 1, generator=g) * 1.3).float()
                ref = torch.sum(vals.cuda(), dim=0).cpu()
                got = orc.CUDA.sum0(vals)
                bad += int(not torch.equal(ref, got))
            print(f"m={m} B={B}: restated order differs from torch.sum on cuda in {bad}/200 trials")
