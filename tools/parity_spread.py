"""Measures — on a GPU box — how far the policy kernels' softmax tables are from the reference's, for EVERY fixture,
and how far the reference is from itself (CPU/MKL run vs CUDA/cuBLAS run of the same weights, same inputs) and from the
fp64 evaluation of the same network.  The tolerances in tests/test_gpu_golden.py and tests/test_gpu_cuda_reference.py
are set to 2x the worst numbers this prints; nothing there is a round guess.

    python tools/parity_spread.py [--md profiles/parity_spread_r02.md] [--json profiles/parity_spread_r02.json]

Columns: max |dp| (absolute), max |dp|/p over entries with p >= 1e-6, max |d log(p+1e-9)| over the SAMPLED entries (what
PPO consumes, train_ppo.py:410-411), sampled-index mismatches of the kernel's own draw on the fixture's Exp(1) values.
"""
import argparse
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import consolver_b200 as cb  # noqa: E402
from golden_io import Golden, names  # noqa: E402

DT = {None: None, "float16": torch.float16, "bfloat16": torch.bfloat16}


def fp64_tables(g, rows, variant, feats=None):
    """the policy evaluated in float64 from the fp32 master weights (no autocast): the 'truth' both builds approximate"""
    sd = {k: v.double() for k, v in g.state_dict.items()}
    x = rows.double()
    x = x / 999.0 if variant == "sd" else x
    if feats is not None:
        x = torch.cat([x, feats.double()], -1)
    h = torch.relu(x @ sd["mlp.0.weight"].t() + sd["mlp.0.bias"])
    h = torch.relu(h @ sd["mlp.2.weight"].t() + sd["mlp.2.bias"])
    lg = h @ sd["mlp.4.weight"].t() + sd["mlp.4.bias"]
    A, K = sd["action_values"].shape
    lg = lg.view(-1, A, K)
    if variant == "fm":
        lg = lg / 0.01
    return torch.softmax(lg, -1)


def spread(a, b):
    a, b = a.double(), b.double()
    d = (a - b).abs()
    big = b >= 1e-6
    rel = (d[big] / b[big]).max().item() if big.any() else 0.0
    return d.max().item(), rel


def run_fixture(name):
    g = Golden(name)
    m = g.meta
    kind = m["kind"]
    if kind == "sd":
        s = cb.PPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
    else:
        s = cb.FMPPOScheduler(factor_net_kwargs=dict(m["factor_net_kwargs"]), **m["config"])
    s.reference_device = m.get("device", "cpu")
    s.factor_net.load_state_dict(g.state_dict)
    if m.get("policy_dtype"):
        s.factor_net.to("cuda", dtype=DT[m["policy_dtype"]])
    else:
        s.factor_net.cuda()
    if kind == "sd":
        s.set_timesteps(m["n"], device="cuda")
    else:
        s.set_timesteps(m["n"], device="cuda", sigmas=np.linspace(1.0, 1 / m["n"], m["n"]), mu=m["mu"])
        if m["use_begin_index"]:
            s.set_begin_index(0)
    s.replay = {"q": [g[f"q_{i}"].cuda() for i in range(m["n"])]}
    ctx = torch.autocast("cuda", DT[m["autocast"]]) if m.get("autocast") else contextlib.nullcontext()
    conv = m["config"].get("use_conv", False)
    out = dict(name=name, kind=kind, device=m.get("device", "cpu"), autocast=m.get("autocast"),
               policy_dtype=m.get("policy_dtype"), model_dtype=m["dtype"], use_conv=conv,
               dp_abs=0.0, dp_rel=0.0, dlogp=0.0, idx_mismatch=0, idx_total=0, latent_mismatch_steps=0,
               truth_kernel_rel=0.0, truth_ref_rel=0.0)
    x = g["x_T"].cuda()
    with ctx:
        for i, t in enumerate(s.timesteps):
            mo = g[f"eps_{i}"] if kind == "sd" else g[f"v_{i}"]
            x_in = g[f"prev_{i - 1}"].cuda() if i else x                 # stay on the reference trajectory
            res = s.step(mo.cuda(), t, x_in, return_dict=False)
            lp = s.last_policy()
            tab = lp["probs_table"].cpu()
            ref = g[f"probs_full_{i}"].float() if conv else g[f"probs_full_{i}"][0].float()
            a, r = spread(tab, ref)
            out["dp_abs"], out["dp_rel"] = max(out["dp_abs"], a), max(out["dp_rel"], r)
            ref_lp = torch.log(g[f"probs_{i}"].float() + 1e-9)
            same = lp["idx"].cpu() == g[f"idx_{i}"]
            if same.any():
                out["dlogp"] = max(out["dlogp"], (lp["logp"].cpu() - ref_lp)[same].abs().max().item())
            out["idx_mismatch"] += int((~same).sum())
            out["idx_total"] += same.numel()
            out["latent_mismatch_steps"] += int(not (res[0].dtype == g[f"prev_{i}"].dtype and
                                                     torch.equal(res[0].cpu(), g[f"prev_{i}"])))
            if not conv and not m.get("autocast") and not m.get("policy_dtype"):
                truth = fp64_tables(g, g[f"condx_{i}"][:1].float(), kind)[0]
                out["truth_kernel_rel"] = max(out["truth_kernel_rel"], spread(tab, truth)[1])
                out["truth_ref_rel"] = max(out["truth_ref_rel"], spread(ref, truth)[1])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", default=os.path.join(ROOT, "gpurun_out", "parity_spread.md"))
    ap.add_argument("--json", default=os.path.join(ROOT, "gpurun_out", "parity_spread.json"))
    args = ap.parse_args()
    rows = []
    for prefix in ("sd_", "sd16_", "fm_", "cuda_"):
        for name in names(prefix):
            try:
                rows.append(run_fixture(name))
            except Exception as e:  # noqa: BLE001
                rows.append(dict(name=name, error=repr(e)[:300]))
            print(json.dumps(rows[-1]), flush=True)
    # the reference against itself: CPU-made vs CUDA-made fixtures with the same seed / weights / inputs
    self_spread = []
    for cuda_name in names("cuda_"):
        cpu_name = cuda_name[len("cuda_"):].replace("autocastlayout", "autocast")
        if cpu_name in names(""):
            gc, gg = Golden(cpu_name), Golden(cuda_name)
            worst = (0.0, 0.0)
            for i in range(gc.meta["n"]):
                a, r = spread(gc[f"probs_full_{i}"].float(), gg[f"probs_full_{i}"].float())
                worst = (max(worst[0], a), max(worst[1], r))
            self_spread.append(dict(cpu=cpu_name, cuda=cuda_name, dp_abs=worst[0], dp_rel=worst[1]))
    os.makedirs(os.path.dirname(args.json), exist_ok=True)
    with open(args.json, "w") as f:
        json.dump(dict(gpu=torch.cuda.get_device_name(0), torch=torch.__version__, fixtures=rows,
                       reference_cpu_vs_cuda=self_spread), f, indent=1)
    with open(args.md, "w") as f:
        f.write("# Policy-table parity: measured spreads (tools/parity_spread.py, " + torch.cuda.get_device_name(0) + ")\n\n")
        f.write("| fixture | made on | autocast | max abs dp | max rel dp (p>=1e-6) | max dlogp (sampled) | own-draw idx mismatch | "
                "latent steps differing | kernel vs fp64 rel | reference vs fp64 rel |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            if "error" in r:
                f.write(f"| {r['name']} | ERROR {r['error']} |\n")
                continue
            f.write(f"| {r['name']} | {r['device']} | {r['autocast'] or r['policy_dtype'] or '-'} | {r['dp_abs']:.2e} | "
                    f"{r['dp_rel']:.2e} | {r['dlogp']:.2e} | {r['idx_mismatch']}/{r['idx_total']} | "
                    f"{r['latent_mismatch_steps']} | {r['truth_kernel_rel']:.2e} | {r['truth_ref_rel']:.2e} |\n")
        f.write("\n## The reference against itself: CPU (MKL) run vs CUDA (cuBLAS) run, same weights and inputs\n\n"
                "| CPU fixture | CUDA fixture | max abs dp | max rel dp |\n|---|---|---|---|\n")
        for r in self_spread:
            f.write(f"| {r['cpu']} | {r['cuda']} | {r['dp_abs']:.2e} | {r['dp_rel']:.2e} |\n")
    print("wrote", args.md)


if __name__ == "__main__":
    main()
