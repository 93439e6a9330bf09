"""Solver-step microbench (BASELINE config 2): fused SD step kernel over a batch sweep, through the C ABI,
CUDA-event timed, rotating buffer sets so the working set exceeds L2.  Prints one JSON object per point."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from consolver_b200 import _lib  # noqa: E402

L2_BYTES = 126 * 2 ** 20


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def bench_copy(nbytes_total, iters=50):
    """torch b.copy_(a) moving the same total bytes (half read, half written): the 'copy at this size' yardstick"""
    n = nbytes_total // 2 // 4
    nsets = max(2, min(64, -(-3 * L2_BYTES // nbytes_total)))
    bufs = [(torch.randn(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(nsets)]
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for a, b in bufs:
            b.copy_(a)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for i in range(iters):
            a, b = bufs[i % nsets]
            b.copy_(a)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / iters * 1e3)
    ts.sort()
    return round(ts[2], 3), round(2 * n * 4 / ts[2] / 1e3, 1)


def bench_sd(B, n_hist=4, pair=True, N=4 * 64 * 64, dtype=torch.float32, iters=200, flags=0, min_bytes=3 * L2_BYTES,
             order_dim=4, graph=False):
    lib = _lib.load()
    es = torch.empty((), dtype=dtype).element_size()
    tensors = (n_hist - 1) + 2 + (2 if pair else 1) + (1 if pair else 0)   # reads + writes (x', slot)
    bytes_per_launch = tensors * B * N * es
    nsets = max(2, min(64, -(-min_bytes // bytes_per_launch)))
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda: torch.randn(B, N, device="cuda", generator=g).to(dtype)  # noqa: E731
    sets = []
    for _ in range(nsets):
        sets.append(dict(e0=mk(), cond=mk() if pair else None, x=mk(), hist=[mk() for _ in range(n_hist - 1)],
                         out=torch.empty(B, N, device="cuda", dtype=dtype),
                         slot=torch.empty(B, N, device="cuda", dtype=dtype) if pair else None))
    coef = torch.randn(B, order_dim + 2, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    code = _lib.dtype_code(dtype)

    def launch(s):
        nonlocal stream
        rc = lib.consolver_step_sd(code, s["e0"].data_ptr(), s["cond"].data_ptr() if pair else None, 3.0,
                                   s["slot"].data_ptr() if pair else None,
                                   _lib.ptr_array([h.data_ptr() for h in s["hist"]]), n_hist, s["x"].data_ptr(),
                                   s["out"].data_ptr(), None, 0, coef.data_ptr(), order_dim + 2, order_dim,
                                   0.8378, 0.5460, 0.9151, 0.4033, flags, B, N, stream)
        assert rc == 0, rc

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        stream = side.cuda_stream
        for i in range(max(3, nsets)):
            launch(sets[i % nsets])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    cg = None
    if graph:
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            stream = torch.cuda.current_stream().cuda_stream
            for i in range(iters):
                launch(sets[i % nsets])
        cg.replay()
        torch.cuda.synchronize()
    times = []
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if cg is not None:
            cg.replay()
        else:
            for i in range(iters):
                launch(sets[i % nsets])
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) / iters * 1e3)  # us per launch
    times.sort()
    us = times[len(times) // 2]
    return dict(B=B, n_hist=n_hist, pair=pair, dtype=str(dtype).split(".")[-1], us_median=round(us, 3),
                us_best=round(times[0], 3), bytes=bytes_per_launch, gbs=round(bytes_per_launch / us / 1e3, 1),
                gbs_best=round(bytes_per_launch / times[0] / 1e3, 1), nsets=nsets)


def torch_reference_step(pair, x, hist, coef, guidance, sc):
    """The reference's op sequence for one steady-state step, as stock torch ops on the GPU: CFG combine
    (denoise_ppo.py:97-100), the dead stack copy (scheduler_ppo.py:222), the python-sum combine (:272) and
    _get_prev_sample (:306-332).  17 full-size elementwise kernels + one stack — the bar the fused kernel replaces."""
    u, c = pair.chunk(2)
    eps = u + guidance * (c - u)
    ets = [eps] + hist
    torch.stack(ets, dim=1)
    eff = sum(cj.view(-1, 1) * e for cj, e in zip(coef.unbind(1), ets))
    sa_t, sb_t, sa_p, sb_p = sc
    x0 = (x - sb_t * eff) / sa_t
    return sa_p * x0 + sb_p * eff, eps


def bench_torch_reference(B, mode="eager", N=4 * 64 * 64, iters=30, min_bytes=3 * L2_BYTES):
    per = 8 * B * N * 4
    nsets = max(2, min(16, -(-min_bytes // per)))
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    sets = [dict(pair=mk(2 * B, N), x=mk(B, N), hist=[mk(B, N) for _ in range(3)]) for _ in range(nsets)]
    coef = mk(B, 4)
    sc = [torch.tensor(v) for v in (0.8378, 0.5460, 0.9151, 0.4033)]     # 0-d CPU scalars, as in the reference
    fn = torch_reference_step
    if mode == "compile":
        fn = torch.compile(torch_reference_step, dynamic=False)
    for i in range(3):
        fn(sets[i % nsets]["pair"], sets[i % nsets]["x"], sets[i % nsets]["hist"], coef, 3.0, sc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            s = sets[i % nsets]
            fn(s["pair"], s["x"], s["hist"], coef, 3.0, sc)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / iters * 1e3)
    ts.sort()
    return dict(B=B, impl=f"torch-{mode} reference op sequence on the GPU", us_median=round(ts[2], 2),
                gbs_algorithmic=round(per / ts[2] / 1e3, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,4,16,64,256,1024,4096")
    ap.add_argument("--threads", default="0")
    ap.add_argument("--unroll", default="0")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--pdl", type=int, default=0)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--copy", type=int, default=1)
    ap.add_argument("--torch-ref", default="", help="comma list of eager,compile: also time the reference op sequence")
    a = ap.parse_args()
    peak, src = load_peak()
    lib = _lib.load()
    for th in [int(v) for v in a.threads.split(",")]:
        for un in [int(v) for v in a.unroll.split(",")]:
            assert lib.consolver_set_step_launch(th, un) == 0
            for B in [int(v) for v in a.batches.split(",")]:
                r = bench_sd(B, iters=a.iters, flags=8 if a.pdl else 0, graph=bool(a.graph))
                if a.copy:
                    r["copy_us"], r["copy_gbs"] = bench_copy(r["bytes"])
                r.update(graph=a.graph, threads=th, unroll=un, frac=round(r["gbs"] / peak, 3), peak=peak, peak_src=src)
                print(json.dumps(r), flush=True)
    lib.consolver_set_step_launch(0, 0)
    for mode in [m for m in a.torch_ref.split(",") if m]:
        for B in [int(v) for v in a.batches.split(",")]:
            try:
                r = bench_torch_reference(B, mode)
                r.update(frac=round(r["gbs_algorithmic"] / peak, 3))
                print(json.dumps(r), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"B": B, "impl": f"torch-{mode}", "error": repr(e)[:200]}), flush=True)


if __name__ == "__main__":
    main()
