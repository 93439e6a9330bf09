"""Baseline-solver step kernels on the same yardstick as tools/microbench.py: the AMED / DPM-Solver++ second-order
step with a CFG pair (reads u, c, x, m1; writes x', m0 = 6 latent-sized tensors) and the flow-matching heun second
stage (reads v, kept v, kept x; writes x' = 4 tensors), graph-captured, rotating buffer sets larger than L2,
CUDA-event timed.  One JSON object per point."""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from consolver_b200 import _lib  # noqa: E402

L2_BYTES = 126 * 2 ** 20


def _time(launch, nsets, iters):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(max(3, nsets)):
            launch(i % nsets, side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        st = torch.cuda.current_stream().cuda_stream
        for i in range(iters):
            launch(i % nsets, st)
    cg.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cg.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / iters * 1e3)
    ts.sort()
    return ts[2], ts[0]


def bench_dpm(B, N=4 * 64 * 64, dtype=torch.float32, iters=200):
    lib = _lib.load()
    es = torch.empty((), dtype=dtype).element_size()
    nbytes = 6 * B * N * es
    nsets = max(2, min(64, -(-3 * L2_BYTES // nbytes)))
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda: torch.randn(B, N, device="cuda", generator=g).to(dtype)  # noqa: E731
    sets = [dict(u=mk(), c=mk(), x=mk(), m1=mk(), out=torch.empty(B, N, device="cuda", dtype=dtype),
                 slot=torch.empty(B, N, device="cuda", dtype=dtype)) for _ in range(nsets)]
    code = _lib.dtype_code(dtype)
    upd = _lib.DpmUpdate(cx=0.71, a0=-0.21, a1=-0.105, rinv=1.3)
    upd_ref = ctypes.byref(upd)

    def launch(k, st):
        s = sets[k]
        rc = lib.consolver_step_dpm(code, code, s["u"].data_ptr(), s["c"].data_ptr(), 7.5, s["slot"].data_ptr(),
                                    s["m1"].data_ptr(), None, s["x"].data_ptr(), s["out"].data_ptr(), None, 0,
                                    _lib.DPM_CONVERT_DIV, 0.83, 0.55, upd_ref, B, N, st)
        assert rc == 0, rc

    med, best = _time(launch, nsets, iters)
    return dict(kernel="dpm_step (2nd order, CFG pair)", B=B, dtype=str(dtype).split(".")[-1], bytes=nbytes,
                us_median=round(med, 3), gbs=round(nbytes / med / 1e3, 1), gbs_best=round(nbytes / best / 1e3, 1))


def bench_fm_heun2(B, N=4096 * 64, dtype=torch.bfloat16, iters=100):
    lib = _lib.load()
    es = torch.empty((), dtype=dtype).element_size()
    nbytes = 4 * B * N * es
    nsets = max(2, min(64, -(-3 * L2_BYTES // nbytes)))
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda: torch.randn(B, N, device="cuda", generator=g).to(dtype)  # noqa: E731
    sets = [dict(v=mk(), v1=mk(), x=mk(), out=torch.empty(B, N, device="cuda", dtype=dtype)) for _ in range(nsets)]
    ones = torch.ones(B, 4, device="cuda")
    code = _lib.dtype_code(dtype)

    def launch(k, st):
        s = sets[k]
        rc = lib.consolver_step_fm(code, code, s["v"].data_ptr(), None, _lib.ptr_array([s["v1"].data_ptr()]), 2,
                                   s["x"].data_ptr(), s["out"].data_ptr(), None, 0, ones.data_ptr(), 4, 2, -0.05,
                                   _lib.FLAG_LOWP_COMBINE, B, N, st)
        assert rc == 0, rc

    med, best = _time(launch, nsets, iters)
    return dict(kernel="fm_step heun 2nd stage (LOWP_COMBINE)", B=B, dtype=str(dtype).split(".")[-1], bytes=nbytes,
                us_median=round(med, 3), gbs=round(nbytes / med / 1e3, 1), gbs_best=round(nbytes / best / 1e3, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="64,256,1024,4096")
    ap.add_argument("--fm-batches", default="8,64,512")
    a = ap.parse_args()
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    for B in [int(b) for b in a.batches.split(",") if b]:
        r = bench_dpm(B)
        r["frac_of_copy_peak"] = round(r["gbs"] / peak, 3)
        print(json.dumps(r), flush=True)
    for B in [int(b) for b in a.fm_batches.split(",") if b]:
        r = bench_fm_heun2(B)
        r["frac_of_copy_peak"] = round(r["gbs"] / peak, 3)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
