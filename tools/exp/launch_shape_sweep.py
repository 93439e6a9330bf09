"""Finer launch-shape sweep (threads x vectors per thread) over the batch range, SD fp32 n_hist=4 CFG pair and FM bf16.
Two passes per setting (median-of-7 each).  python tools/exp/launch_shape_sweep.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from consolver_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
for rep in range(2):
    for threads, unroll in ((0, 0), (128, 1), (256, 1), (128, 2), (256, 2), (64, 1), (64, 2)):
        assert lib.consolver_set_step_launch(threads, unroll) == 0
        row = dict(rep=rep, threads=threads, unroll=unroll)
        for B in (32, 64, 128, 256, 512, 1024, 4096):
            us, bytes_, _ = bench.time_step_kernel(B, 4, dev, iters=64 if B <= 512 else 16)
            row[f"sd{B}"] = round(bytes_ / us / 1e3 / 6550.7, 4)
        for B in (4, 16, 64, 256):
            us, bytes_ = bench.time_fm_kernel(B, dev)
            row[f"fm{B}"] = round(bytes_ / us / 1e3 / 6550.7, 4)
        print(json.dumps(row), flush=True)
lib.consolver_set_step_launch(0, 0)
