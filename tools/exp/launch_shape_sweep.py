"""Launch-shape sweep of the step kernel (threads per CTA x vectors per thread, consolver_set_step_launch) for the FM / FLUX
shape and the SD shape at the bench batches; also the whole FM preview leg under each setting.
python tools/exp/launch_shape_sweep.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from consolver_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
out = []
for threads, unroll in ((0, 0), (256, 1), (256, 2), (128, 2), (512, 1), (512, 2), (128, 1)):
    assert lib.consolver_set_step_launch(threads, unroll) == 0
    row = dict(threads=threads, unroll=unroll)
    for B in (8, 16, 64):
        us, bytes_ = bench.time_fm_kernel(B, dev)
        row[f"fm_B{B}_us"] = round(us, 3)
        row[f"fm_B{B}_frac"] = round(bytes_ / us / 1e3 / 6550.7, 4)
    for B in (64, 256):
        us, bytes_, _ = bench.time_step_kernel(B, 4, dev)
        row[f"sd_B{B}_us"] = round(us, 3)
        row[f"sd_B{B}_frac"] = round(bytes_ / us / 1e3 / 6550.7, 4)
    if (threads, unroll) in ((0, 0), (256, 2), (512, 1), (512, 2)):
        row["fm_preview"] = bench.fm_preview_throughput(dev)["value"]
    print(json.dumps(row), flush=True)
    out.append(row)
lib.consolver_set_step_launch(0, 0)
