"""SURVEY §8(d) "timing the reference CPU path": the torch-CPU oracle port of the 8-step SD preview on the host cores
of the box, best of 5 for B in {1, 16, 64, 256}, with all host threads and with one.  One JSON object per point.
Test/bench infrastructure (it executes oracle/), never part of the product."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (reuses the bench's synthetic workload and oracle preview closure)


def best_of(run, reps=5):
    run()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "")
    except OSError:
        pass
    ncpu = os.cpu_count() or 1
    for B in (1, 16, 64, 256):
        run = bench._oracle_preview_fn(B)
        row = {"B": B, "host_cpus": ncpu, "cpu_model": model, "torch": torch.__version__}
        for threads in (ncpu, 1):
            torch.set_num_threads(threads)
            t = best_of(run, reps=5 if B * threads <= 4096 or threads > 1 else 3)
            key = "all_threads" if threads == ncpu else "one_thread"
            row[key] = {"threads": torch.get_num_threads(), "best_ms_per_preview_batch": round(t * 1e3, 2),
                        "previews_per_s": round(B / t, 1),
                        "algorithmic_gbs": round(bench.TENSORS_PER_PREVIEW * B * 65536 / t / 1e9, 2)}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
