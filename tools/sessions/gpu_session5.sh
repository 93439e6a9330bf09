#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python tools/microbench.py --batches 64,256,1024,4096 --torch-ref eager,compile > gpurun_out/micro_vs_torch.jsonl 2> gpurun_out/micro_vs_torch.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/micro_vs_torch.jsonl | cut -c1-220; tail -3 gpurun_out/micro_vs_torch.err
