#!/bin/bash
# baseline solvers: parity tests + kernel bandwidth
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_baselines.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_baselines.log
timeout 300 python tools/baseline_bench.py > gpurun_out/baseline_bench.jsonl 2>&1
tail -8 gpurun_out/pytest_baselines.log; cat gpurun_out/baseline_bench.jsonl
