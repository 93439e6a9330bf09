#!/bin/bash
# mixed-precision SD step tests, TMA bulk-copy variant of the step kernel, CPU reference sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_scheduler_edges.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_mixed.log
timeout 240 tools/exp/step_variants 64 256 1024 4096 > gpurun_out/step_variants_tma.txt 2>&1
echo "variants rc=$?" >> gpurun_out/step_variants_tma.txt
timeout 300 python tools/cpu_sweep.py > gpurun_out/cpu_sweep.jsonl 2>&1
tail -5 gpurun_out/pytest_mixed.log; grep -i "tma\|product\|rc=\|error" gpurun_out/step_variants_tma.txt; cat gpurun_out/cpu_sweep.jsonl | cut -c1-400
