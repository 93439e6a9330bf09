#!/bin/bash
# full GPU suite after the 16-bit parity work
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
