#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --eager --steps 50 --warmup 5 --cpu-budget 1 > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
python tools/microbench.py --batches 64,256 --threads 128,256,512 --unroll 1,2 --copy 0 > gpurun_out/micro_sweep.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy -s 12 -c 4 -o gpurun_out/policy_r01 \
    python bench.py --steps 2 --warmup 3 --no-extras --eager > gpurun_out/ncu_policy.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench_ours.json; tail -2 gpurun_out/bench_ours.err
