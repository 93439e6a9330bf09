#!/bin/bash
# one gpurun call: parity tests, bench (both arms), microbench, ncu launch list + full capture of the step kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --eager --steps 50 --warmup 5 --cpu-budget 1 > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/microbench.py --batches 1,16,64,128,256,512,1024,4096 > gpurun_out/micro_graph.jsonl 2>&1
python tools/microbench.py --batches 64,256 --pdl 1 --copy 0 > gpurun_out/micro_pdl.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 2 -o gpurun_out/step_B256_r01 \
    python tools/microbench.py --batches 256 --graph 0 --copy 0 --iters 4 > gpurun_out/ncu_full256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 26 -c 2 -o gpurun_out/step_B64_r01 \
    python tools/microbench.py --batches 64 --graph 0 --copy 0 --iters 4 > gpurun_out/ncu_full64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 4 -c 2 -o gpurun_out/policy_r01 \
    python bench.py --steps 2 --warmup 3 --no-extras --eager > gpurun_out/ncu_policy.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ours.json | cut -c1-600
