#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
NG=$(nvidia-smi -L | wc -l)
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 100 --warmup 10 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_n$NG.json 2> gpurun_out/bench_ref_n$NG.err
python tools/ppo_bench.py > gpurun_out/ppo_n1.json 2> gpurun_out/ppo_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 tools/ppo_bench.py > gpurun_out/ppo_n$NG.json 2> gpurun_out/ppo_n$NG.err
tail -3 gpurun_out/pytest_gpu.log; for f in gpurun_out/bench_n*.json gpurun_out/ppo_n*.json gpurun_out/bench_ref_n*.json; do echo $f; cut -c1-250 $f; done; tail -3 gpurun_out/*.err | cut -c1-300
