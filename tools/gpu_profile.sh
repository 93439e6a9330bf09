#!/bin/bash
# ncu evidence for profiles/: launch list of the bench step + full captures of the step and policy kernels
mkdir -p gpurun_out
R=${1:-r01}
bash tools/launchlist.sh > gpurun_out/launch_summary_$R.txt 2>&1
cp gpurun_out/launches.csv gpurun_out/launches_$R.csv
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 2 -o gpurun_out/step_B256_$R \
    python tools/microbench.py --batches 256 --graph 0 --copy 0 --iters 4 > gpurun_out/ncu_full256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 26 -c 2 -o gpurun_out/step_B64_$R \
    python tools/microbench.py --batches 64 --graph 0 --copy 0 --iters 4 > gpurun_out/ncu_full64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/step_B4096_$R \
    python tools/microbench.py --batches 4096 --graph 0 --copy 0 --iters 3 > gpurun_out/ncu_full4096.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy -s 12 -c 4 -o gpurun_out/policy_$R \
    python bench.py --steps 2 --warmup 3 --no-extras --eager > gpurun_out/ncu_policy.log 2>&1
cat gpurun_out/launch_summary_$R.txt | head -8
ls -la gpurun_out/*.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:ppo_ -s 4 -c 2 -o gpurun_out/ppo_$R \
    python tools/ppo_bench.py --iters 3 > gpurun_out/ncu_ppo.log 2>&1
ls -la gpurun_out/ppo_$R.ncu-rep
