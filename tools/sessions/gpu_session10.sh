#!/bin/bash
# round-end style validation: smoke, default bench, reference arm, launch list, dpm kernel capture, sanitizer on the new paths
mkdir -p gpurun_out
( time python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
bash tools/launchlist.sh > gpurun_out/launch_summary_r01b.txt 2>&1
cp gpurun_out/launches.csv gpurun_out/launches_r01b.csv
ncu --set full --clock-control none --import-source on -k regex:dpm_step -s 4 -c 1 -o gpurun_out/dpm_B256_r01 \
    python tools/baseline_bench.py --batches 256 --fm-batches "" > gpurun_out/ncu_dpm.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_baselines.py tests/test_gpu_golden.py -m gpu -q -k "amed or fm_baseline or sd_16bit" > gpurun_out/sanitizer_memcheck_r01b.log 2>&1
tail -3 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench_default.json; echo; tail -3 gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_reference.json; echo
head -6 gpurun_out/launch_summary_r01b.txt; tail -4 gpurun_out/sanitizer_memcheck_r01b.log; ls -la gpurun_out/*.ncu-rep
