#!/bin/bash
# round 2, after the PDL-ordering fix and the one-vector launch default: GPU suite, smoke, both bench arms with the
# driver's flags, and fresh ncu evidence of the step kernel (launch list + --set full at B = 64 / 256 / 4096)
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3) > gpurun_out/pytest_gpu_final.log 2>&1
tail -5 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_final2_reference.json 2> gpurun_out/bench_r02_final2_reference.err
(time python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_final2.json") if l.startswith("{")][-1])
r=d["roofline"]
print(d["value"], d["e2e"]["value"], r["frac"], r["chained"]["frac"], r["in_timed_region"], d["clocks"])
print([(p["batch"], p["frac"]) for p in d["roofline_sweep"]], d["fm_flux_preview"]["value"], d["ppo_rollout"]["value"])
print(d["cpu_baseline"]["value"], d["torch_eager_gpu"].get("value"), d["torch_compile_gpu"].get("value"), d["with_denoiser"].get("value"))
print(open("gpurun_out/bench_r02_final2_reference.json").read()[:300])
PY
R=r02b
bash tools/launchlist.sh > gpurun_out/launch_summary_$R.txt 2>&1
cp gpurun_out/launches.csv gpurun_out/launches_$R.csv
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 26 -c 2 -o gpurun_out/step_B64_$R -f \
    python tools/microbench.py --batches 64 --graph 0 --copy 0 --iters 4 > gpurun_out/ncu_full64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 2 -o gpurun_out/step_B256_$R -f \
    python tools/microbench.py --batches 256 --graph 0 --copy 0 --iters 4 > gpurun_out/ncu_full256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/step_B4096_$R -f \
    python tools/microbench.py --batches 4096 --graph 0 --copy 0 --iters 3 > gpurun_out/ncu_full4096.log 2>&1
head -8 gpurun_out/launch_summary_$R.txt
ls -la gpurun_out/*$R.ncu-rep
