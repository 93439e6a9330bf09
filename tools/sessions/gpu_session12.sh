#!/bin/bash
# final default bench line of the round for profiles/
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_final.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["clocks_e2e"])
print(d["fm_flux_preview"]); print(d["with_denoiser"]["value"], d["with_denoiser"]["preview_latency"])
print(d["cpu_baseline"]["value"], d["gpu_launches"])
PY
tail -2 gpurun_out/bench_final.err
