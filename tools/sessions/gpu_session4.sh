#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -4 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_ours.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('with_denoiser'), d.get('roofline_fm_flux_bf16'))
PY
tail -3 gpurun_out/bench_ours.err
