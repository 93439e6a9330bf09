#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python tools/microbench.py --batches 64,256,1024,4096 > gpurun_out/micro_graph.jsonl 2>&1
python bench.py --no-denoiser > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -4 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
for l in open('gpurun_out/micro_graph.jsonl'):
    if l.startswith('{'):
        r=json.loads(l); print(r['B'], r['us_median'], r['gbs'], r['frac'], r.get('copy_gbs'))
d=json.loads([l for l in open('gpurun_out/bench_ours.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline'], d['roofline_sweep'])
PY
