#!/bin/bash
# 2-GPU sanity after the round's later changes: multi-GPU test, default-flag bench and reference arm under torchrun
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 1000 --warmup 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
cat gpurun_out/pytest_multi.log
python - <<'PY'
import json
for f in ('gpurun_out/bench_n2.json','gpurun_out/bench_ref_n2.json'):
    ls=[l for l in open(f) if l.startswith('{')]
    print(f, len(ls))
    if ls:
        d=json.loads(ls[-1]); print(d.get('impl'), d['n_gpus'], d['value'], d['e2e']['value'], d.get('clocks'), d.get('gpu_launches'))
PY
tail -3 gpurun_out/bench_n2.err
