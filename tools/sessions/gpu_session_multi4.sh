#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $NG --steps 1000 --warmup 20 --cpu-budget 3 > gpurun_out/scale_n$NG.json 2> gpurun_out/scale_n$NG.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29622 tools/ppo_bench.py > gpurun_out/ppo_n$NG.json 2> gpurun_out/ppo_n$NG.err
tail -2 gpurun_out/pytest_gpu.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_n4.json')):
    ls=[l for l in open(f) if l.startswith('{')]
    if ls:
        d=json.loads(ls[-1]); print(f, d['n_gpus'], d['value'], d['e2e']['value'], d.get('with_denoiser',{}).get('value'), d['clocks'])
for f in sorted(glob.glob('gpurun_out/ppo_n4.json')):
    ls=[l for l in open(f) if l.startswith('{')]
    if ls: print(f, ls[-1][:260])
PY
tail -3 gpurun_out/scale_n$NG.err
