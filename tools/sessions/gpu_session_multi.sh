#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    if [ $N -eq 1 ]; then
      python bench.py --steps 1000 --warmup 20 --no-denoiser --cpu-budget 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 1000 --warmup 20 --no-denoiser --cpu-budget 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    fi
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29713 tools/ppo_bench.py > gpurun_out/ppo_n$NG.json 2> gpurun_out/ppo_n$NG.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_n*.json')):
    ls=[l for l in open(f) if l.startswith('{')]
    if ls:
        d=json.loads(ls[-1]); print(f, d['n_gpus'], d['value'], d['e2e']['value'], d['clocks'])
for f in sorted(glob.glob('gpurun_out/ppo_n*.json')):
    ls=[l for l in open(f) if l.startswith('{')]
    if ls: print(f, ls[-1][:300])
PY
