#!/bin/bash
# round-end style scaling run: N = 1, 2, 4, 8 (as many as the box has), default flags
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    EXTRA=""
    if [ $N -eq 2 ] || [ $N -eq 4 ]; then EXTRA="--no-denoiser --cpu-budget 2"; fi
    if [ $N -eq 1 ]; then
      python bench.py $EXTRA > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N $EXTRA > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err
    fi
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29733 tools/ppo_bench.py > gpurun_out/final_ppo_n$NG.json 2> gpurun_out/final_ppo_n$NG.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final_n*.json')):
    ls=[l for l in open(f) if l.startswith('{')]
    if ls:
        d=json.loads(ls[-1]); print(f, d['n_gpus'], d['value'], d['e2e']['value'], d.get('with_denoiser',{}).get('value'), d['solver_loop'], d['clocks']['reasons'])
for f in sorted(glob.glob('gpurun_out/final_ppo_n*.json')):
    ls=[l for l in open(f) if l.startswith('{')]
    if ls: print(f, ls[-1][:200])
PY
