#!/bin/bash
# Round-2 eight-GPU records: BASELINE config 3 (global batch 8192, carried on the 8-GPU bench line) and config 5
# (ViT-L/14, train.fit with the slow loop every 100 steps).  usage: tools/r2_n8.sh N
N=${1:-8}
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "8" ]; then
  timeout 600 $T --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
  tail -2 $O/r2_bench_n8.err | cut -c1-300
fi
timeout 600 $T --master-port 29542 tools/fit_bench.py --config vit_l14 --steps 210 --interval 100 --overlap 20 --json $O/r2_c5_n${N}_ovl.json > $O/r2_c5_n${N}_ovl.log 2>&1
tail -1 $O/r2_c5_n${N}_ovl.log | cut -c1-600
timeout 600 $T --master-port 29543 tools/fit_bench.py --config vit_l14 --steps 210 --interval 100 --overlap 0 --json $O/r2_c5_n${N}_sync.json > $O/r2_c5_n${N}_sync.log 2>&1
tail -1 $O/r2_c5_n${N}_sync.log | cut -c1-600
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_n8.json').read().splitlines() if l.startswith('{')][-1])
    print('N=8 bench', d['ms_per_step'], d['value'], d.get('c3_global_batch_8192'), d.get('smoothed_map',{}).get('ms_per_step'))
except Exception as e: print('bench parse', e)
PY
