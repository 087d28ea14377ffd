#!/bin/bash
# Eight GPUs, final library: bench line (weak scaling, 256 pairs per GPU) with the config-3 block (global batch 8192)
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2l_bench_n8.json 2> $O/r2l_bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2l_bench_n8.json').read().splitlines() if l.startswith('{')][-1])
    print('n8', d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['ms_per_step'],2), d['e2e']['loss'], d['smoothed_map']['ms_per_step'], d.get('c3_global_batch_8192'), d['clocks'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2l_bench_n8.err').read()[-2000:])
PY
