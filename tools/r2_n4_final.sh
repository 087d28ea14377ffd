#!/bin/bash
# Four GPUs, final library: bench line (weak scaling, 256 pairs per GPU)
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2k_bench_n4.json 2> $O/r2k_bench_n4.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2k_bench_n4.json').read().splitlines() if l.startswith('{')][-1])
    print('n4', d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['ms_per_step'],2), d['e2e']['loss'], d['smoothed_map']['ms_per_step'], d['clocks'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r2k_bench_n4.err').read()[-2000:])
PY
