#!/bin/bash
# Two GPUs, final library: NCCL parity test, attention checks, 2-GPU bench line (text tower on its own stream + bucketed
# bf16 gradient exchange), 1-GPU line on the same box.
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -rA > $O/r2h_dist.log 2>&1; tail -5 $O/r2h_dist.log
timeout 600 python -m pytest tests -x -q -m gpu -k "attn or attention" > $O/r2h_attn_tests.log 2>&1; tail -2 $O/r2h_attn_tests.log
timeout 200 python tools/attn_bench.py > $O/r2h_attn_bench.txt 2>&1; grep vision $O/r2h_attn_bench.txt
timeout 200 python tools/attn_trace.py 2>/dev/null | sed -n 4,7p
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2h_bench_n2.json 2> $O/r2h_bench_n2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err
python - <<'PY'
import json
for f in ('r2h_bench_n2','r2h_bench_n1'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().splitlines() if l.startswith('{')][-1])
        print(f, d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['ms_per_step'],2), d['e2e']['loss'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
