#!/bin/bash
# vision-tower attention forward: O read-out through shared memory + TMA stores; early hand-back of block 0's TMEM
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "attn or attention or dual or train or model" > $O/r2g_tests.log 2>&1; tail -3 $O/r2g_tests.log
timeout 200 python tools/attn_trace.py > $O/r2g_trace.txt 2>&1; head -8 $O/r2g_trace.txt
timeout 200 python tools/attn_bench.py > $O/r2g_attn_bench.txt 2>&1; cat $O/r2g_attn_bench.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2g_bench.json 2> $O/r2g_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2g_bench.json').read().splitlines() if l.startswith('{')][-1])
print('bench', round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['ms_per_step'],2), d['e2e']['loss'], d['clocks'])
PY
