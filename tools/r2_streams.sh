#!/bin/bash
# Round-2 A/B: text tower on its own stream (fork/join inside the captured step) against one stream.
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2_gputests_k.log 2>&1; tail -4 $O/r2_gputests_k.log
FORDE_TOWER_STREAMS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2d_streams1.json 2> $O/r2d_streams1.err
FORDE_TOWER_STREAMS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2d_streams0.json 2> $O/r2d_streams0.err
FORDE_TOWER_STREAMS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2d_streams1b.json 2> $O/r2d_streams1b.err
for f in r2d_streams1 r2d_streams0 r2d_streams1b; do python - $O/$f.json <<'PY'
import json,sys
try:
    txt=open(sys.argv[1]).read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['ms_per_step'],2), d['e2e']['loss'], round(d['stateful_layer']['ms_per_step_in_kernels'],2), d['clocks']['sm_mhz'] if d['clocks'] else None)
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
