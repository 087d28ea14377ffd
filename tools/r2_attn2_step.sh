#!/bin/bash
# vision-tower attention on this repository's kernels (forde_attn2) against cuDNN, inside the train step
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2_gputests_l.log 2>&1; tail -4 $O/r2_gputests_l.log
for pf in 1 0 1 0; do FORDE_ATTN2_L2PF=$pf timeout 200 python tools/attn_bench.py 2>&1 | grep "^vision"; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2e_attn_own.json 2> $O/r2e_attn_own.err
FORDE_ATTN_MAX_TOKENS=128 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2e_attn_cudnn.json 2> $O/r2e_attn_cudnn.err
for f in r2e_attn_own r2e_attn_cudnn; do python - $O/$f.json <<'PY'
import json,sys
try:
    txt=open(sys.argv[1]).read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['ms_per_step'],2), d['e2e']['loss'], d['clocks']['sm_mhz'] if d['clocks'] else None)
except Exception as e: print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
