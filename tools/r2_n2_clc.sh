#!/bin/bash
# Round-2 two-GPU A/B of the dynamic (cluster launch control) tile schedule against the static one, and of overlapping the
# gradient exchange with the backward at all.
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2_gputests_c.log 2>&1; tail -4 $O/r2_gputests_c.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2b_n1_clc1.json 2> $O/r2b_n1_clc1.err
FORDE_CLC=0 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2b_n1_clc0.json 2> $O/r2b_n1_clc0.err
timeout 300 $T --master-port 29521 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2b_n2_clc1.json 2> $O/r2b_n2_clc1.err
FORDE_CLC=0 timeout 300 $T --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2b_n2_clc0.json 2> $O/r2b_n2_clc0.err
FORDE_GRAD_OVERLAP=0 timeout 300 $T --master-port 29523 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2b_n2_clc1_noovl.json 2> $O/r2b_n2_clc1_noovl.err
FORDE_GRAD_COMM=fp32 timeout 300 $T --master-port 29524 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2b_n2_clc1_fp32.json 2> $O/r2b_n2_clc1_fp32.err
for f in r2b_n1_clc1 r2b_n1_clc0 r2b_n2_clc1 r2b_n2_clc0 r2b_n2_clc1_noovl r2b_n2_clc1_fp32; do python - $O/$f.json <<'PY'
import json,sys
try:
    txt=open(sys.argv[1]).read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], round(d['ms_per_step'],2), round(d['value']), round(d['stateful_layer']['ms_per_step_in_kernels'],2), d['clocks']['sm_mhz'] if d['clocks'] else None)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
