#!/bin/bash
# Round-2 two-GPU attribution of what is left between 1 and 2 GPUs: bucket shapes, no exchange at all, NCCL CTA cap.
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map"
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2_gputests_d.log 2>&1; tail -4 $O/r2_gputests_d.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2c_n1.json 2> $O/r2c_n1.err
i=0
run() { name=$1; shift; i=$((i+1)); env "$@" timeout 300 $T --master-port $((29530+i)) $B > $O/r2c_$name.json 2> $O/r2c_$name.err; }
run taper4 FORDE_BUCKET_TAPER=1
run notaper4 FORDE_BUCKET_TAPER=0
run skip FORDE_SKIP_GRAD_EXCHANGE=1
run layers2 FORDE_BUCKET_LAYERS=2
run layers1 FORDE_BUCKET_LAYERS=1
run layers12 FORDE_BUCKET_LAYERS=12
run cta8 NCCL_MAX_CTAS=8
run cta16 NCCL_MAX_CTAS=16
for f in n1 taper4 notaper4 skip layers2 layers1 layers12 cta8 cta16; do python - $O/r2c_$f.json <<'PY'
import json,sys
try:
    txt=open(sys.argv[1]).read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], round(d['ms_per_step'],2), round(d['value']), round(d['stateful_layer']['ms_per_step_in_kernels'],2), d['clocks']['sm_mhz'] if d['clocks'] else None)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
