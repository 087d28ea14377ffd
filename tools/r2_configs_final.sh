#!/bin/bash
# Final-code records of the other BASELINE configs on ONE B200: C1 (tiny, batch 32), C4 (MoE slot + decoder LM step),
# C5's model (ViT-L/14) as a bench line and through train.fit with the slow loop every 100 steps.
O=gpurun_out
timeout 300 python bench.py --config tiny --batch 32 --steps 20 --warmup 3 --no-smoothed-map > $O/r2j_bench_c1_tiny.json 2> $O/r2j_c1.err
timeout 400 python bench.py --config vit_l14 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2j_bench_c5_vit_l14.json 2> $O/r2j_c5.err
timeout 300 python tools/moe_bench.py > $O/r2j_moe_bench_c4.txt 2>&1
timeout 300 python tools/llm_bench.py > $O/r2j_llm_config4_step.txt 2>&1
timeout 600 python tools/fit_bench.py --config vit_l14 --steps 210 --interval 100 --overlap 20 --json $O/r2j_c5_fit_n1_overlap20.json > $O/r2j_c5_fit.log 2>&1
python - <<'PY'
import json
for f in ('r2j_bench_c1_tiny','r2j_bench_c5_vit_l14'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().splitlines() if l.startswith('{')][-1])
        print(f, round(d['ms_per_step'],2), round(d['value']), d['roofline']['kernel'], d['roofline']['frac'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2j_moe_bench_c4.txt; tail -3 gpurun_out/r2j_llm_config4_step.txt; tail -2 gpurun_out/r2j_c5_fit.log | cut -c1-500
