#!/bin/bash
# Last refresh of the round: GPU suite + smoke, bench line, attention micro-benchmark, ncu --set full of the non-GEMM kernels
O=gpurun_out
sha256sum forde-vision-language_b200/libforde_b200.so | cut -c1-16 > $O/r2q_lib_sha.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2q_gputests.log 2>&1; tail -3 $O/r2q_gputests.log
python __graft_entry__.py --smoke 2>&1 | tail -1 > $O/r2q_smoke.txt; cat $O/r2q_smoke.txt
timeout 600 python bench.py > $O/r2q_bench_n1.json 2> $O/r2q_bench_n1.err; tail -c 200 $O/r2q_bench_n1.json
timeout 300 python tools/attn_bench.py > $O/r2q_attn_bench.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn|layernorm|colsum' -c 18 -f -o $O/r2q_aux python tools/aux_once.py > $O/r2q_ncu.log 2>&1
ncu -i $O/r2q_aux.ncu-rep --page raw --csv > $O/r2q_aux_raw.csv 2>/dev/null
rm -f $O/r2q_aux.ncu-rep
python tools/ncu_aux_table.py $O/r2q_aux_raw.csv | cut -c1-200
