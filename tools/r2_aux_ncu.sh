#!/bin/bash
# ncu --set full of the non-GEMM kernels of a block (attention, add/norm, column sum) + the stream-equivalence test
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/r2n_tests.log 2>&1; tail -3 $O/r2n_tests.log
timeout 200 python tools/aux_once.py > $O/r2n_once.log 2>&1 || tail -5 $O/r2n_once.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn|layernorm|colsum' -c 18 -f -o $O/r2n_aux python tools/aux_once.py > $O/r2n_ncu.log 2>&1
ncu -i $O/r2n_aux.ncu-rep --page raw --csv > $O/r2n_aux_raw.csv 2>/dev/null
python tools/ncu_aux_table.py $O/r2n_aux_raw.csv
