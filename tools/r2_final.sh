#!/bin/bash
# End-of-round evidence on ONE B200 from the library that ships: GPU test suite, bench line, reference arm,
# launch list of a train step (ncu, gpu__time_duration), ncu --set full of the six StatefulLayer GEMMs + attention.
O=gpurun_out
sha256sum forde-vision-language_b200/libforde_b200.so | cut -c1-16 > $O/r2f_lib_sha.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2f_gputests.log 2>&1; tail -3 $O/r2f_gputests.log
timeout 600 python bench.py > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err; tail -c 600 $O/r2f_bench_n1.json
timeout 300 python tools/layer_bench.py --assign zeros > $O/r2f_layer_zeros.txt 2>&1
timeout 300 python tools/layer_bench.py --assign blocks > $O/r2f_layer_blocks.txt 2>&1
timeout 300 python tools/layer_bench.py --M 19712 --d 512 --H 2048 --assign zeros > $O/r2f_layer_text.txt 2>&1
timeout 300 python tools/attn_bench.py > $O/r2f_attn_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2f_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-smoothed-map --no-graph > $O/r2f_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forde_gemm_kernel -c 6 -f -o $O/r2f_layer \
   python tools/layer_bench.py --once --assign zeros > $O/r2f_ncu_layer.log 2>&1
ncu -i $O/r2f_layer.ncu-rep --page raw --csv > $O/r2f_layer_raw.csv 2>/dev/null
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err; tail -c 400 $O/r2f_bench_ref.json
ls -la $O | head -30
