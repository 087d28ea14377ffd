#!/bin/bash
# Refresh of the end-of-round evidence after the last attention change: GPU suite, bench line, ncu --set full of the six
# StatefulLayer GEMMs (traffic table keyed to this library), attention micro-benchmark.
O=gpurun_out
sha256sum forde-vision-language_b200/libforde_b200.so | cut -c1-16 > $O/r2p_lib_sha.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2p_gputests.log 2>&1; tail -3 $O/r2p_gputests.log
timeout 600 python bench.py > $O/r2p_bench_n1.json 2> $O/r2p_bench_n1.err; tail -c 300 $O/r2p_bench_n1.json
timeout 300 python tools/attn_bench.py > $O/r2p_attn_bench.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forde_gemm_kernel -c 6 -f -o $O/r2p_layer \
   python tools/layer_bench.py --once --assign zeros > $O/r2p_ncu_layer.log 2>&1
ncu -i $O/r2p_layer.ncu-rep --page raw --csv > $O/r2p_layer_raw.csv 2>/dev/null
rm -f $O/r2p_layer.ncu-rep
python __graft_entry__.py --smoke 2>&1 | tail -1
