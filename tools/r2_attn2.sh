#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_configs.py -x -q -m gpu -k attention > $O/r2_attn2_tests.log 2>&1; tail -5 $O/r2_attn2_tests.log
timeout 300 python tools/attn_bench.py > $O/r2_attn2_bench.txt 2>&1; tail -6 $O/r2_attn2_bench.txt
timeout 100 python tools/attn_trace.py 2>&1 | tail -16
