#!/bin/bash
# Round-2 two-GPU experiments (gpurun --gpus 2): NCCL parity test log, gradient-exchange variants, 1024 pairs/GPU, config 5 fit.
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -rA > $O/r2_n2_dist_test.log 2>&1; tail -3 $O/r2_n2_dist_test.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-smoothed-map > $O/r2_n1_samebox.json 2> $O/r2_n1_samebox.err
FORDE_GRAD_COMM=fp32 timeout 300 $T --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2_n2_fp32.json 2> $O/r2_n2_fp32.err
timeout 300 $T --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2_n2_bf16.json 2> $O/r2_n2_bf16.err
NCCL_MAX_CTAS=8 timeout 300 $T --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2_n2_bf16_cta8.json 2> $O/r2_n2_bf16_cta8.err
NCCL_MAX_CTAS=4 timeout 300 $T --master-port 29514 bench.py --gpus 2 --steps 8 --warmup 3 --no-smoothed-map > $O/r2_n2_bf16_cta4.json 2> $O/r2_n2_bf16_cta4.err
timeout 400 $T --master-port 29515 bench.py --gpus 2 --batch 1024 --steps 4 --warmup 3 --no-smoothed-map > $O/r2_n2_b1024.json 2> $O/r2_n2_b1024.err
timeout 400 $T --master-port 29516 tools/fit_bench.py --config vit_l14 --steps 210 --interval 100 --overlap 0 --json $O/r2_c5_n2_sync.json > $O/r2_c5_n2_sync.log 2>&1
timeout 400 $T --master-port 29517 tools/fit_bench.py --config vit_l14 --steps 210 --interval 100 --overlap 10 --json $O/r2_c5_n2_ovl.json > $O/r2_c5_n2_ovl.log 2>&1
for f in r2_n1_samebox r2_n2_fp32 r2_n2_bf16 r2_n2_bf16_cta8 r2_n2_bf16_cta4 r2_n2_b1024; do echo "$f: $(python -c "import json,sys; d=json.load(open('$O/$f.json')); print(d['ms_per_step'], d['value'], d['stateful_layer']['kernels'].get('sl_dz_bwd[50432x768x3072]', d['stateful_layer']['kernels'].get('sl_dz_bwd[201728x768x3072]')))" 2>&1 | tail -1)"; done
tail -2 $O/r2_c5_n2_sync.log | cut -c1-900; tail -2 $O/r2_c5_n2_ovl.log | cut -c1-900
