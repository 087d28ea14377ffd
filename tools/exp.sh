timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for m in mixed zeros blocks; do timeout 100 python tools/layer_bench.py --iters 10 --assign $m 2>&1 | grep -E "StatefulLayer|up_fwd|dz_bwd|layer fwd"; done
timeout 100 python tools/layer_bench.py --iters 10 --assign zeros --M 19712 --d 512 --H 2048 2>&1 | grep -E "StatefulLayer|up_fwd|dz_bwd|layer fwd"
