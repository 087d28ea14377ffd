for d in 0 64; do echo "FORDE_DEBUG=$d"; FORDE_DEBUG=$d timeout 200 python tools/layer_bench.py --iters 10 2>&1 | grep -E "up_fwd|dz_bwd"; done
