"""Wave-quantisation report for the GEMM shapes of a config (host arithmetic only, no GPU needed).

Restates the launch planning of csrc/forde_api.cu (choose_pair, fill_common, choose_split_k): tile = 128 x 256, or
256 x 256 for a 2-CTA pair; pairs are used when there are at least sm/2 pair tiles (split-K counted for the fp32
accumulate GEMMs); one persistent worker (CTA or CTA pair) per SM (pair of SMs).  Efficiency = work items /
(workers x waves): how much of the last wave is idle.
Usage: python tools/wave_report.py [--config vit_b16] [--batch 256] [--sms 148]"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BM, BN, BK = 128, 256, 64


def plan(M, N, K, can_split_k, sms):
    pair_tiles = math.ceil(M / (2 * BM)) * math.ceil(N / BN)
    work = pair_tiles
    kb = math.ceil(K / BK)
    if can_split_k and kb // 8 > 1:
        work *= kb // 8
    pair = work >= sms // 2
    tile_m = 2 * BM if pair else BM
    tiles = math.ceil(M / tile_m) * math.ceil(N / BN)
    workers = sms // 2 if pair else sms
    split = 1
    if can_split_k:
        split = max(1, min(workers // tiles, max(kb // 8, 1)))
        per = math.ceil(kb / split)
        split = math.ceil(kb / per)
    items = tiles * split
    waves = math.ceil(items / workers)
    return dict(pair=pair, tiles=tiles, split_k=split, items=items, workers=workers, waves=waves,
                eff=items / (workers * waves))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="vit_b16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sms", type=int, default=148)
    a = ap.parse_args()
    import forde_b200  # noqa: F401
    from forde_b200 import dual_encoder
    cfg = dual_encoder.CONFIGS[a.config]()
    towers = [("vision", a.batch * cfg.vision_tokens, cfg.vision.width, cfg.vision.hidden),
              ("text", a.batch * cfg.text_len, cfg.text.width, cfg.text.hidden)]
    print(f"{a.config}, {a.batch} pairs, {a.sms} SMs   (GEMM rows x cols x reduction as launched: D[M][N] = A . B^T)")
    print(f"{'tower':7s}{'GEMM':26s}{'M':>7s}{'N':>7s}{'K':>7s}  pair  items  workers  waves  efficiency")
    for name, M, d, H in towers:
        shapes = [
            ("sl_up_fwd  (H, M, d)", H, M, d, False), ("sl_down_fwd (M, d, H)", M, d, H, False),
            ("sl_dz_bwd  (H, M, d)", H, M, d, False), ("sl_dx_bwd  (M, d, H)", M, d, H, False),
            ("sl_dw_up   (d, H, M)", d, H, M, True), ("sl_dw_down (H, d, M)", H, d, M, True),
            ("dense d->d fwd/dx", M, d, d, False), ("dense d->d dW", d, d, M, True),
        ]
        for label, m, n, k, sk in shapes:
            p = plan(m, n, k, sk, a.sms)
            print(f"{name:7s}{label:26s}{m:7d}{n:7d}{k:7d}  {'yes' if p['pair'] else 'no ':4s}{p['items']:6d}{p['workers']:8d}"
                  f"{p['waves']:7d}{100 * p['eff']:10.1f} %")


if __name__ == "__main__":
    main()
