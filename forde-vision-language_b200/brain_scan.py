"""Slow-loop diagnostics written to disk: the "Brain Scan" heat map and the feature-space scatter.

Mirrors /root/reference/src/utils/logging.py:7-52 (`plot_brain_scan(grid, step, epoch)`,
`plot_feature_space(stats, assignments, step, epoch)`: same arguments, same output file names).  The reference draws
with matplotlib/seaborn, which this image does not have, so the PNGs are rasterised here with numpy and encoded
with zlib; no axes or colour bar, one colour per cluster id.  Host-side only (runs once per slow-loop interval).
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

# viridis end points and midpoint for ids 0..2, then a tab10-like cycle for larger cluster counts
_PALETTE = np.array(
    [[68, 1, 84], [33, 145, 140], [253, 231, 37], [214, 39, 40], [148, 103, 189], [140, 86, 75], [227, 119, 194],
     [127, 127, 127], [188, 189, 34], [23, 190, 207]], dtype=np.uint8)


def _write_png(path: str, rgb: np.ndarray) -> None:
    """rgb: (h, w, 3) uint8."""
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()  # filter byte 0 per row

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def plot_brain_scan(smoothed_assignments_grid, step, epoch, out_dir: str = ".", cell: int = 8) -> str:
    """Heat map of the smoothed neuron-assignment grid, one `cell` x `cell` square per neuron
    (logging.py:7-20).  Returns the path written: brain_scan_epoch{epoch}_step{step}.png."""
    grid = np.asarray(smoothed_assignments_grid)
    if grid.ndim != 2:
        raise ValueError(f"brain scan needs a 2-D grid, got shape {grid.shape}")
    ids = np.clip(grid.astype(np.int64), 0, len(_PALETTE) - 1)
    img = _PALETTE[ids]
    img = np.repeat(np.repeat(img, cell, axis=0), cell, axis=1)
    path = os.path.join(out_dir, f"brain_scan_epoch{epoch}_step{step}.png")
    _write_png(path, img)
    return path


def plot_feature_space(flattened_stats, assignments, step, epoch, out_dir: str = ".", size: int = 512):
    """Scatter of the first two stat columns coloured by cluster (logging.py:23-52).  Like the reference it warns
    and returns None when there are fewer than two features."""
    stats = np.asarray(flattened_stats, dtype=np.float64)
    a = np.asarray(assignments).astype(np.int64).reshape(-1)
    if stats.ndim != 2 or stats.shape[1] < 2:
        print(f"Warning: Cannot plot feature space for {stats.shape[-1] if stats.ndim else 0} features. Need at least 2.")
        return None
    if stats.shape[0] != a.shape[0]:
        raise ValueError("flattened_stats and assignments disagree on the number of neurons")
    img = np.full((size, size, 3), 255, np.uint8)
    if stats.shape[0]:
        xy = stats[:, :2]
        lo, hi = xy.min(axis=0), xy.max(axis=0)
        span = np.where(hi > lo, hi - lo, 1.0)
        pad = 8
        px = (pad + (xy[:, 0] - lo[0]) / span[0] * (size - 1 - 2 * pad)).astype(np.int64)
        py = (size - 1 - pad - (xy[:, 1] - lo[1]) / span[1] * (size - 1 - 2 * pad)).astype(np.int64)
        col = _PALETTE[np.clip(a, 0, len(_PALETTE) - 1)]
        for dy in (-1, 0, 1):  # 3 x 3 markers
            for dx in (-1, 0, 1):
                img[np.clip(py + dy, 0, size - 1), np.clip(px + dx, 0, size - 1)] = col
    path = os.path.join(out_dir, f"feature_space_epoch{epoch}_step{step}.png")
    _write_png(path, img)
    return path
