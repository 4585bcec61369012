"""Dense query batches on 8 slices of the C2 grid: 32 k full-length cell series (k_cell_tiles4) and a count-only search
batch of 16 k windows (k_count_tiles4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dcdf_b200 import Context, Superchunk, synth, _ffi
T, R, C = 512, 721, 1440
data = synth.raster(T, R, C, device="cuda")
ctx = Context(0)
sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
rng = np.random.default_rng(7)
nq = 32768
q = np.stack([np.zeros(nq, np.int64), np.full(nq, T, np.int64), rng.integers(0, R, nq), rng.integers(0, C, nq)], axis=1)
sc.cell_batch(q)
print("cell series ms", ctx.last_kernel_ms(_ffi.KT_CELL), "cells", nq * T)
nw = 16384
side = rng.integers(8, 257, nw)
top = rng.integers(0, R - 8, nw); left = rng.integers(0, C - 8, nw)
t0 = rng.integers(0, max(T - 64, 1), nw)
cubes = np.stack([t0, np.minimum(t0 + 64, T), top, np.minimum(top + side, R), left, np.minimum(left + side, C)], axis=1)
lo = rng.integers(270 * 32, 300 * 32, nw)
counts, _ = sc.search_batch(cubes, lo, lo + 48, want_cells=False)
vol = int(((cubes[:, 1] - cubes[:, 0]) * (cubes[:, 3] - cubes[:, 2]) * (cubes[:, 5] - cubes[:, 4])).sum())
print("count-only search ms", ctx.last_kernel_ms(_ffi.KT_SEARCH), "cells scanned", vol, "matches", int(counts.sum()))
