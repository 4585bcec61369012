"""Short run of the random-access queries for ncu captures: 8 slices of the C2 grid, 4096 full-length cell series
(configs[2] style) and 2048 value-range searches over 8..256-sided windows (configs[3] style)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dcdf_b200 import Context, Superchunk, synth, _ffi
T, R, C = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 721, 1440
data = synth.raster(T, R, C, device="cuda")
ctx = Context(0)
for a in sys.argv[3:]:
    k, v = a.split("=")
    ctx.set_option(k, int(v))
sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
rng = np.random.default_rng(7)
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
q = np.stack([np.zeros(nq, np.int64), np.full(nq, T, np.int64), rng.integers(0, R, nq), rng.integers(0, C, nq)], axis=1)
sc.cell_batch(q[:64])
series = sc.cell_batch(q)
print("cell series ms", ctx.last_kernel_ms(_ffi.KT_CELL), "cells", nq * T,
      "ok", bool(np.array_equal(series[5], data[:, int(q[5, 2]), int(q[5, 3])].cpu().numpy())))
nw = 2048
side = rng.integers(8, 257, nw)
top = rng.integers(0, R - 8, nw); left = rng.integers(0, C - 8, nw)
t0 = rng.integers(0, max(T - 64, 1), nw)
cubes = np.stack([t0, np.minimum(t0 + 64, T), top, np.minimum(top + side, R), left, np.minimum(left + side, C)], axis=1)
lo = rng.integers(270 * 32, 300 * 32, nw)
counts, cells = sc.search_batch(cubes, lo, lo + 48)
vol = int(((cubes[:, 1] - cubes[:, 0]) * (cubes[:, 3] - cubes[:, 2]) * (cubes[:, 5] - cubes[:, 4])).sum())
print("search ms", ctx.last_kernel_ms(_ffi.KT_SEARCH), "cells scanned", vol, "matches", int(counts.sum()))
