"""GPU parity tests of the batched window decoder (k_window_tiles4, decode_tile4.cuh), through the C-ABI.

Snapshot::fill_window (snapshot.rs:204-301) / Log::fill_window (log.rs:311-508) / Chunk::fill_window (chunk.rs:152-158) /
Superchunk::fill_window (superchunk.rs:402-457): the decoded window must equal the encoder's input (float encodings,
round trip) and the oracle's fixed-point window (raw output) bit for bit.  The cases aim at the paths of the kernel:
16-byte / 8-byte / scalar stores (window widths and offsets of every parity), windows that start inside a block
(the snapshot is expanded but not emitted), trees of every depth from 2x2 to 64x64, clipped tiles, elided tiles,
uniform and `equal` sub-trees, NaN cells, 64-bit values, structures larger than the staging buffer (global-memory
path), and the earlier generations of the kernel that stay selectable by environment variable.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from dcdf_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _field(T, R, C, seed, nan_frac=0.0, flat=True):
    """Float raster in units of 1/16 with drifting smooth part, per-cell jitter, flat patches and optional NaNs."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:R, 0:C]
    base = 4000 + 3 * y - 2 * x
    q = np.empty((T, R, C), np.int64)
    for t in range(T):
        jitter = rng.integers(-6, 7, (R, C)) * (rng.random((R, C)) < 0.4)
        q[t] = base + 5 * t + jitter
    if flat:
        q[:, : R // 3, : C // 2] = 4321                   # uniform over space and time: single-value sub-trees / elision
        q[:, R // 2:, C // 2:] = base[R // 2:, C // 2:] + 7 * np.arange(T)[:, None, None]  # snapshot + constant: `equal` nodes
    a = (q / 16.0).astype(np.float32)
    if nan_frac:
        a[:, rng.random((R, C)) < nan_frac] = np.nan
        a[rng.random((T, R, C)) < nan_frac / 4] = np.nan
    return a


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                                 b.view(np.uint32) if b.dtype == np.float32 else b)


def _canon(a):
    """NaN payloads: from_fixed yields the canonical quiet NaN (fixed.rs:81-86); compare NaN positions + finite bits."""
    a = a.copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a


def test_superchunk_windows_of_every_alignment(ctx):
    from dcdf_b200 import Superchunk
    T, R, C = 21, 150, 203
    data = _field(T, R, C, 1, nan_frac=0.05)
    sc = Superchunk.build(ctx, data, [2, 6], chunk_size=8)
    rng = np.random.default_rng(2)
    cubes = [[0, T, 0, R, 0, C], [3, 19, 1, 149, 1, 202], [5, 6, 64, 128, 64, 128], [2, 21, 0, 64, 0, 64],
             [7, 15, 63, 66, 127, 130], [0, T, 149, 150, 0, C], [1, 20, 10, 140, 4, 200], [1, 20, 10, 140, 8, 72]]
    for _ in range(40):
        t0, t1 = sorted(rng.integers(0, T + 1, 2)); r0, r1 = sorted(rng.integers(0, R + 1, 2)); c0, c1 = sorted(rng.integers(0, C + 1, 2))
        if t0 < t1 and r0 < r1 and c0 < c1:
            cubes.append([t0, t1, r0, r1, c0, c1])
    out, off = sc.window_batch(cubes)
    for c, a, b in zip(cubes, off[:-1], off[1:]):
        got = out[int(a):int(b)].reshape(c[1] - c[0], c[3] - c[2], c[5] - c[4])
        assert _same(_canon(got), _canon(data[c[0]:c[1], c[2]:c[3], c[4]:c[5]])), f"window {c}"
    # single windows (output offset 0) incl. raw fixed point against the oracle, tile by tile
    w = sc.window(4, 17, 3, 131, 6, 198)
    assert _same(_canon(w), _canon(data[4:17, 3:131, 6:198]))
    ref = orc.superchunk_build(np.ascontiguousarray(data[:8]), [2, 6])  # the oracle's own decode of the first time slice
    assert _same(_canon(sc.window(0, 8, 0, R, 0, C)), _canon(ref.window_f32(0, 8, 0, R, 0, C)))
    assert _same(_canon(sc.window(2, 7, 5, 133, 9, 190)), _canon(ref.window_f32(2, 7, 5, 133, 9, 190)))
    sc.close()


@pytest.mark.parametrize("rows,cols", [(2, 2), (1, 2), (3, 3), (4, 4), (5, 7), (8, 8), (13, 16), (17, 9), (32, 32), (33, 20), (64, 64), (64, 17), (40, 64)])
def test_chunk_windows_every_tree_depth(ctx, rows, cols):
    from dcdf_b200 import Chunk
    T = 14
    data = _field(T, rows, cols, 10 + rows * 100 + cols, nan_frac=0.08 if rows * cols > 8 else 0.0, flat=rows >= 8)
    got = Chunk.build(ctx, data, fractional_bits=4)
    ref = orc.chunk_build(data, fractional_bits=4)
    assert got.write_to() == ref.serialize()
    assert np.array_equal(got.window(0, T, 0, rows, 0, cols, raw=True), ref.window(0, T, 0, rows, 0, cols))
    assert _same(_canon(got.window(0, T, 0, rows, 0, cols)), _canon(data))
    # windows that start inside a block, one-cell windows, odd offsets
    for (a, b, t, bo, l, r) in [(1, T, 0, rows, 0, cols), (T - 1, T, rows - 1, rows, cols - 1, cols), (3, 9, rows // 3, rows, cols // 2, cols),
                                (5, 6, 0, max(1, rows - 1), 0, max(1, cols - 1))]:
        assert np.array_equal(got.window(a, b, t, bo, l, r, raw=True), ref.window(a, b, t, bo, l, r)), (a, b, t, bo, l, r)


def test_int64_values_and_structures_beyond_the_staging_buffer(ctx):
    """Large random integers: 5-byte DAC codes (64-bit expansion) and Snapshots of ~30 KB, which are decoded from
    global memory because they do not fit the staging buffer."""
    from dcdf_b200 import Chunk
    rng = np.random.default_rng(5)
    T, R, C = 6, 64, 64
    data = rng.integers(-(1 << 36), 1 << 36, (T, R, C)).astype(np.int64)
    data[2] = data[1] + 3                      # a Log that is `equal` at the root
    data[4, :32] = data[3, :32]                # half the tile unchanged
    got = Chunk.build(ctx, data)
    ref = orc.chunk_build(data)
    assert got.write_to() == ref.serialize()
    assert np.array_equal(got.window(0, T, 0, R, 0, C), data)
    assert np.array_equal(got.window(1, 5, 3, 61, 2, 63), data[1:5, 3:61, 2:63])
    d32 = rng.integers(-(1 << 20), 1 << 20, (T, R, C)).astype(np.int32)   # 3-byte codes: 32-bit expansion, large structures
    got32 = Chunk.build(ctx, d32)
    assert got32.write_to() == orc.chunk_build(d32).serialize()
    assert np.array_equal(got32.window(0, T, 0, R, 0, C), d32)
    assert np.array_equal(got32.window(2, 6, 1, 64, 5, 60), d32[2:6, 1:64, 5:60])


def test_device_output_and_elided_tiles(ctx):
    import torch
    from dcdf_b200 import Superchunk
    T, R, C = 16, 130, 140
    data = _field(T, R, C, 9)
    data[:, :64, :64] = 1.5                   # a whole subchunk uniform at every instant -> Elided (superchunk.rs:145-151)
    sc = Superchunk.build(ctx, data, [2, 6], chunk_size=8)
    out = torch.empty((T, R, C), device="cuda", dtype=torch.float32)
    sc.window(0, T, 0, R, 0, C, out=out)
    assert torch.equal(out.cpu(), torch.from_numpy(data))
    view = torch.empty((T - 3) * 100 * 101 + 1, device="cuda", dtype=torch.float32)[1:]  # 4-byte aligned output only
    sc.window_batch([[3, T, 20, 120, 30, 131]], out=view)
    assert torch.equal(view.cpu().reshape(T - 3, 100, 101), torch.from_numpy(data[3:T, 20:120, 30:131]))
    sc.close()


@pytest.mark.parametrize("option", ["window_cells", "window_wide"])
def test_other_window_kernels_stay_bit_exact(option):
    """The per-cell kernel (used for trees larger than 64x64) and the 64-bit tile expansion, selected per context."""
    from dcdf_b200 import Context, Superchunk
    ctx = Context(0)
    ctx.set_option(option, 1)
    data = _field(19, 150, 203, 3, nan_frac=0.05)
    sc = Superchunk.build(ctx, data, [2, 6], chunk_size=8)
    for c in ([0, 19, 0, 150, 0, 203], [3, 18, 5, 149, 7, 202], [9, 10, 64, 128, 0, 64]):
        w = sc.window(*c)
        assert _same(_canon(w), _canon(data[c[0]:c[1], c[2]:c[3], c[4]:c[5]])), c
    sc.close()
    ctx.close()


@pytest.mark.parametrize("rows,cols", [(2, 2), (1, 2), (3, 3), (4, 4), (5, 7), (8, 8), (17, 9), (33, 20), (64, 64), (100, 130), (256, 256)])
def test_block_walker_every_tree_depth(rows, cols):
    """k_window_blocks (one thread per 4x4 block, the descent shared by its cells): the kernel of trees larger than 64x64,
    forced onto the small ones too; raw fixed values against the oracle's fill_window, windows clipped at every block phase."""
    from dcdf_b200 import Chunk, Context
    ctx = Context(0)
    ctx.set_option("window_cells", 1)
    T = 14
    data = _field(T, rows, cols, 77 + rows * 100 + cols, nan_frac=0.08 if rows * cols > 8 else 0.0, flat=rows >= 8)
    got = Chunk.build(ctx, data, fractional_bits=4)
    ref = orc.chunk_build(data, fractional_bits=4)
    assert got.write_to() == ref.serialize()
    wins = [(0, T, 0, rows, 0, cols), (1, T, 0, rows, 0, cols), (T - 1, T, rows - 1, rows, cols - 1, cols),
            (3, 9, rows // 3, rows, cols // 2, cols), (5, 6, 0, max(1, rows - 1), 0, max(1, cols - 1))]
    for k in range(1, 8):                                   # every phase of the 4x4 block grid on both edges
        wins.append((2, 7, min(k, rows - 1), max(min(k, rows - 1) + 1, rows - k % 3), min(7 - k, cols - 1), max(min(7 - k, cols - 1) + 1, cols - k % 4)))
    for (a, b, t, bo, l, r) in wins:
        assert np.array_equal(got.window(a, b, t, bo, l, r, raw=True), ref.window(a, b, t, bo, l, r)), (a, b, t, bo, l, r)
    got.close()
    ctx.close()
