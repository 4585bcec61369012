"""GPU parity tests aimed at the full-tile encoder (csrc/encode_v4.cuh: 64x64 tiles, 31-bit fixed values):
every DAC length class on both candidates, uniform / equal sub-trees at each level, the heuristic mix, the
254-log cap, non-vector strides, and the direct-to-arena emission path.  Bit-exact against the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["staged", "direct"])
def ctx(request):
    """Every test of this module runs twice: with the shared-memory emission image (default) and with option
    stage_limit = 0, which forces every structure through the direct (global-memory) emission path."""
    from dcdf_b200 import Context
    c = Context(0)
    if request.param == "direct":
        c.set_option("stage_limit", 0)
    yield c
    c.close()


def _check(ctx, data, fractional_bits=0, round_=False):
    from dcdf_b200 import Chunk
    ref = orc.chunk_build(data, fractional_bits=fractional_bits, round_=round_)
    got = Chunk.build(ctx, data, fractional_bits=fractional_bits, round=round_)
    assert got.block_instants() == ref.block_instants(), "heuristic decisions differ"
    a, b = got.write_to(), ref.serialize()
    if a != b:
        n = min(len(a), len(b))
        i = next((j for j in range(n) if a[j] != b[j]), n)
        raise AssertionError(f"bytes differ at {i} (gpu {len(a)}, oracle {len(b)}): {a[i:i+16].hex()} vs {b[i:i+16].hex()}")
    T, R, C = data.shape
    assert np.array_equal(got.window(0, T, 0, R, 0, C, raw=True), ref.window(0, T, 0, R, 0, C))
    return got


def _frames(rng, kind, T=10):
    """[T, 64, 64] int64 rasters built to hit specific tree shapes."""
    base = rng.integers(0, 50, (64, 64))
    out = [base]
    for i in range(1, T):
        f = out[-1].copy()
        if kind == "sparse":            # a few cells change: logs with mostly equal quads
            m = rng.random((64, 64)) < 0.03
            f[m] += rng.integers(-3, 4, m.sum())
        elif kind == "offset":          # whole-field offsets: single-node equal logs
            f = f + int(rng.integers(-5, 6))
        elif kind == "blocks":          # changes confined to one quadrant / 16x16 / 8x8 block
            s = [32, 16, 8, 4, 2][i % 5]
            r, c = rng.integers(0, 64 // s, 2) * s
            f[r:r + s, c:c + s] += rng.integers(-200, 200, (s, s))
        elif kind == "two":             # 2-byte entries: diffs up to +-20000
            f = f + rng.integers(-20000, 20000, (64, 64))
        elif kind == "three":           # 3-byte entries
            f = f + rng.integers(-3_000_000, 3_000_000, (64, 64)) * (rng.random((64, 64)) < 0.3)
        elif kind == "four":            # 4-byte entries (|fixed| stays below 2^30)
            f = rng.integers(-2 ** 28, 2 ** 28, (64, 64)) * (rng.random((64, 64)) < 0.5)
        elif kind == "mixed":
            j = i % 6
            if j == 0:
                f = rng.integers(0, 3000, (64, 64))
            elif j == 1:
                f = f + 7
            elif j == 2:
                f[:32, :32] = 5
            elif j == 3:
                f[rng.random((64, 64)) < 0.5] += 300
            elif j == 4:
                f[10:12, 20:22] -= 100000
            else:
                f = f * 0 + 9
        out.append(f)
    return np.stack(out).astype(np.int64)


@pytest.mark.parametrize("kind", ["sparse", "offset", "blocks", "two", "three", "four", "mixed"])
def test_full_tile_int(ctx, kind):
    rng = np.random.default_rng(["sparse", "offset", "blocks", "two", "three", "four", "mixed"].index(kind) + 20)
    data = _frames(rng, kind)
    _check(ctx, data)
    _check(ctx, data.astype(np.int32))


@pytest.mark.parametrize("rows,cols", [(64, 40), (40, 64), (33, 64), (64, 33), (50, 50), (63, 63), (37, 64)])
def test_clipped_tile_with_64_side_tree(ctx, rows, cols):
    """Clipped tiles that keep a 64-side tree go through the full-tile encoder with None cells; large values make the
    entries of cells outside the raster (quad max - 0) two to four bytes long."""
    rng = np.random.default_rng(rows * 64 + cols)
    for scale, span in ((1, 50), (3000, 2000), (200_000, 100_000), (2 ** 27, 2 ** 26)):
        base = scale + rng.integers(0, span, (rows, cols))
        frames = [base]
        for i in range(1, 9):
            f = frames[-1].copy()
            if i % 4 == 0:
                f = scale + rng.integers(0, span, (rows, cols))
            elif i % 4 == 1:
                f = f + 5
            else:
                m = rng.random((rows, cols)) < 0.2
                f[m] += rng.integers(-span // 4 - 1, span // 4 + 1, m.sum())
            frames.append(f)
        data = np.stack(frames).astype(np.int64)
        _check(ctx, data)
    f32 = (np.stack(frames) % 4096 / 8.0).astype(np.float32)
    f32[rng.random(f32.shape) < 0.1] = np.nan
    _check(ctx, f32, fractional_bits=3)


def test_full_tile_uniform_levels(ctx):
    """Uniform and equal sub-trees at every level of the quadtree, including the root."""
    rng = np.random.default_rng(11)
    frames = []
    for s in [64, 32, 16, 8, 4, 2, 1]:
        f = np.repeat(np.repeat(rng.integers(0, 1000, (64 // s, 64 // s)), s, 0), s, 1)
        frames.append(f)
        frames.append(f + 3)                 # equal at the root
        g = f.copy()
        g[:s, :s] += 1                       # equal everywhere but one node
        frames.append(g)
    _check(ctx, np.stack(frames).astype(np.int64))


def test_full_tile_float_nan_round(ctx):
    rng = np.random.default_rng(12)
    base = rng.integers(0, 4000, (64, 64))
    frames = [base]
    for i in range(11):
        f = frames[-1].copy()
        m = rng.random((64, 64)) < (0.9 if i % 4 == 3 else 0.05)
        f[m] += rng.integers(-40, 40, m.sum())
        frames.append(f)
    f32 = (np.stack(frames) / 8.0).astype(np.float32)
    f32[rng.random(f32.shape) < 0.02] = np.nan
    f32[3, 16:32, :] = np.nan              # NaN sub-trees (fixed value 0)
    _check(ctx, f32, fractional_bits=3)
    f64 = rng.normal(100, 30, (6, 64, 64))
    _check(ctx, f64, fractional_bits=10, round_=True)
    _check(ctx, f64.astype(np.float32), fractional_bits=6, round_=True)


def test_full_tile_log_cap(ctx):
    rng = np.random.default_rng(13)
    base = rng.integers(0, 50, (64, 64)).astype(np.int64)
    frames = []
    for i in range(300):
        f = base.copy()
        f[i % 64, (7 * i) % 64] += 1
        frames.append(f)
    got = _check(ctx, np.stack(frames))
    assert got.block_instants() == [255, 45]


def test_full_tile_strided_and_device_views(ctx):
    import torch
    from dcdf_b200 import Chunk
    rng = np.random.default_rng(14)
    big = rng.integers(0, 500, (9, 70, 131)).astype(np.int32)
    big[1:] = big[0] + rng.integers(-2, 3, big[1:].shape) * (rng.random(big[1:].shape) < 0.1)
    view = big[1:8, 3:67, 5:69]                       # rows not 16-byte aligned -> scalar loads
    ref = orc.chunk_build(np.ascontiguousarray(view))
    got = Chunk.build(ctx, torch.from_numpy(big).cuda()[1:8, 3:67, 5:69])
    assert got.write_to() == ref.serialize()
    view2 = big[:, 2:66, 1:129:2]                     # column stride 2
    ref = orc.chunk_build(np.ascontiguousarray(view2))
    got = Chunk.build(ctx, torch.from_numpy(big).cuda()[:, 2:66, 1:129:2])
    assert got.write_to() == ref.serialize()


def test_superchunk_many_full_tiles(ctx):
    """A 3x4 grid of tiles (some clipped) over 70 instants: units race on the arena and the work lists."""
    from dcdf_b200 import Superchunk, synth
    data = synth.raster_slice(0, 70, 150, 230).numpy()
    sc = Superchunk.build(ctx, data, [2, 6])
    ref = orc.superchunk_build(data, [2, 6])
    kinds, child = ref.node_refs(0)
    chunks = sc.chunk_bytes(0)
    for slot, (k, c) in enumerate(zip(kinds, child)):
        assert (chunks[slot] is None) == (k == 0)
        if k:
            assert chunks[slot] == ref.node_bytes(int(c)), f"slot {slot}"
    assert np.array_equal(sc.window(0, 70, 0, 150, 0, 230), data)
