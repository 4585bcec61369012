"""Cell series through the tile decoder (k_cell_tiles4): a tile that many series of one batch fall into is expanded once per
instant and hands out the asked-for cells; the result must be what Chunk / Superchunk::fill_cell give cell by cell
(chunk.rs:133-150, superchunk.rs:353-400) -- here: the input raster itself, the oracle's raw values, and the per-cell walker.

Run on the B200 box with `pytest -m gpu`.
"""
import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 64, 0])
def ctx(request):
    """cell_tile_min = 1: every tile takes the tile decoder; 64: the default mix; 0: the per-cell walker only."""
    from dcdf_b200 import Context
    c = Context(0)
    c.set_option("cell_tile_min", request.param)
    yield c
    c.close()


def _queries(rng, T, R, C, n, *, full=False):
    rows, cols = rng.integers(0, R, n), rng.integers(0, C, n)
    if full:
        s, e = np.zeros(n, np.int64), np.full(n, T, np.int64)
    else:
        s = rng.integers(0, T, n)
        e = np.minimum(T, s + rng.integers(0, T, n))        # some empty, some crossing slices
    return np.stack([s, e, rows, cols], axis=1).astype(np.int64)


def _check(sc, data, q, raw_ref=None):
    got = sc.cell_batch(q)
    for qi, series in zip(q, got):
        s, e, r, c = (int(x) for x in qi)
        lo, hi = min(s, e), max(s, e)
        assert np.array_equal(series, data[lo:hi, r, c], equal_nan=True), f"series {qi}"
    if raw_ref is not None:
        graw = sc.cell_batch(q, raw=True)
        for qi, series in zip(q, graw):
            s, e, r, c = (int(x) for x in qi)
            assert np.array_equal(series, raw_ref[min(s, e):max(s, e), r, c]), f"raw series {qi}"


def test_dense_series_superchunk_f32(ctx):
    from dcdf_b200 import Superchunk
    rng = np.random.default_rng(11)
    T, R, C = 40, 100, 130
    data = (rng.integers(-40, 40, (T, R, C)) / 8).astype(np.float32)
    data[:, 70:, :20] = 2.5                         # an elided corner tile
    data[3:9, 10:30, 5:25] = np.nan
    sc = Superchunk.build(ctx, data, [2, 6], compute_bits=True, chunk_size=8)
    ref_raw = np.concatenate([orc.superchunk_build(data[t:t + 8], [2, 6], compute_bits=True).window_raw(0, min(8, T - t), 0, R, 0, C)
                              for t in range(0, T, 8)])
    q = _queries(rng, T, R, C, 3000)
    q = np.concatenate([q, q[:200], _queries(rng, T, R, C, 500, full=True)])     # repeated cells, full-length series
    q[5, 0], q[5, 1] = q[5, 1], q[5, 0]                                               # a reversed range
    _check(sc, data, q, ref_raw)
    # every cell of one tile, full length
    rr, cc = np.meshgrid(np.arange(64, 100), np.arange(64, 128), indexing="ij")
    qa = np.stack([np.zeros(rr.size, np.int64), np.full(rr.size, T, np.int64), rr.ravel(), cc.ravel()], axis=1)
    _check(sc, data, qa, ref_raw)
    sc.close()


def test_dense_series_int_and_wide_values(ctx):
    from dcdf_b200 import Superchunk
    rng = np.random.default_rng(12)
    T, R, C = 20, 70, 90
    data = rng.integers(-2**40, 2**40, (T, R, C)).astype(np.int64)     # 64-bit expansion
    sc = Superchunk.build(ctx, data, [1, 6], chunk_size=7)
    _check(sc, data, _queries(rng, T, R, C, 2000))
    sc.close()
    small = rng.integers(-100, 100, (T, R, C)).astype(np.int32)
    sc = Superchunk.build(ctx, small, [1, 6], chunk_size=20)
    _check(sc, small, _queries(rng, T, R, C, 2000))
    sc.close()


def test_dense_series_plain_chunk_and_small_trees(ctx):
    from dcdf_b200 import Chunk
    rng = np.random.default_rng(13)
    for shape in [(30, 64, 64), (12, 22, 8), (9, 4, 4), (9, 2, 2), (5, 3, 1)]:
        data = (rng.integers(0, 30, shape) / 2).astype(np.float32)
        ch = Chunk.build(ctx, data, fractional_bits=1)
        T, R, C = shape
        _check(ch, data, _queries(rng, T, R, C, 400))
        ch.close()


def test_dense_series_nested_superchunk(ctx):
    """Three k2_levels entries (superchunks inside a superchunk, BASELINE configs[4]'s shape of tree)."""
    from dcdf_b200 import Superchunk
    rng = np.random.default_rng(14)
    T, R, C = 24, 200, 230
    data = (rng.integers(0, 64, (T, R, C)) / 4).astype(np.float32)
    data[:, 128:192, 64:128] = 1.25           # an elided leaf
    sc = Superchunk.build(ctx, data, [1, 1, 6], compute_bits=True, chunk_size=8)
    _check(sc, data, _queries(rng, T, R, C, 6000))
    sc.close()
