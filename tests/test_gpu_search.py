"""GPU parity tests of the tile search kernel (k_search_tiles4, decode_search4.cuh) through the C-ABI: result cells AND
their order must equal the reference traversal (Snapshot::search_window snapshot.rs:310-421, Log::search_window
log.rs:519-702, Chunk::iter_search chunk.rs:213-228) as restated by the oracle.  The rasters have large uniform areas,
areas that equal the snapshot plus a constant, smooth gradients (whole sub-trees inside the band: the reference then
pushes a node's rectangle row-major) and per-cell noise, on 64x64 tiles, clipped tiles and small trees."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as orc
from test_gpu_window import _field

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["default", "shared", "per_window"])
def ctx(request):
    """shared: the counting pass always decodes every touched (time slice, tile) once for all windows of the batch
    (search_share_min = 1); per_window: never; default: from three windows per touched tile on."""
    from dcdf_b200 import Context
    c = Context(0)
    if request.param == "shared":
        c.set_option("search_share_min", 1)
    if request.param == "per_window":
        c.set_option("search_share_min", 0)
    yield c
    c.close()


def _bands(fixed, rng, n):
    lo, hi = int(fixed.min()), int(fixed.max())
    out = [(lo, hi), (lo - 5, lo - 1), (hi + 1, hi + 9), (int(np.median(fixed)), int(np.median(fixed)))]
    for _ in range(n):
        a, b = sorted(int(x) for x in rng.integers(lo - 3, hi + 4, 2))
        out.append((a, b))
        c = int(rng.integers(lo, hi + 1))
        out.append((c, c + int(rng.integers(0, max(2, (hi - lo) // 6)))))
    return out


@pytest.mark.parametrize("rows,cols,kind", [(64, 64, "float"), (64, 64, "int"), (50, 64, "float"), (64, 37, "int"), (32, 32, "float"),
                                            (17, 9, "int"), (8, 8, "float"), (4, 4, "int"), (3, 2, "int"), (2, 2, "int")])
def test_chunk_search_order_exact(ctx, rows, cols, kind):
    from dcdf_b200 import Chunk
    T = 12
    rng = np.random.default_rng(rows * 1000 + cols)
    if kind == "float":
        data = _field(T, rows, cols, 77 + rows + cols, nan_frac=0.04 if rows * cols > 16 else 0.0, flat=rows >= 8)
        got, ref = Chunk.build(ctx, data, fractional_bits=4), orc.chunk_build(data, fractional_bits=4)
    else:
        y, x = np.mgrid[0:rows, 0:cols]
        data = np.stack([(100 + 2 * y + x + 3 * t + (rng.integers(-2, 3, (rows, cols)) * (rng.random((rows, cols)) < 0.3))) for t in range(T)]).astype(np.int32)
        data[:, : rows // 2, : cols // 2] = 250                      # uniform quadrant
        data[5] = data[4] + 11                                       # a whole instant "equal + constant"
        data[7] = 999                                                # a single-node instant
        data[8] = 999                                                # single-node log over a single-node snapshot (or its log)
        got, ref = Chunk.build(ctx, data), orc.chunk_build(data)
    assert got.write_to() == ref.serialize()
    fixed = ref.window(0, T, 0, rows, 0, cols)
    wins = [(0, T, 0, rows, 0, cols), (1, T - 1, rows // 4, rows, 0, max(1, cols - 1)), (3, 9, 0, max(1, rows // 2 + 1), cols // 3, cols),
            (T - 1, T, rows - 1, rows, cols - 1, cols)]
    for (a, b, t, bo, l, r) in wins:
        for lo, hi in _bands(fixed, rng, 6):
            g = got.search(a, b, t, bo, l, r, lo, hi)
            o = ref.search(a, b, t, bo, l, r, lo, hi)
            assert np.array_equal(g, o), f"window {(a, b, t, bo, l, r)} band {(lo, hi)}: {len(g)} vs {len(o)} cells or order differs"


def test_superchunk_search_batch_against_oracle_chunks(ctx):
    """Per (window, subchunk) the cells of one instant come out in the chunk's traversal order; subchunks row-major,
    instants ascending inside a subchunk (decode.cuh: job order)."""
    from dcdf_b200 import Superchunk
    T, R, C = 10, 150, 140
    data = _field(T, R, C, 4, nan_frac=0.03)
    data[:, 64:128, 0:64] = 2.25                                     # an elided subchunk
    sc = Superchunk.build(ctx, data, [2, 6], chunk_size=4)
    bits = sc.info(0).fractional_bits
    cubes = [[0, T, 0, R, 0, C], [1, 9, 30, 131, 20, 139], [3, 4, 64, 128, 0, 64], [2, 10, 100, 150, 100, 140]]
    rng = np.random.default_rng(3)
    for lo_f, hi_f in [(2.25, 2.25), (240.0, 251.0), (0.0, 1000.0), (247.0, 247.5)]:
        lo, hi = orc.to_fixed(lo_f, bits, False, np.float32), orc.to_fixed(hi_f, bits, False, np.float32)
        counts, cells = sc.search_batch(cubes, lo, hi)
        pos = 0
        for c, n in zip(cubes, counts):
            mine = cells[pos:pos + int(n)]
            pos += int(n)
            want = []
            for cr in range(c[2] // 64, (c[3] - 1) // 64 + 1):
                for cc in range(c[4] // 64, (c[5] - 1) // 64 + 1):
                    tile = np.ascontiguousarray(data[:, cr * 64:min(cr * 64 + 64, R), cc * 64:min(cc * 64 + 64, C)])
                    t_, b_ = max(c[2], cr * 64) - cr * 64, min(c[3], cr * 64 + 64) - cr * 64
                    l_, r_ = max(c[4], cc * 64) - cc * 64, min(c[5], cc * 64 + 64) - cc * 64
                    for s0 in range(0, T, 4):                          # one Chunk per time slice
                        a, b = max(c[0], s0), min(c[1], s0 + 4)
                        if a >= b:
                            continue
                        ref = orc.chunk_build(np.ascontiguousarray(tile[s0:s0 + 4]), fractional_bits=bits)
                        hits = ref.search(a - s0, b - s0, t_, b_, l_, r_, lo, hi)
                        want += [(int(t) + s0, int(r) + cr * 64, int(cl) + cc * 64) for t, r, cl in hits]
            assert [tuple(x) for x in mine.tolist()] == want, f"cube {c} band {(lo_f, hi_f)}"
    sc.close()


@pytest.mark.parametrize("option", ["search_dfs", "window_wide", "search_no_cache"])
def test_depth_first_search_kernel_and_wide_expansion_stay_bit_exact(option):
    import test_gpu_window as t
    from dcdf_b200 import Chunk, Context
    ctx = Context(0)
    ctx.set_option(option, 1)
    data = t._field(9, 64, 50, 5, nan_frac=0.04)
    got, ref = Chunk.build(ctx, data, fractional_bits=4), orc.chunk_build(data, fractional_bits=4)
    for lo, hi in ((7900, 8000), (0, 0), (7000, 9000)):
        assert np.array_equal(got.search(1, 9, 3, 60, 2, 49, lo, hi), ref.search(1, 9, 3, 60, 2, 49, lo, hi))
    got.close()
    ctx.close()


def test_superchunk_search_prunes_with_the_superchunk_min_max_like_the_reference(ctx):
    """Superchunk::search first asks the superchunk's own min / max Dacs -- in the SUPERCHUNK's fractional bits --
    whether a subchunk can have cells in range (superchunk.rs:480-493) and only then searches the subchunk with the same
    bounds in the SUBCHUNK's bits.  When the two differ the pruning changes the answer, so it has to be reproduced:
    results and order against the oracle's Superchunk::search, two-level and nested."""
    from dcdf_b200 import Superchunk
    rng = np.random.default_rng(77)
    data = (rng.integers(1600, 1920, (9, 128, 192)) / 16.0).astype(np.float32)        # 4 fractional bits
    data[:, :64, 64:128] = (rng.integers(200, 240, (9, 64, 64)) / 2.0).astype(np.float32)   # this subchunk alone: 1 bit
    data[:, 64:, :64] = 107.5                                                           # elided subchunk
    for levels in ([2, 6], [1, 1, 6]):
        got = Superchunk.build(ctx, data, levels)
        ref = orc.superchunk_build(data, levels)
        assert got.info(0).fractional_bits == 4
        cubes = [[0, 9, 0, 128, 0, 192], [2, 7, 10, 100, 50, 150], [0, 9, 64, 128, 0, 64], [3, 4, 0, 64, 64, 128]]
        bands = [(400, 500), (3300, 3500), (3441, 3441), (3201, 3841), (0, 10 ** 6), (430, 431), (6000, 100)]
        for lo, hi in bands:
            counts, cells = got.search_batch(cubes, lo, hi)
            rcounts, rcells, _ = ref.search_batch(cubes, lo, hi)
            assert counts.tolist() == rcounts.tolist(), (levels, lo, hi, counts.tolist(), rcounts.tolist())
            if len(levels) > 2:
                # the reference gathers subchunk streams unordered; the C-ABI's order is leaf subchunks row-major over the
                # whole window, the oracle's recursion goes region by region: bring the oracle's cells into leaf order
                pos, parts = 0, []
                for n in rcounts:
                    w = rcells[pos:pos + int(n)]
                    pos += int(n)
                    parts.append(w[np.argsort((w[:, 1] // 64) * 1000 + w[:, 2] // 64, kind="stable")])
                rcells = np.concatenate(parts) if parts else rcells
            assert np.array_equal(cells, rcells), (levels, lo, hi)
        series, _ = ref.cell_batch([[0, 9, 3, 70], [2, 8, 100, 5], [0, 9, 127, 191]])
        mine = got.cell_batch([[0, 9, 3, 70], [2, 8, 100, 5], [0, 9, 127, 191]], raw=True)
        for a, b in zip(mine, series):
            assert np.array_equal(a, b)
        got.close()


def test_single_node_logs_follow_the_reference_traversal(ctx):
    """SURVEY Appendix B #15 through every counting path: a Log that is one node (the tile is uniform, or equal to its
    snapshot plus a constant, at that instant) is tested by the reference with min_t = 0 at the root and read as
    "snapshot + root entry" below it.  Counts and cells against the oracle for many bands, several windows per call."""
    import fixtures as fx
    from dcdf_b200 import Superchunk
    rng = np.random.default_rng(5)
    a = fx.array8_3()
    frames = [a[0], np.full((8, 8), 5, np.int64), a[0] + 3, a[0] - 2, np.full((8, 8), -4, np.int64), a[1], np.full((8, 8), 9, np.int64), a[1] + 1]
    data = np.stack([np.tile(f, (2, 2)) for f in frames])          # four 8x8 subchunks with the same history
    data[:, 8:, 8:] += 1
    got, ref = Superchunk.build(ctx, data, [1, 3]), orc.superchunk_build(data, [1, 3])
    T = len(frames)
    cubes, los, his = [], [], []
    for lo in range(-8, 14, 2):
        for width in (0, 1, 3, 20):
            t0 = int(rng.integers(0, T - 1))
            r0, c0 = int(rng.integers(0, 14)), int(rng.integers(0, 14))
            cubes.append([t0, int(rng.integers(t0 + 1, T + 1)), r0, int(rng.integers(r0 + 1, 17)), c0, int(rng.integers(c0 + 1, 17))])
            los.append(lo); his.append(lo + width)
    counts, cells = got.search_batch(cubes, los, his)
    rcounts, rcells, _ = ref.search_batch(cubes, los, his)
    assert counts.tolist() == rcounts.tolist()
    assert np.array_equal(cells, rcells)
    got.close()


def test_logs_over_a_single_node_snapshot_follow_the_reference_root_test(ctx):
    """A block whose Snapshot is one node (the tile is uniform at that instant): the reference tests the root of every Log of
    the block with snapshot.min.get(0) + log.min.get(0), and the empty min Dac of a single-node Snapshot reads 0
    (log.rs:527-548, dac.rs:80-93) -- bands below the wrong bound find nothing, bands above it may find everything.  Counts
    and cells against the oracle in every counting mode."""
    import fixtures as fx
    from dcdf_b200 import Superchunk
    rng = np.random.default_rng(6)
    a = fx.array8_3()
    frames = [np.full((8, 8), -13, np.int64), a[0] - 40, a[0] - 41, a[1] - 40, np.full((8, 8), 7, np.int64), a[0] - 39, a[1] + 50, a[0] - 40]
    data = np.stack([np.tile(f, (2, 2)) for f in frames])
    data[:, 8:, :8] -= 3
    got, ref = Superchunk.build(ctx, data, [1, 3]), orc.superchunk_build(data, [1, 3])
    T = len(frames)
    cubes, los, his = [], [], []
    for lo in range(-60, 70, 5):
        for width in (0, 4, 12, 200):
            t0 = int(rng.integers(0, T - 1))
            r0, c0 = int(rng.integers(0, 14)), int(rng.integers(0, 14))
            cubes.append([t0, int(rng.integers(t0 + 1, T + 1)), r0, int(rng.integers(r0 + 1, 17)), c0, int(rng.integers(c0 + 1, 17))])
            los.append(lo); his.append(lo + width)
    cubes.append([0, T, 0, 16, 0, 16]); los.append(-30); his.append(1000)
    cubes.append([0, T, 0, 16, 0, 16]); los.append(-1000); his.append(-36)     # everything below a bound
    counts, cells = got.search_batch(cubes, los, his)
    only_counts, _ = got.search_batch(cubes, los, his, want_cells=False)
    rcounts, rcells, _ = ref.search_batch(cubes, los, his)
    assert counts.tolist() == rcounts.tolist()
    assert only_counts.tolist() == rcounts.tolist()
    assert np.array_equal(cells, rcells)
    got.close()
