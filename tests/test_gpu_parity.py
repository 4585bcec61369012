"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle, bit for bit.

Run on the B200 box with `pytest -m gpu`.  Nothing here reads /root/reference.
"""
import itertools
import os

import numpy as np
import pytest

import fixtures as fx
import oracle_lib as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from dcdf_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _diff(a, b, what):
    """Readable first difference of two byte strings."""
    if a == b:
        return
    n = min(len(a), len(b))
    i = next((j for j in range(n) if a[j] != b[j]), n)
    raise AssertionError(f"{what}: bytes differ at offset {i} (gpu len {len(a)}, oracle len {len(b)}): "
                         f"gpu {a[i:i+16].hex()} vs oracle {b[i:i+16].hex()}")


def _check_chunk(ctx, data, fractional_bits=0, round_=False, queries=True):
    from dcdf_b200 import Chunk
    ref = orc.chunk_build(data, fractional_bits=fractional_bits, round_=round_)
    got = Chunk.build(ctx, data, fractional_bits=fractional_bits, round=round_)
    assert got.block_instants() == ref.block_instants(), "heuristic decisions differ"
    _diff(got.write_to(), ref.serialize(), f"chunk {data.shape} {data.dtype}")
    assert got.stats["size"] == ref.stats["size"] == got.size()
    assert (got.stats["snapshots"], got.stats["logs"]) == (ref.stats["snapshots"], ref.stats["logs"])
    info = ref.info()
    assert got.shape == info["shape"] and got.encoding == info["encoding"] and got.fractional_bits == info["fractional_bits"]
    if queries:
        T, R, Cc = data.shape
        w = got.window(0, T, 0, R, 0, Cc, raw=True)
        assert np.array_equal(w, ref.window(0, T, 0, R, 0, Cc))
    return got, ref


# ----------------------------------------------------------------------------- encode: Chunk::build
def test_chunk_build_reference_fixtures(ctx):
    _check_chunk(ctx, fx.array8(100))                       # testing.rs:200-240 (i64)
    _check_chunk(ctx, fx.array8(100, np.int32))
    _check_chunk(ctx, fx.array(16, 30))
    _check_chunk(ctx, fx.array(64, 7))
    _check_chunk(ctx, fx.farray8(100), fractional_bits=3)   # testing.rs:251-330 (f32 + NaN)
    _check_chunk(ctx, fx.farray(16, 20, np.float64), fractional_bits=3)
    _check_chunk(ctx, fx.farray(64, 6), fractional_bits=3)


@pytest.mark.parametrize("rows,cols", [(2, 2), (3, 2), (4, 4), (5, 7), (8, 8), (9, 9), (16, 16), (17, 23), (33, 20), (64, 64), (50, 64), (64, 37), (1, 40), (40, 1)])
def test_chunk_build_shapes_and_padding(ctx, rows, cols):
    rng = np.random.default_rng(rows * 100 + cols)
    base = rng.integers(0, 40, (rows, cols))
    frames = [base]
    for i in range(9):
        nxt = frames[-1].copy()
        if i % 4 == 3:
            nxt = rng.integers(0, 5000, (rows, cols))          # unrelated frame -> new snapshot
        else:
            m = rng.random((rows, cols)) < 0.1
            nxt[m] += rng.integers(-3, 4, m.sum())
        frames.append(nxt)
    data = np.stack(frames).astype(np.int64)
    _check_chunk(ctx, data)
    f = (data / 8.0).astype(np.float32)
    f[rng.random(f.shape) < 0.05] = np.nan
    _check_chunk(ctx, f, fractional_bits=3)


def test_chunk_build_wide_values_and_negatives(ctx):
    rng = np.random.default_rng(5)
    data = rng.integers(-2 ** 45, 2 ** 45, (6, 40, 40)).astype(np.int64)      # 6-7 byte DAC codes, int64 path
    data[3] = data[2]
    _check_chunk(ctx, data)
    data = rng.integers(-2 ** 31, 2 ** 31 - 1, (5, 32, 32)).astype(np.int32)
    _check_chunk(ctx, data)
    f = rng.normal(0, 1000, (6, 33, 47)).astype(np.float32)                   # needs ~20+ fractional bits
    kind, bits = orc.suggest_fraction(f)
    _check_chunk(ctx, f, fractional_bits=bits, round_=(kind == "Round"))
    f64 = rng.normal(0, 10, (4, 16, 16))
    _check_chunk(ctx, f64, fractional_bits=20, round_=True)                   # rounding branch fixed.rs:48-49
    _check_chunk(ctx, f.astype(np.float32), fractional_bits=8, round_=True)


def test_chunk_heuristic_and_log_cap(ctx):
    rng = np.random.default_rng(3)
    base = rng.integers(0, 50, (16, 16)).astype(np.int64)
    data = np.stack([base] * 5 + [rng.integers(1000, 2000, (16, 16)).astype(np.int64)] * 3)
    got, ref = _check_chunk(ctx, data)
    assert got.block_instants() == [5, 3]
    got, ref = _check_chunk(ctx, np.stack([base] * 300), queries=False)        # 254-log cap chunk.rs:62
    assert got.block_instants() == [255, 45]
    uniform = np.full((4, 32, 32), 42, np.int64)                               # single-node trees
    _check_chunk(ctx, uniform)
    nan = np.full((3, 8, 8), np.nan, np.float32)                               # all NaN -> all zeros
    _check_chunk(ctx, nan, fractional_bits=0)


def test_chunk_build_strided_view_and_device_input(ctx):
    import torch
    from dcdf_b200 import Chunk
    big = fx.farray(64, 12)
    view = big[2:11, 5:50, 3:60:1]
    ref = orc.chunk_build(np.ascontiguousarray(view), fractional_bits=3)
    got = Chunk.build(ctx, view, fractional_bits=3)
    _diff(got.write_to(), ref.serialize(), "strided host view")
    dev = torch.from_numpy(big).cuda()
    got = Chunk.build(ctx, dev[2:11, 5:50, 3:60], fractional_bits=3)
    _diff(got.write_to(), ref.serialize(), "strided device view")
    # reversed ndarray views have negative strides (mmbuffer.rs:573-594 slices any ArrayViewMut3)
    for rev in (np.s_[::-1], np.s_[:, ::-1], np.s_[:, :, ::-1], np.s_[::-1, ::-2, ::-1]):
        v = big[rev]
        ref = orc.chunk_build(np.ascontiguousarray(v), fractional_bits=3)
        _diff(Chunk.build(ctx, v, fractional_bits=3).write_to(), ref.serialize(), f"negative strides {rev}")
    from dcdf_b200 import Superchunk, synth
    field = synth.raster_slice(0, 9, 150, 200).numpy()
    for rev in (np.s_[::-1], np.s_[:, ::-1, ::-1]):
        v = field[rev]
        sc = Superchunk.build(ctx, v, [2, 6])
        refs = orc.superchunk_build(np.ascontiguousarray(v), [2, 6])
        kinds, child = refs.node_refs(0)
        chunks = sc.chunk_bytes(0)
        for slot, c in enumerate(child):
            if kinds[slot]:
                _diff(chunks[slot], refs.node_bytes(int(c)), f"superchunk over a reversed view, slot {slot}")
        assert np.array_equal(sc.window(0, 9, 0, 150, 0, 200), v)
        sc.close()


@pytest.mark.parametrize("shape", [(5, 65, 65), (4, 70, 130), (3, 200, 300), (6, 128, 128), (3, 1, 100)])
def test_big_chunk_build_multi_level(ctx, shape):
    """Chunk::build for padded sides above 64 (dense pyramid + device-wide scans path)."""
    rng = np.random.default_rng(shape[1])
    base = rng.integers(0, 30, shape[1:])
    frames = []
    for i in range(shape[0]):
        f = base.copy()
        m = rng.random(shape[1:]) < 0.02 * i
        f[m] += rng.integers(-300, 300, m.sum())
        frames.append(f if i != 2 else rng.integers(-10 ** 6, 10 ** 6, shape[1:]))
    data = np.stack(frames).astype(np.int64)
    got, ref = _check_chunk(ctx, data)
    f32 = (data / 16.0).astype(np.float32)
    f32[rng.random(f32.shape) < 0.1] = np.nan
    got, ref = _check_chunk(ctx, f32, fractional_bits=4)
    T, R, Cc = shape
    g = got.search(0, T, R // 3, R, 0, Cc // 2 + 1, -100, 100)
    o = ref.search(0, T, R // 3, R, 0, Cc // 2 + 1, -100, 100)
    assert np.array_equal(g, o)


def test_c1_config_chunk_build_and_get_window(ctx):
    """BASELINE configs[0]: synthetic 256x256x100 f32 raster, Chunk build (Snapshot + Log) then get_window."""
    from dcdf_b200 import synth
    data = synth.raster_slice(0, 100, 256, 256).numpy()
    got, ref = _check_chunk(ctx, data, fractional_bits=4)
    assert len(got.block_instants()) > 1 and max(got.block_instants()) > 1    # both snapshots and logs
    w = got.window(10, 90, 17, 250, 3, 256)
    assert np.array_equal(w, data[10:90, 17:250, 3:256])
    series = got.cell(0, 100, 255, 0)
    assert np.array_equal(series, data[:, 255, 0])


@pytest.mark.parametrize("bad,code", [("precision", 2), ("inf", 1), ("overflow", 3)])
def test_chunk_build_data_errors(ctx, bad, code):
    from dcdf_b200 import Chunk, DcdfError
    data = fx.farray(16, 4).copy()
    bits = 3
    if bad == "precision":
        bits = 2                                    # fixed.rs:51-57
    elif bad == "inf":
        data[2, 3, 4] = np.inf                      # fixed.rs:39-41
    else:
        data[1, 0, 0] = 3e38
        bits = 40                                   # fixed.rs:66-69
    with pytest.raises(orc.OracleError) as eo:
        orc.chunk_build(data, fractional_bits=bits)
    with pytest.raises(DcdfError) as eg:
        Chunk.build(ctx, data, fractional_bits=bits)
    assert eg.value.code == eo.value.code == code


def test_chunk_build_rejects_what_the_reference_cannot_do(ctx):
    from dcdf_b200 import Chunk, DcdfError
    with pytest.raises(DcdfError) as e:
        Chunk.build(ctx, np.zeros((3, 1, 1), np.int64))       # snapshot.rs:166 would panic
    assert e.value.code == 8
    with pytest.raises(DcdfError) as e:
        Chunk.build(ctx, fx.array8(3), k=3)
    assert e.value.code == 8


# ----------------------------------------------------------------------------- a1 / a2 / a3
def test_fraction_minmax_fixed(ctx):
    assert ctx.suggest_fraction(fx.fixed_array()) == ("Precise", 3)                      # fixed.rs:311-323
    assert ctx.suggest_fraction(np.array([[[16.0, 1.0 / 16.0]]])) == ("Precise", 4)
    assert ctx.suggest_fraction(np.array([[[16.0, 0.1]]])) == ("Precise", 55)
    assert ctx.suggest_fraction(np.array([[[316.0, 0.1]]])) == ("Round", 53)
    nan = np.float32("nan")
    assert ctx.suggest_fraction(np.array([[[nan, 16.0, nan, 1.0 / 16.0]]], dtype=np.float32)) == ("Precise", 4)
    assert ctx.suggest_fraction(np.array([[[nan, nan]]], dtype=np.float32)) == ("Precise", 0)
    assert ctx.suggest_fraction(np.array([[[-100.5, 2.0]]])) == ("Precise", 0)            # saturating-cast path
    assert ctx.suggest_fraction(np.array([[[-1.5, 2.0]]])) == ("Precise", 1)
    rng = np.random.default_rng(11)
    for dt in (np.float32, np.float64):
        a = (rng.integers(-4000, 9000, (7, 70, 130)) / 64.0).astype(dt)
        a[rng.random(a.shape) < 0.3] = np.nan
        assert ctx.suggest_fraction(a) == orc.suggest_fraction(a)
        assert ctx.suggest_fraction(a[:, 3:40, 5:77]) == orc.suggest_fraction(a[:, 3:40, 5:77])
        mn, mx = ctx.min_max(a, 6)
        rmn, rmx = orc.min_max(a, 6)
        assert np.array_equal(mn, rmn) and np.array_equal(mx, rmx)
    b = rng.normal(0, 50, (3, 20, 20)).astype(np.float32)
    assert ctx.suggest_fraction(b) == orc.suggest_fraction(b)
    v = np.array([1.5, -1.5, 0.0625, 0.0, -0.0, np.nan, 0.1], np.float64)
    assert ctx.to_fixed(v, 16, True).tolist() == [orc.to_fixed(x, 16, True) for x in v]
    assert ctx.to_fixed(np.array([1.5, -1.5], np.float32), 1).tolist() == [7, -5]        # fixed.rs:209-223
    f = ctx.from_fixed(np.array([7, -5, 3, 1, 0]), 1)
    assert f[:4].tolist() == [1.5, -1.5, 0.5, 0.0] and np.isnan(f[4])


# ----------------------------------------------------------------------------- decode: Chunk queries
def test_chunk_queries_against_oracle(ctx):
    from dcdf_b200 import Chunk
    data = fx.array8(100)
    ref = orc.chunk_build(data)
    got = Chunk.read_from(ctx, ref.serialize())                    # oracle bytes -> GPU decoder
    assert got.shape == (100, 8, 8) and got.block_instants() == ref.block_instants()
    irc = np.array(list(itertools.product(range(100), range(8), range(8))))
    assert np.array_equal(got.get_batch(irc), data.reshape(-1))    # chunk.rs:427-439
    qs = [[r * 6, 100 - c * 6, r, c] for r in range(8) for c in range(8)]
    for q, series in zip(qs, got.cell_batch(qs)):                  # chunk.rs:441-456
        assert np.array_equal(series, data[q[0]:q[1], q[2], q[3]])
    for top, bottom, left, right in [(0, 8, 0, 8), (1, 3, 2, 7), (5, 6, 0, 8), (7, 8, 7, 8), (6, 2, 7, 1)]:
        start, end = top * left % 50, 60 + bottom * right % 40
        t, b, l, r = min(top, bottom), max(top, bottom), min(left, right), max(left, right)
        assert np.array_equal(got.window(start, end, top, bottom, left, right), data[start:end, t:b, l:r])
        for lower, upper in ((4, 6), (7, 5), (9, 9), (0, 100)):
            g = got.search(start, end, top, bottom, left, right, lower, upper)
            o = ref.search(start, end, top, bottom, left, right, lower, upper)
            assert np.array_equal(g, o), "search results / order differ from the reference traversal"


def test_float_chunk_decode_and_format_errors(ctx):
    from dcdf_b200 import Chunk, DcdfError
    data = fx.farray(16, 12)
    got = Chunk.build(ctx, data, fractional_bits=3)
    w = got.window(0, 12, 0, 16, 0, 16)
    assert w.dtype == np.float32 and np.array_equal(w, data, equal_nan=True)
    assert np.array_equal(got.window(0, 12, 0, 16, 0, 16, raw=True), orc.chunk_build(data, fractional_bits=3).window(0, 12, 0, 16, 0, 16))
    with pytest.raises(DcdfError) as e:
        got.window(0, 13, 0, 16, 0, 16)                            # mmarray.rs:218-229
    assert e.value.code == 5
    ser = got.write_to()
    for bad in (ser[:-3], b"\x07" + ser[1:], ser + b"\x00"):
        with pytest.raises(DcdfError) as e:
            Chunk.read_from(ctx, bad)
        assert e.value.code == 6


def test_log_search_reference_bug_parity(ctx):
    """results must match the reference (and its oracle), including the quirk."""
    from dcdf_b200 import Chunk
    a = fx.array8_3()
    data = np.stack([a[0], np.full((8, 8), 5, np.int64)])
    got = Chunk.build(ctx, data)
    ref = orc.chunk_build(data)
    assert got.block_instants() == ref.block_instants()
    for lower, upper in ((5, 5), (3, 4), (0, 9)):
        assert np.array_equal(got.search(0, 2, 0, 8, 0, 8, lower, upper), ref.search(0, 2, 0, 8, 0, 8, lower, upper))


# ----------------------------------------------------------------------------- Superchunk
def _cmp_node(got, s, gnode, ref, rnode, blob, path="root"):
    """Walk the GPU node tree and the oracle's tree in lock step."""
    rinfo, ginfo = ref.node_info(rnode), got.info(s, gnode)
    for f in ("sidelen", "chunks_sidelen", "subsidelen", "levels", "fractional_bits", "encoding", "n_refs"):
        assert getattr(ginfo, f) == getattr(rinfo, f), (path, f, getattr(ginfo, f), getattr(rinfo, f))
    assert tuple(ginfo.shape) == tuple(rinfo.shape), path
    rk, rchild = ref.node_refs(rnode)
    gk, goff, gsize, gbits, gchild = got.refs(s, gnode, with_children=True)
    assert gk.tolist() == rk.tolist(), f"{path}: Elided / External pattern differs"
    chunks = got.chunk_bytes(s, gnode, blob)
    for slot, (k, child) in enumerate(zip(rk, rchild)):
        if k == 0:
            assert chunks[slot] is None and gchild[slot] == -1
            continue
        cinfo = ref.node_info(int(child))
        if cinfo.kind == 1:
            _diff(chunks[slot], ref.node_bytes(int(child)), f"slice {s} {path} slot {slot}")
            assert gbits[slot] == cinfo.fractional_bits and gchild[slot] == -1
        else:
            assert gchild[slot] >= 0
            _cmp_node(got, s, int(gchild[slot]), ref, int(child), blob, f"{path}/{slot}")
    _diff(got.bytes(s, 1, gnode), ref.node_bytes(rnode, 1), f"slice {s} {path} max DAC")
    _diff(got.bytes(s, 2, gnode), ref.node_bytes(rnode, 2), f"slice {s} {path} min DAC")
    st, rs = ginfo.stats, rinfo.stats
    assert (st.elided, st.external, st.snapshots, st.logs, st.size) == (rs.elided, rs.external, rs.snapshots, rs.logs, rs.size), path


def _check_superchunk(ctx, data, levels, chunk_size=0, fractional_bits=0, round_=False, compute_bits=True):
    from dcdf_b200 import Superchunk
    got = Superchunk.build(ctx, data, levels, fractional_bits=fractional_bits, round=round_, compute_bits=compute_bits, chunk_size=chunk_size)
    T = data.shape[0]
    cs = chunk_size or T
    assert got.n_slices == (T + cs - 1) // cs
    for s in range(got.n_slices):
        sub = data[s * cs:(s + 1) * cs]
        ref = orc.superchunk_build(sub, levels, fractional_bits=fractional_bits, round_=round_, compute_bits=compute_bits)
        _cmp_node(got, s, 0, ref, 0, got.bytes(s, 0))
    return got


def test_superchunk_reference_structures(ctx):
    _check_superchunk(ctx, fx.array8(20), [3, 0])                 # superchunk.rs:1005-1019: all from the DAC
    _check_superchunk(ctx, fx.array(16, 20), [2, 2])              # superchunk.rs:1068-1093
    _check_superchunk(ctx, fx.array(17, 20), [2, 3])              # superchunk.rs:1097-1131: 8 external / 8 elided
    _check_superchunk(ctx, np.zeros((10, 16, 16), np.int64) + 42, [2, 2])   # elide everything :1135-1173
    _check_superchunk(ctx, fx.farray(32, 13), [2, 3], chunk_size=5)
    _check_superchunk(ctx, fx.farray(16, 9, np.float64), [1, 3], chunk_size=4)
    _check_superchunk(ctx, fx.array(40, 6, np.int32), [3, 3])


def test_nested_superchunks(ctx):
    """Superchunk::build recursion (superchunk.rs:171): more than two k2_levels entries."""
    from dcdf_b200 import synth
    _check_superchunk(ctx, fx.farray(32, 9), [1, 2, 2])                       # superchunk.rs:1177-1196
    data = synth.raster_slice(0, 11, 150, 200).numpy()
    got = _check_superchunk(ctx, data, [1, 1, 6], chunk_size=4)
    assert got.node_count() == 5
    assert np.array_equal(got.window(1, 10, 5, 150, 60, 200), data[1:10, 5:150, 60:200])
    assert np.array_equal(got.cell(0, 11, 140, 190), data[:, 140, 190])
    _check_superchunk(ctx, fx.array(64, 5), [1, 2, 3])
    # a clipped corner region small enough to be demoted to a plain Chunk (superchunk.rs:153-163)
    d2 = synth.raster_slice(0, 5, 130, 130).numpy()
    got = _check_superchunk(ctx, d2, [1, 1, 6])
    assert np.array_equal(got.window(0, 5, 0, 130, 0, 130), d2)
    # regions elided at the upper level: values must come from the upper node's table
    d3 = np.full((6, 200, 200), 7.5, np.float32)
    d3[:, :100, :100] = synth.raster_slice(0, 6, 100, 100).numpy()
    got = _check_superchunk(ctx, d3, [1, 1, 6])
    assert np.array_equal(got.window(0, 6, 0, 200, 0, 200), d3)
    assert got.get(3, 199, 199) == 7.5


def test_nested_bad_levels_inside_recursion(ctx):
    from dcdf_b200 import DcdfError, Superchunk
    data = fx.array(84, 4)       # region (1,1) is 20x20: needs 5 levels, neither <= 2 (demotion) nor == 2 + 4
    with pytest.raises(orc.OracleError) as eo:
        orc.superchunk_build(data, [1, 2, 4])
    with pytest.raises(DcdfError) as eg:
        Superchunk.build(ctx, data, [1, 2, 4])
    assert eg.value.code == eo.value.code == 4


def test_superchunk_bad_levels(ctx):
    from dcdf_b200 import DcdfError, Superchunk
    with pytest.raises(DcdfError) as e:
        Superchunk.build(ctx, fx.array(16, 4), [2, 3])            # superchunk.rs:105-110
    assert e.value.code == 4


def test_superchunk_cpc_fixture_and_queries(ctx):
    p = os.path.join(os.path.dirname(__file__), "golden", "cpc_day_360x720.npz")
    day = np.load(p)["day"]
    data = np.stack([day, day, np.roll(day, 3, axis=1)])
    got = _check_superchunk(ctx, data, [4, 6])                    # Precise(29): 64-bit path, 64% NaN
    assert got.info(0).fractional_bits == 29
    w = got.window(0, 3, 0, 360, 0, 720)
    assert np.array_equal(w, data, equal_nan=True)                # py-dcdf/tests/test_dcdf.py:357-365
    ref = orc.superchunk_build(data, [4, 6])
    rng = np.random.default_rng(1)
    irc = np.stack([rng.integers(0, 3, 500), rng.integers(0, 360, 500), rng.integers(0, 720, 500)], axis=1)
    vals = got.get_batch(irc)
    assert np.array_equal(vals, data[irc[:, 0], irc[:, 1], irc[:, 2]], equal_nan=True)
    fixed, bits = ref.get_batch(irc)
    mine = got.get_batch(irc, raw=True)
    assert np.array_equal(mine, fixed)


def test_superchunk_time_slices_queries_and_search(ctx):
    from dcdf_b200 import synth
    data = synth.raster_slice(0, 23, 150, 200).numpy()
    got = _check_superchunk(ctx, data, [2, 6], chunk_size=8)
    assert np.array_equal(got.window(3, 21, 10, 140, 7, 199), data[3:21, 10:140, 7:199])
    series = got.cell_batch([[0, 23, 149, 199], [5, 17, 64, 64], [22, 2, 0, 63]])
    assert np.array_equal(series[0], data[:, 149, 199]) and np.array_equal(series[1], data[5:17, 64, 64])
    assert np.array_equal(series[2], data[2:22, 0, 63])
    cubes = [[0, 23, 0, 150, 0, 200], [4, 9, 60, 70, 120, 135], [7, 8, 0, 1, 0, 1]]
    out, off = got.window_batch(cubes)
    for c, a, b in zip(cubes, off[:-1], off[1:]):
        assert np.array_equal(out[int(a):int(b)].reshape(c[1] - c[0], c[3] - c[2], c[5] - c[4]), data[c[0]:c[1], c[2]:c[3], c[4]:c[5]])
    # search: per slice the reference answers at chunk level; compare as sets against brute force and
    # exactly (order included) against the oracle chunk by chunk
    bits = got.info(0).fractional_bits
    fixed = np.vectorize(lambda v: orc.to_fixed(v, bits, False, np.float32))(data)
    lo, hi = int(np.percentile(fixed, 40)), int(np.percentile(fixed, 45))
    counts, cells = got.search_batch(cubes, lo, hi)
    pos = 0
    for c, n in zip(cubes, counts):
        mine = cells[pos:pos + int(n)]
        pos += int(n)
        assert {tuple(x) for x in mine.tolist()} == fx.brute_search3(fixed, c[0], c[1], c[2], c[3], c[4], c[5], lo, hi)
        assert len(mine) == len({tuple(x) for x in mine.tolist()})


# ----------------------------------------------------------------------------- full-size properties
def test_c2_shaped_slice_round_trip(ctx):
    """ERA5-shaped grid (721 x 1440), one 64-instant slice generated on the device: encode -> decode every
    cell -> equals the input; and encoding is deterministic (a checksum of the chunk bytes repeats)."""
    import torch
    from dcdf_b200 import Superchunk, synth
    dev = synth.raster_slice(0, 64, 721, 1440, device="cuda")
    sc = Superchunk.build(ctx, dev, [5, 6])
    info = sc.info(0)
    assert info.fractional_bits == 4 and info.stats.external == 12 * 23
    out = torch.empty_like(dev)
    sc.window(0, 64, 0, 721, 0, 1440, out=out)
    assert torch.equal(out, dev)
    blob = sc.bytes(0, 0)
    sc2 = Superchunk.build(ctx, dev, [5, 6])
    assert sc2.bytes(0, 0) == blob and sc2.bytes(0, 1) == sc.bytes(0, 1)
    # spot-check three subchunks byte-for-byte against the oracle (corner, edge, interior)
    host = dev.cpu().numpy()
    chunks = sc.chunk_bytes(0)
    for (r, c) in ((0, 0), (11, 22), (5, 9)):
        tile = np.ascontiguousarray(host[:, r * 64:(r + 1) * 64, c * 64:(c + 1) * 64])
        kind, bits = orc.suggest_fraction(tile)
        ref = orc.chunk_build(tile, fractional_bits=bits)
        _diff(chunks[r * 32 + c], ref.serialize(), f"subchunk ({r},{c})")


def test_c5_shaped_nested_round_trip(ctx):
    """0.1-degree global grid (1801 x 3600), k2_levels [2,4,6] as in examples/example.py:201: 16 nested
    superchunks over 29 x 57 leaf subchunks.  Encode -> decode == input; three subchunks vs the oracle."""
    import torch
    from dcdf_b200 import Superchunk, synth
    dev = synth.raster_slice(0, 8, 1801, 3600, device="cuda")
    sc = Superchunk.build(ctx, dev, [2, 4, 6])
    assert sc.node_count() == 1 + 8                      # 2 x 4 in-bounds 1024-side regions of the 4 x 4 root grid
    info = sc.info(0)
    assert (info.sidelen, info.chunks_sidelen, info.subsidelen) == (4096, 1024, 4)
    out = torch.empty_like(dev)
    sc.window(0, 8, 0, 1801, 0, 3600, out=out)
    assert torch.equal(out, dev)
    host = dev.cpu().numpy()
    kinds, off, size, bits, child = sc.refs(0, 0, with_children=True)
    assert kinds.tolist()[:4] == [2, 2, 2, 2] and kinds.tolist()[8:] == [0] * 8
    blob = sc.bytes(0, 0)
    node = int(child[5])                                 # region row 1, col 1: rows 1024..1801, cols 1024..2048
    ninfo = sc.info(0, node)
    assert tuple(ninfo.shape) == (8, 777, 1024) and ninfo.subsidelen == 16
    chunks = sc.chunk_bytes(0, node, blob)
    for (r, c) in ((0, 0), (12, 15), (7, 3)):
        tile = np.ascontiguousarray(host[:, 1024 + r * 64:1024 + (r + 1) * 64, 1024 + c * 64:1024 + (c + 1) * 64])
        kind, b = orc.suggest_fraction(tile)
        _diff(chunks[r * 16 + c], orc.chunk_build(tile, fractional_bits=b).serialize(), f"nested subchunk ({r},{c})")
