"""Storage side of the boundary on the GPU: content addressing + node assembly of built superchunks
(Superchunk::build's tail + Resolver::save) against the oracle's restatement, and opening superchunks from STORED bytes
(Superchunk::load_from) so that every batched query runs on data this process did not build.

Run on the B200 box with `pytest -m gpu`.  Nothing here reads /root/reference.
"""
import hashlib

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


def _check_save(ctx, data, levels, chunk_size=0, **kw):
    from dcdf_b200 import Superchunk
    got = Superchunk.build(ctx, data, levels, chunk_size=chunk_size, **kw)
    T = data.shape[0]
    cs = chunk_size or T
    saved = []
    for s in range(got.n_slices):
        ref = orc.superchunk_build(data[s * cs:(s + 1) * cs], levels, **{("round_" if k == "round" else k): v for k, v in kw.items()})
        rs = ref.save()
        rnodes, rstats = rs.nodes(), rs.stats()
        nodes, stats = got.save(s)
        assert [(c.hex(), t, len(b)) for c, t, b in nodes] == [(c.hex(), t, len(b)) for c, t, b in rnodes], f"slice {s}: stored objects differ"
        for (c, t, b), (rc, rt, rb) in zip(nodes, rnodes):
            assert b == rb, f"slice {s}: bytes of node {c.hex()[:16]} (type {t}) differ"
            assert c == bytes([1, 0x12, 0x12, 0x20]) + hashlib.sha256(b).digest()          # testing.rs:172-177
        assert {k: stats[k] for k in rstats} == rstats, (stats, rstats)
        saved.append((nodes, ref))
    return got, saved


def test_saved_nodes_and_cids_match_the_oracle(ctx):
    from dcdf_b200 import synth
    got, saved = _check_save(ctx, fx.array(16, 100), [2, 2])                 # superchunk.rs:1068-1093
    nodes, _ = saved[0]
    assert sum(1 for _, t, _ in nodes if t == 4) == 4                        # 16 External references, 4 distinct subchunks
    got.close()
    got, saved = _check_save(ctx, fx.array(17, 100), [2, 3])                 # superchunk.rs:1097-1131: 8 external / 8 elided
    got.close()
    _check_save(ctx, np.zeros((10, 16, 16), np.int64) + 42, [2, 2])[0].close()   # elide everything :1135-1173
    _check_save(ctx, fx.farray(32, 9), [1, 2, 2])[0].close()                 # nested :1177-1196
    data = synth.raster_slice(0, 21, 150, 200).numpy()
    _check_save(ctx, data, [2, 6], chunk_size=8)[0].close()
    _check_save(ctx, data, [1, 1, 6])[0].close()
    noisy = synth.raster_slice(0, 9, 130, 140, nan_ocean=True, scale=0.1).numpy()
    _check_save(ctx, noisy, [2, 6])[0].close()                               # int64 path chunks
    _check_save(ctx, noisy, [2, 6], fractional_bits=7, round=True)[0].close()


def _query_checks(sc, data, ref_slices, cs):
    T, R, Cc = data.shape
    w = sc.window(0, T, 0, R, 0, Cc)
    assert np.array_equal(w, data, equal_nan=True)
    rng = np.random.default_rng(3)
    q = np.stack([np.zeros(40, np.int64), np.full(40, T, np.int64), rng.integers(0, R, 40), rng.integers(0, Cc, 40)], axis=1)
    for qi, series in zip(q, sc.cell_batch(q)):
        assert np.array_equal(series, data[:, qi[2], qi[3]], equal_nan=True)
    irc = np.stack([rng.integers(0, T, 300), rng.integers(0, R, 300), rng.integers(0, Cc, 300)], axis=1)
    assert np.array_equal(sc.get_batch(irc), data[irc[:, 0], irc[:, 1], irc[:, 2]], equal_nan=True)
    for s, ref in enumerate(ref_slices):
        t0, t1 = s * cs, min((s + 1) * cs, T)
        raw = ref.window_raw(0, t1 - t0, 0, R, 0, Cc)
        lo, hi = int(np.percentile(raw, 30)), int(np.percentile(raw, 40))
        cubes_local = [[0, t1 - t0, 0, R, 0, Cc], [0, t1 - t0, R // 3, R - 1, 5, Cc // 2]]
        rcounts, rcells, _ = ref.search_batch(cubes_local, lo, hi)
        cubes = [[t0 + a, t0 + b, c, d, e, f] for a, b, c, d, e, f in cubes_local]
        counts, cells = sc.search_batch(cubes, lo, hi)
        assert counts.tolist() == rcounts.tolist()
        if rcells is not None and len(rcells):
            rcells = rcells.copy()
            rcells[:, 0] += t0
            # nested superchunks: the oracle's recursion goes region by region, the C-ABI orders leaf subchunks row-major
            # over the whole window (the reference gathers subchunk streams unordered, superchunk.rs:500-513)
            pos, parts = 0, []
            for n in rcounts:
                w = rcells[pos:pos + int(n)]
                pos += int(n)
                parts.append(w[np.argsort((w[:, 1] // 64) * 1000 + w[:, 2] // 64, kind="stable")])
            rcells = np.concatenate(parts)
        assert np.array_equal(cells, rcells)


def test_open_from_stored_bytes_and_query(ctx):
    """build -> save -> free -> open from the stored objects -> every query equals the input / the oracle."""
    from dcdf_b200 import Superchunk, synth
    data = synth.raster_slice(0, 21, 150, 200).numpy()
    data[:, :64, 64:128] = 3.5                                               # an Elided subchunk
    cs = 8
    got, saved = _check_save(ctx, data, [2, 6], chunk_size=cs)
    store, roots = {}, []
    for nodes, _ in saved:
        for c, t, b in nodes:
            store[c] = b
        roots.append(nodes[-1][0])
    refs = [ref for _, ref in saved]
    got.close()
    sc = Superchunk.open(ctx, roots, store)
    assert sc.shape == data.shape and sc.n_slices == 3
    _query_checks(sc, data, refs, cs)
    nodes2, _ = sc.save(1)                                                   # an opened superchunk saves to the same objects
    assert [(c, b) for c, _, b in nodes2] == [(c, b) for c, _, b in saved[1][0]]
    sc.close()


def test_open_objects_written_by_the_oracle(ctx):
    """Bytes this library never produced: the oracle's store -> dcdf_superchunk_open -> queries; two-level and nested."""
    from dcdf_b200 import Superchunk, synth
    data = synth.raster_slice(0, 10, 150, 200, nan_ocean=True).numpy()
    for levels in ([2, 6], [1, 1, 6]):
        ref = orc.superchunk_build(data, levels)
        nodes = ref.save().nodes()
        store = {c: b for c, _, b in nodes}
        sc = Superchunk.open(ctx, [nodes[-1][0]], store)
        assert sc.node_count() == (1 if len(levels) == 2 else 5)
        _query_checks(sc, data, [ref], 10)
        sc.close()
    ints = fx.array(40, 6, np.int32)
    ref = orc.superchunk_build(ints, [3, 3])
    nodes = ref.save().nodes()
    sc = Superchunk.open(ctx, [nodes[-1][0]], {c: b for c, _, b in nodes})
    assert np.array_equal(sc.window(0, 6, 0, 40, 0, 40), ints)
    sc.close()


def test_open_rejects_damaged_or_missing_objects(ctx):
    from dcdf_b200 import DcdfError, Superchunk, synth
    data = synth.raster_slice(0, 6, 100, 130).numpy()
    ref = orc.superchunk_build(data, [2, 6])
    nodes = ref.save().nodes()
    store = {c: b for c, _, b in nodes}
    root = nodes[-1][0]
    chunk_cid = next(c for c, t, _ in nodes if t == 4)
    links_cid = next(c for c, t, _ in nodes if t == 1)
    cases = []
    bad = dict(store); del bad[chunk_cid]; cases.append(("missing chunk", bad, 8))
    bad = dict(store); bad[root] = store[root][:-5]; cases.append(("truncated superchunk node", bad, 6))
    bad = dict(store); bad[root] = b"\x00\x00" + store[root][2:]; cases.append(("bad magic", bad, 6))
    bad = dict(store); bad[links_cid] = store[links_cid][:11]; cases.append(("Links cut short", bad, 6))
    b = bytearray(store[chunk_cid]); b[8 + 6 + 1 + 13 + 3] ^= 0x04; bad = dict(store); bad[chunk_cid] = bytes(b)
    cases.append(("chunk nodemap length changed", bad, 6))
    bad = dict(store); bad[chunk_cid] = store[chunk_cid][:-7]; cases.append(("chunk cut short", bad, 6))
    for what, st, code in cases:
        with pytest.raises(DcdfError) as e:
            Superchunk.open(ctx, [root], st)
        assert e.value.code == code, (what, e.value)
    Superchunk.open(ctx, [root], store).close()


def test_single_object_fetch_matches_the_bulk_transfer(ctx):
    """dcdf_saved_node_bytes (one object, host or device destination) against dcdf_saved_all_bytes."""
    import ctypes as C
    from dcdf_b200 import Superchunk, _ffi
    rng = np.random.default_rng(5)
    data = (rng.integers(0, 50, (24, 100, 130)) / 4).astype(np.float32)
    sc = Superchunk.build(ctx, data, [2, 6], compute_bits=True, chunk_size=8)
    nodes, _ = sc.save(1)
    lib = ctx._lib
    h = C.c_void_p()
    ctx.check(lib.dcdf_superchunk_save(ctx._h, sc._h, 1, C.byref(h)))
    try:
        for i in (0, len(nodes) // 2, len(nodes) - 1):
            buf = np.zeros(len(nodes[i][2]), np.uint8)
            ctx.check(lib.dcdf_saved_node_bytes(ctx._h, h, i, buf.ctypes.data_as(C.c_void_p), len(buf), 0))
            assert buf.tobytes() == nodes[i][2]
        small = np.zeros(3, np.uint8)
        assert lib.dcdf_saved_node_bytes(ctx._h, h, 0, small.ctypes.data_as(C.c_void_p), 3, 0) != 0   # destination too small
    finally:
        lib.dcdf_saved_free(h)
    sc.close()
