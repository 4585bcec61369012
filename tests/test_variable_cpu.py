"""Host logic of dcdf_b200.variable that needs no GPU: the __getitem__ dispatch of py-dcdf/dcdf/__init__.py:281-336
(scalar -> get, (slice, int, int) -> cell, anything else -> window + squeeze of the scalar axes given explicitly)."""
import numpy as np
import pytest

from dcdf_b200.variable import MMArray3


class Fake(MMArray3):
    def __init__(self, a):
        self.a, self.calls = a, []

    @property
    def shape(self):
        return list(self.a.shape)

    def get(self, i, r, c):
        self.calls.append("get")
        return self.a[i, r, c]

    def cell(self, s, e, r, c):
        self.calls.append("cell")
        return self.a[s:e, r, c]

    def window(self, s, e, t, b, l, r):
        self.calls.append("window")
        return self.a[s:e, t:b, l:r]


def test_getitem_follows_the_reference_dispatch():
    a = np.arange(5 * 6 * 7).reshape(5, 6, 7)
    v = Fake(a)
    assert v[2, 3, 4].data == a[2, 3, 4] and v.calls[-1] == "get"               # test_dcdf.py:235-299 style
    assert np.array_equal(v[1:4, 3, 4].data, a[1:4, 3, 4]) and v.calls[-1] == "cell"
    assert np.array_equal(v[:, 3, 4].data, a[:, 3, 4]) and v.calls[-1] == "cell"
    for idx in [np.s_[1:4, 2:5, 3:6], np.s_[2, 2:5, 3:6], np.s_[1:4, 2, 3:6], np.s_[1:4, 2:5, 3], np.s_[2, 3, 1:5], np.s_[2], np.s_[1:3],
                np.s_[2, 3], np.s_[:, :, :], np.s_[:2, :, 4:]]:
        got = v[idx]
        assert np.array_equal(got.data, a[idx]), idx
        assert v.calls[-1] == "window"
        assert np.array_equal(got[0], a[idx][0])                                  # _Slice forwards indexing
    with pytest.raises(IndexError):
        v[1, 2, 3, 4]


from stub_superchunk import StubSuperchunk as _StubSuperchunk, fake_superchunk_node as _fake_superchunk_node  # noqa: E402


def test_variable_record_round_trip_through_the_span_tree():
    """write_to / load (dataset.rs:1018-1075) and the walk of the span tree, on stand-in superchunk nodes (no GPU)."""
    from dcdf_b200 import span as sp
    from dcdf_b200.api import DcdfError
    from dcdf_b200.variable import Variable
    store = {}
    v = Variable(None, store, [2, 2], chunk_size=20, round=2, span_size=3, name="dates")       # dataset.rs:1256
    with pytest.raises(ValueError):
        v.write_to()
    v.rows, v.cols = 16, 16
    v.tree = sp.SpanTree(store, 16, 16, 20, 3, sp.ENCODINGS["float32"])
    lengths = [20] * 10 + [7]
    for i, n in enumerate(lengths):
        b = _fake_superchunk_node(i, n, 16, 16, 2 + i % 3)
        cid = sp.cid_of(b)
        store[cid] = b
        v.tree.append(cid, n)
        v.roots.append(cid); v.instants.append(n); v.slice_bits.append(2 + i % 3); v.stats.append(None)
    rec = v.write_to()
    assert rec[:6] == b"\x05dates" and rec[6:8] == b"\x01\x02"
    assert rec[8:16] == (3).to_bytes(4, "big") + (20).to_bytes(4, "big") and rec[16] == 2
    assert rec[17:25] == (2).to_bytes(4, "big") * 2 and rec[25] == 32 and rec[26:] == v.cid and len(rec) == 26 + 36
    w = Variable.load(None, store, rec)
    assert (w.name, w.round, w.span_size, w.chunk_size, w.k2_levels, w.dtype) == ("dates", 2, 3, 20, (2, 2), np.float32)
    assert (w.roots, w.instants, w.slice_bits) == (v.roots, v.instants, v.slice_bits)
    assert w.shape == [207, 16, 16] and w.cid == v.cid and w.tree.tail() == (v.roots[-1], 7)
    u = Variable(None, store, [2, 2], chunk_size=20, round=None, span_size=3, name="apples")
    u.rows, u.cols, u.tree = 16, 16, v.tree
    assert Variable.load(None, store, u.write_to()).round is None
    for bad in (rec[:-1], rec + b"\x00", rec[:25] + b"\x07" + rec[26:]):
        with pytest.raises(DcdfError):
            Variable.load(None, store, bad)


def test_append_drives_the_span_tree_like_the_reference(monkeypatch):
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import span_oracle as so
    from dcdf_b200 import variable as var
    monkeypatch.setattr(var, "Superchunk", _StubSuperchunk)
    rng = np.random.default_rng(5)
    data = rng.standard_normal((61, 6, 7)).astype(np.float32)
    store = {}
    v = var.Variable(None, store, [1, 2], chunk_size=4, span_size=2)
    ov = so.OVariable({}, [6, 7], 4, 2, 32)
    cuts = [0, 3, 4, 13, 14, 30, 31, 61]
    for a, b in zip(cuts, cuts[1:]):
        had_tail = ov.tail_data() is not None
        n_keep = len(v.roots) - (1 if had_tail else 0)
        v.append(data[a:b])
        ov.append(list(zip(v.roots[n_keep:], v.instants[n_keep:])), had_tail)    # the reference's walk on the same chunk CIDs
        assert v.cid == ov.cid, (a, b)
        assert v.shape == [b, 6, 7] and np.array_equal(v.window(0, b, 0, 6, 0, 7), data[:b])
    assert v.instants == [4] * 15 + [1]
    one = var.Variable(None, {}, [1, 2], chunk_size=4, span_size=2)
    one.append(data)
    assert one.cid == v.cid                                      # the tree does not depend on how the instants arrived
    w = var.Variable.load(None, store, v.write_to())
    assert w.roots == v.roots and np.array_equal(w.window(7, 50, 2, 3, 3, 4)[:, 0, 0], data[7:50, 2, 3])
    w.append(data[:5])
    assert w.instants == [4] * 16 + [2] and w.shape[0] == 66


def test_a_failed_append_leaves_the_variable_as_it_was(monkeypatch):
    from dcdf_b200 import variable as var
    from dcdf_b200.api import DcdfError

    class Failing(_StubSuperchunk):
        @classmethod
        def build(cls, ctx, data, *a, **k):
            if np.isnan(np.asarray(data)).any():
                raise DcdfError(1, "non-finite value")             # fixed.rs:40
            return super().build(ctx, data, *a, **k)

    monkeypatch.setattr(var, "Superchunk", Failing)
    data = np.arange(10 * 4 * 4, dtype=np.float32).reshape(10, 4, 4)
    v = var.Variable(None, {}, [1, 1], chunk_size=4, span_size=2)
    v.append(data[:6])
    before = (list(v.roots), list(v.instants), v.cid, v.tree.tail())
    bad = data[6:].copy()
    bad[1, 2, 3] = np.nan
    with pytest.raises(DcdfError):
        v.append(bad)
    assert (v.roots, v.instants, v.cid, v.tree.tail()) == before
    assert np.array_equal(v.window(0, 6, 0, 4, 0, 4), data[:6])
    assert v.append(data[6:6]).shape == [6, 4, 4] and v.cid == before[2]   # nothing to append: nothing changes
    v.append(data[6:])
    assert v.instants == [4, 4, 2] and np.array_equal(v.window(0, 10, 0, 4, 0, 4), data)


def test_cache_is_lru_by_bytes_and_loads_once_under_concurrent_requests(monkeypatch):
    """cache.rs:37-232 semantics on the device-resident cache, with a stand-in for the opened handle."""
    import threading
    import time
    from dcdf_b200 import variable as var
    opened, closed = [], []

    class Handle:
        def __init__(self, key):
            self.key = key

        def total_bytes(self):
            return 100

        def close(self):
            closed.append(self.key)

    class Opener:
        @staticmethod
        def open(ctx, cids, store):
            time.sleep(0.02)                                        # a load in flight while other threads ask for the key
            opened.append(tuple(cids))
            return Handle(tuple(cids))

    monkeypatch.setattr(var, "Superchunk", Opener)
    cache = var.ChunkCache(None, {}, cache_bytes=250)              # room for two handles
    got = []
    threads = [threading.Thread(target=lambda: got.append(cache.get([b"a"]))) for _ in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(opened) == 1 and len({id(h) for h in got}) == 1 and (cache.hits, cache.misses) == (7, 1)
    cache.get([b"b"])
    cache.get([b"a"])                                               # a is now the most recently used
    cache.get([b"c"])                                               # 300 bytes > 250: the least recently used (b) goes
    assert closed == [(b"b",)] and cache.evictions == 1 and cache.resident_bytes == 200
    cache.get([b"b"])
    assert closed == [(b"b",), (b"a",)] and opened.count((b"b",)) == 2
    cache.invalidate([b"c"])
    assert closed[-1] == (b"c",) and cache.resident_bytes == 100
    cache.close()
    assert cache.resident_bytes == 0 and sorted(closed)[-1] == (b"c",)
