"""Variable.append (tail re-encode), the device-resident superchunk cache and the MMArray3 / __getitem__ surface, driven
the way the product drives the boundary (dataset.rs:268-325, :834-878; cache.rs:37-232; py-dcdf/dcdf/__init__.py:281-336)."""
import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from dcdf_b200 import Context
    c = Context(0)
    yield c
    c.close()


def test_append_with_tail_reencode_equals_one_build(ctx):
    """Appending 37 + 5 + 80 instants with chunk_size 16 re-encodes the incomplete last slice each time
    (dataset.rs:283-297) and must end with exactly the objects of one 122-instant append; queries see all of it."""
    from dcdf_b200 import Variable, synth
    data = synth.raster_slice(0, 122, 150, 200).numpy()
    store_a, store_b = {}, {}
    va = Variable(ctx, store_a, [2, 6], chunk_size=16, span_size=3)
    for a, b in ((0, 37), (37, 42), (42, 122)):
        va.append(data[a:b])
    vb = Variable(ctx, store_b, [2, 6], chunk_size=16, span_size=3)
    vb.append(data)
    assert va.roots == vb.roots and va.instants == vb.instants == [16] * 7 + [10]
    assert set(store_b) <= set(store_a)                       # the re-encoded tails left superseded objects behind, like the reference
    for s, root in enumerate(vb.roots):                      # every slice equals the oracle's superchunk node for that slice
        rn = orc.superchunk_build(data[s * 16:(s + 1) * 16], [2, 6]).save().nodes()
        assert root == rn[-1][0] and store_b[root] == rn[-1][2]
    assert va.shape == [122, 150, 200]
    assert np.array_equal(va.window(0, 122, 0, 150, 0, 200), data)
    assert np.array_equal(va[30:100, 149, 199].data, data[30:100, 149, 199])
    assert va[121, 3, 5].data == data[121, 3, 5]
    assert np.array_equal(va[5, 10:140, :].data, data[5, 10:140, :])
    assert np.array_equal(va[40:60].data, data[40:60])
    from dcdf_b200 import DcdfError
    with pytest.raises(DcdfError) as e:
        va.window(0, 123, 0, 150, 0, 200)
    assert e.value.code == 5
    # the span tree (dataset.rs:880-987): three appends with two tail updates end at the root one append builds, and at
    # the root the step-by-step restatement of the reference reaches from the same chunk CIDs
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import span_oracle as so
    assert va.cid == vb.cid and store_b[vb.cid] == store_a[va.cid]
    ov = so.OVariable({}, [150, 200], 16, 3, 32)
    ov.append(list(zip(vb.roots, vb.instants)), False)
    assert ov.cid == vb.cid
    # re-open from the stored record in a fresh object (Variable::load_from, dataset.rs:1042-1075) and carry on appending
    vc = Variable.load(ctx, store_a, va.write_to())
    assert (vc.roots, vc.instants, vc.slice_bits, vc.shape) == (va.roots, va.instants, va.slice_bits, va.shape)
    assert np.array_equal(vc.window(100, 122, 0, 150, 0, 200), data[100:122])
    more = synth.raster_slice(122, 142, 150, 200).numpy()
    vc.append(more)
    vb.append(more)
    assert vc.cid == vb.cid and vc.roots == vb.roots and vc.instants == [16] * 8 + [14]
    assert np.array_equal(vc.window(110, 142, 0, 150, 0, 200), np.concatenate([data, more])[110:142])
    va.close(); vb.close(); vc.close()


def test_device_cache_is_lru_by_bytes_and_reloads_from_the_store(ctx):
    from dcdf_b200 import Variable, synth
    data = synth.raster_slice(0, 96, 130, 140).numpy()
    store = {}
    v = Variable(ctx, store, [2, 6], chunk_size=8, span_size=2)      # 12 slices -> 6 groups
    v.append(data)
    one = v.cache.get(v.roots[0:2]).total_bytes()
    v.cache.invalidate()
    v.cache.cache_bytes = int(one * 2.5)                               # room for two groups
    v.cache.hits = v.cache.misses = v.cache.evictions = 0
    assert np.array_equal(v.window(0, 16, 0, 130, 0, 140), data[0:16])          # group 0: miss
    assert np.array_equal(v.window(3, 9, 5, 100, 5, 100), data[3:9, 5:100, 5:100])   # hit
    assert (v.cache.hits, v.cache.misses) == (1, 1)
    assert np.array_equal(v.window(10, 40, 0, 130, 0, 140), data[10:40])        # groups 0 (hit), 1, 2 (misses) -> group 0 evicted
    assert v.cache.misses == 3 and v.cache.evictions >= 1 and v.cache.resident_bytes <= v.cache.cache_bytes
    assert np.array_equal(v.cell(0, 96, 64, 64), data[:, 64, 64])               # everything, through reloads
    assert v.cache.resident_bytes <= v.cache.cache_bytes
    v.close()


def test_search_with_float_bounds(ctx):
    """mmarray.rs:407-417 is todo!() upstream: bounds in the variable's units, exact against numpy, NaN never matches;
    a subchunk with fewer fractional bits than its slice takes the decoded-window route."""
    from dcdf_b200 import Variable, synth
    data = synth.raster_slice(0, 40, 150, 200, nan_ocean=True).numpy()
    data[:, :64, 64:128] = np.round(data[:, :64, 64:128])                          # this subchunk: 0 fractional bits
    store = {}
    v = Variable(ctx, store, [2, 6], chunk_size=16, span_size=2)
    v.append(data)
    for lo, hi in ((1.0, 2.5), (-1.0, 0.0), (0.3, 0.31), (5.0, 3.0), (-3.0, 100.0)):
        got = v.search(3, 37, 10, 140, 20, 190, lo, hi)
        a, b = min(lo, hi), max(lo, hi)
        w = data[3:37, 10:140, 20:190]
        want = np.argwhere((w >= a) & (w <= b)) + np.array([3, 10, 20])
        assert {tuple(x) for x in got.tolist()} == {tuple(x) for x in want.tolist()}, (lo, hi)
        assert len(got) == len(want)
    ints = (synth.raster_slice(0, 20, 100, 100).numpy() * 16).astype(np.int32)
    vi = Variable(ctx, {}, [1, 6], chunk_size=8, dtype=np.int32)
    vi.append(ints)
    got = vi.search(0, 20, 0, 100, 0, 100, 4400.5, 4500)
    want = np.argwhere((ints >= 4401) & (ints <= 4500))
    assert {tuple(x) for x in got.tolist()} == {tuple(x) for x in want.tolist()}
    v.close(); vi.close()
