"""examples/example.py (the reference's example script on this package) and DirStore, host logic only: init / copy /
query with a HEAD file over a directory store; the codec is the stand-in of tests/stub_superchunk.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

from stub_superchunk import StubSuperchunk  # noqa: E402


def test_dir_store_is_content_addressed_and_persistent(tmp_path):
    from dcdf_b200 import DirStore
    from dcdf_b200 import span as sp
    store = DirStore(tmp_path / "objects", verify=True)
    a, b = b"first object", b"second object " * 1000
    ca, cb = sp.cid_of(a), sp.cid_of(b)
    store[ca] = a
    store[cb] = b
    store[ca] = a                                                   # writing an existing object is a no-op
    assert store[ca] == a and store.get(cb) == b and ca in store and len(store) == 2 and set(store) == {ca, cb}
    missing = sp.cid_of(b"missing")
    assert missing not in store and store.get(missing) is None and b"short" not in store
    with pytest.raises(KeyError):
        store[missing]
    with pytest.raises(ValueError):
        store[missing] = b"something else"                          # verify: the bytes must hash to the CID
    again = DirStore(tmp_path / "objects")
    assert again[cb] == b and len(again) == 2
    assert not [f for _, _, fs in os.walk(tmp_path) for f in fs if len(f) != 72]   # no temp files left behind


def test_example_life_cycle_on_the_host_logic(tmp_path, monkeypatch, capsys):
    import example
    from dcdf_b200 import DirStore, variable as var
    monkeypatch.setattr(var, "Superchunk", StubSuperchunk)

    class SmallCpc(example.CpcPrecip):                              # same layout, a 45x90 grid so the stand-in stays cheap
        shape = (45, 90)

        @staticmethod
        def factory(ctx, store):
            monkeypatch.setattr(example.CpcPrecip, "shape", (45, 90))
            return example.CpcPrecip.factory(ctx, store)

        @staticmethod
        def source(start, stop, device="cpu"):
            from dcdf_b200 import synth
            return synth.raster_slice(start, stop, 45, 90, hourly=False, base=0, device=device).clamp_(min=0)

    root = str(tmp_path)
    store = DirStore(os.path.join(root, "objects"))
    ds0 = example.initialize_dataset(SmallCpc, None, store, root)
    assert ds0.precip.shape == [0, 45, 90] and (ds0.precip.span_size, ds0.precip.chunk_size, ds0.precip.k2_levels) == (20000, 64, (4, 6))
    with pytest.raises(SystemExit):
        example.initialize_dataset(SmallCpc, None, store, root)
    ds1 = example.copy_data(SmallCpc, None, store, root, 150, commit_every=2)
    assert ds1.precip.shape == [150, 45, 90] and ds1.precip.instants == [64, 64, 22]
    ds2 = example.copy_data(SmallCpc, None, store, root, 50, commit_every=10)     # the 22-instant tail is re-encoded
    assert ds2.precip.shape == [200, 45, 90] and ds2.precip.instants == [64, 64, 64, 8] and ds2.prev is not None
    fresh = DirStore(os.path.join(root, "objects"))                 # a new process would see the same objects
    ds3 = example.query(SmallCpc, None, fresh, root, with_search=False)
    assert ds3.precip.roots == ds2.precip.roots and ds3.cid is not None
    out = capsys.readouterr().out
    assert "Copied 200/200" in out and "Incremental progress saved." in out and "cell (15, 45): 200 instants from 1979-01-01" in out
    src = np.asarray(SmallCpc.source(0, 200))
    assert np.array_equal(ds3.precip[:, 3, 4].data, src[:, 3, 4])
