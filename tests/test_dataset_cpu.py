"""Dataset node and coordinates (dataset.rs:408-471, 676-820; range.rs; time.rs) and the flow of py-dcdf's own tests on
the host logic alone: the codec is replaced by a stand-in (tests/stub_superchunk.py), everything above it is the product's."""
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import span_oracle as so  # noqa: E402

import dataset_flow as flow  # noqa: E402
from stub_superchunk import StubSuperchunk  # noqa: E402

from dcdf_b200 import Coordinate, Dataset, DcdfError  # noqa: E402
from dcdf_b200 import span as sp  # noqa: E402


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64])
def test_new(dtype):
    """test_dcdf.py:69-104 / dataset.rs:1118-1150."""
    ds = flow.make_one(None, {}, dtype)
    assert [c.name for c in ds.coordinates] == ["t", "y", "x"]
    assert ds.shape == (16, 16) and ds.prev is None and ds.cid is None and len(ds.variables) == 0
    with pytest.raises(ValueError):
        len(ds.t)
    assert ds.t[10] == np.datetime64("1970-01-01T00:16:40") and ds.t.dtype == np.datetime64
    assert np.array_equal(ds.t[20:30], np.arange(np.datetime64("1970-01-01T00:33:20"), np.datetime64("1970-01-01T00:50:00"),
                                                 np.timedelta64(100, "s")))
    assert len(ds.get_coordinate("y")) == 16 and ds.y[10] == 40 and ds.y.dtype == dtype
    assert np.array_equal(ds.y[10:], np.arange(40, 160, 20)) and ds.y[10:].dtype == dtype
    assert len(ds.get_coordinate("x")) == 16 and ds.x[10] == 50 and ds.x.dtype == dtype
    assert np.array_equal(ds.x[:10], np.arange(-200, 50, 25))
    with pytest.raises(AttributeError):
        ds.doesnotexist
    with pytest.raises(DcdfError):
        ds.y[16]                                                        # range.rs:45-52 panics
    with pytest.raises(ValueError):
        ds.y[0:4:2]
    with pytest.raises(ValueError):
        Coordinate.range("foo", 0, 1, 10, np.byte)                     # test_dcdf.py:307-309
    assert ds.get_coordinate("nope") is None and ds.get_variable("nope") is None


def test_time_range_vectors_of_the_reference():
    t = Coordinate.time("t", 1000000, 3600)                            # time.rs:30-41
    assert t.get(0) == np.datetime64(1000000, "s") and t.get(100) == np.datetime64(1360000, "s")
    assert t.slice(100, 102).astype(np.int64).tolist() == [1360000, 1363600]


def test_stored_dataset_node_bytes():
    """Every field spelled out against dataset.rs:413-436, :676-768, :1018-1040 (big-endian, extio.rs:196-265)."""
    store = {}
    ds = flow.make_one(None, store, np.float32)
    want = bytes([0xDC, 0xE0, 0, 0, 0, 1, 0])
    want += b"\x01t\x00" + struct.pack(">qq", 0, 100)
    want += b"\x01y\x20" + struct.pack(">ffI", -160, 20, 16)
    want += b"\x01x\x20" + struct.pack(">ffI", -200, 25, 16)
    want += b"\x00" + struct.pack(">II", 16, 16) + b"\x00"
    cid = ds.commit()
    assert store[cid] == want and cid == sp.cid_of(want) and ds.cid is None
    ds2 = ds.add_variable("dates", 10, 20, [2, 2], True, 2)
    empty_span = bytes([0xDC, 0xE0, 0, 0, 0, 1, 2, 3, 32]) + struct.pack(">IIIII", 0, 16, 16, 20, 0)   # dataset.rs:127-129
    assert ds2.dates.cid == sp.cid_of(empty_span) and store[ds2.dates.cid] == empty_span
    assert len(ds.variables) == 0 and ds2.prev is None                  # `ds` was never loaded: no CID to link (dataset.rs:143-147)
    rec = b"\x05dates\x01\x02" + struct.pack(">IIB", 10, 20, 2) + struct.pack(">II", 2, 2) + b"\x20" + ds2.dates.cid
    want2 = want[:-10] + b"\x01" + rec + want[-9:]
    cid2 = ds2.commit()
    assert store[cid2] == want2
    back = Dataset.load(None, store, cid2)
    assert back.cid == cid2 and back.prev is None and back.shape == (16, 16)
    assert [(c.name, c.encoding, c.start, c.step, c.steps) for c in back.coordinates] == \
           [(c.name, c.encoding, c.start, c.step, c.steps) for c in ds2.coordinates]
    assert back.dates.write_to() == rec and back.dates.shape == [0, 16, 16]
    ds3 = back.add_variable("pears", 10, 20, (2, 2), dtype=np.float64)
    assert ds3.prev == cid2 and ds3.cid is None
    cid3 = ds3.commit()
    assert store[cid3].endswith(b"\x01" + cid2) and Dataset.load(None, store, cid3).prev == cid2
    assert Dataset.load(None, store, cid3).ls() == [("dates", ds2.dates.cid), ("pears", ds3.pears.cid), ("prev", cid2)]
    with pytest.raises(ValueError):
        ds3.add_variable("pears", 10, 20, (2, 2))
    for bad in (store[cid3][:-1], store[cid3] + b"\x00", b"\xDC\xE0\x00\x00\x00\x01\x02" + store[cid3][7:], store[cid3][:60]):
        store[b"bad"] = bad
        with pytest.raises(DcdfError):
            Dataset.load(None, store, b"bad")


def test_python_test_suite_of_the_reference_on_the_host_logic(monkeypatch):
    from dcdf_b200 import variable as var
    monkeypatch.setattr(var, "Superchunk", StubSuperchunk)
    store = {}
    ds, test_data, cid = flow.populate(None, store, rounds=False)
    flow.check_metadata(ds)
    flow.check_queries(ds, test_data)
    with pytest.raises(ValueError):
        ds.append("apples", np.arange(10, dtype=np.byte))               # test_dcdf.py:301-304
    with pytest.raises(KeyError):
        ds.append("kiwis", test_data["apples"])
    # the committed first version is untouched by the later appends and still loads (immutability, dataset.rs:268-325)
    first = Dataset.load(None, store, cid)
    assert [v.name for v in first.variables] == list(flow.VARIABLES[:4]) and first.prev is None
    # every variable's span tree is the one the reference's walk builds from the same chunk CIDs (10 chunks per span:
    # grapes -> 19 chunks, two levels; 489 instants -> 25 chunks, two levels)
    final = Dataset.load(None, store, ds.commit())
    flow.check_queries(final, test_data)
    for v in final.variables:
        ov = so.OVariable({}, [16, 16], 20, 10, sp.ENCODINGS[v.dtype.name])
        ov.append(list(zip(v.roots, v.instants)), False)
        assert ov.cid == v.cid, v.name
        assert sum(v.instants) == test_data[v.name].shape[0]


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64])
def test_range_vectors_of_the_reference(dtype):
    """range.rs:144-260: IntRange / FloatRange (-20, 5, 30) against (-20..130).step_by(5): get, every symmetric slice,
    and both out-of-bounds panics."""
    data = np.arange(-20, 130, 5).astype(dtype)
    rng = Coordinate.range("r", -20, 5, 30, dtype)
    assert len(rng) == 30 and rng.slice(0, 30).shape == (30,)
    for i in range(30):
        assert rng.get(i) == data[i] and type(rng.get(i)) is dtype
    for i in range(15):
        got = rng.slice(i, 30 - i)
        assert got.dtype == dtype and np.array_equal(got, data[i:30 - i])
    with pytest.raises(DcdfError):
        rng.get(30)
    with pytest.raises(DcdfError):
        rng.slice(29, 31)


def test_python_test_suite_of_the_reference_over_oracle_written_nodes(monkeypatch):
    """The same flow with the CPU oracle as the codec: every stored object -- subchunks, Links, superchunk nodes, spans,
    the dataset -- is reference-format bytes, rounding variables included, and the host logic reads all of it back."""
    from oracle_superchunk import OracleSuperchunk
    from dcdf_b200 import variable as var
    monkeypatch.setattr(var, "Superchunk", OracleSuperchunk)
    store = {}
    ds, test_data, cid = flow.populate(None, store, rounds=True)
    flow.check_metadata(ds)
    final = Dataset.load(None, store, ds.commit())
    flow.check_queries(final, test_data)
    kinds = {}
    for b in store.values():                                            # what the DAG is made of (node.rs:9-15)
        k = ("dataset",) if b[6] == 0 else ("links",) if b[6] == 1 else ("mmstruct3", b[7])
        kinds[k] = kinds.get(k, 0) + 1
    assert kinds[("dataset",)] == 2 and kinds[("links",)] >= 1
    roots = {r for v in final.variables for r in v.roots}              # the rasters repeat every 60 instants: equal slices
    assert kinds[("mmstruct3", 5)] >= len(roots) >= 12                 # are ONE object (content addressing de-duplicates)
    assert kinds[("mmstruct3", 3)] >= 6 and kinds[("mmstruct3", 4)] >= 1
    for v in final.variables:
        assert v.fractional_bits == (2 if v.round is not None else (0 if v.dtype.kind == "i" else 3)), v.name
