"""The flow of the reference's own Python tests (py-dcdf/tests/test_dcdf.py:46-300; the Rust twin is dataset.rs:1185-1457)
written once, run twice: on the CPU with a stand-in Superchunk (host logic only) and on the GPU with the real codec."""
import itertools

import numpy as np

from fixtures import fixed_array                                     # three 8x8 f32 rasters, multiples of 1/8 (fixed.rs:165-205)

VARIABLES = ("apples", "pears", "bananas", "grapes", "dates", "melons")


def make_data(instants):
    data = np.tile(fixed_array(3, np.float64), [instants // 3 + 1, 2, 2])[:instants]      # test_dcdf.py:46-49
    assert data.shape == (instants, 16, 16)
    return data


def make_one(ctx, store, dtype):
    from dcdf_b200 import Coordinate, Dataset
    t = Coordinate.time("t", 0, np.timedelta64(100, "s"))
    y = Coordinate.range("y", -160, 20, 16, dtype)
    x = Coordinate.range("x", -200, 25, 16, dtype)
    return Dataset.new(ctx, store, [t, y, x], [16, 16])


def populate(ctx, store, rounds=True):
    """test_dcdf.py:108-171.  rounds=False: the stand-in codec does not round, so `dates` / `melons` read back unrounded."""
    from dcdf_b200 import Dataset
    ds = make_one(ctx, store, np.float64)
    assert ds.cid is None
    cuts = {"apples": (360, np.float32, (99, 200)), "pears": (500, np.float64, (189, 400)),
            "bananas": (511, np.int32, (59, 300)), "grapes": (365, np.int64, (179, 300))}
    test_data = {}
    for name, (n, dt, (a, b)) in cuts.items():
        data = make_data(n).astype(dt)
        ds = ds.add_variable(name, 10, 20, (2, 2), dtype=dt)
        for lo, hi in ((0, a), (a, b), (b, n)):
            ds = ds.append(name, data[lo:hi])
        test_data[name] = data
    assert ds.prev is None and len(ds.ls()) == 4
    cid = ds.commit()
    ds = Dataset.load(ctx, store, cid)
    assert ds.cid == cid
    for name, dt in (("dates", np.float32), ("melons", np.float64)):
        data = make_data(489).astype(dt)
        ds = ds.add_variable(name, 10, 20, [2, 2], True, 2, dt)
        ds = ds.append(name, data)
        test_data[name] = ((data * 4 + 0.001).round() / 4).astype(dt) if rounds else data
        assert ds.cid is None and ds.prev == cid
    assert len(ds.ls()) == 7 and ds.ls()[6] == ("prev", cid)          # dataset.rs:1275-1276
    return ds, test_data, cid


def check_metadata(ds):
    """test_dcdf.py:184-246."""
    dtypes = dict(apples=np.float32, pears=np.float64, bananas=np.int32, grapes=np.int64, dates=np.float32, melons=np.float64)
    for name in VARIABLES:
        v = getattr(ds, name)
        assert (v.name, v.span_size, v.chunk_size, v.k2_levels, v.dtype) == (name, 10, 20, (2, 2), np.dtype(dtypes[name]))
        assert v.round == (2 if name in ("dates", "melons") else None)


def check_queries(ds, test_data):
    """test_dcdf.py:249-299: get / cell / window through __getitem__, then every slice permutation."""
    for var in VARIABLES:
        data, v = test_data[var], ds.get_variable(var)
        instants, rows, cols = v.shape
        assert (instants, rows, cols) == data.shape
        for instant in range(0, instants, 13):
            for row in range(0, rows, 4):
                for col in range(0, cols, 3):
                    assert v[instant, row, col].data == data[instant, row, col]
        for row in range(0, rows, 4):
            for col in range(0, cols, 3):
                start = row + col
                end = instants - start
                assert np.array_equal(v[start:end, row, col].data, data[start:end, row, col])
        for top in range(0, rows // 2, 4):
            bottom = top + rows // 2
            for left in range(0, cols // 2, 3):
                right = left + cols // 2
                start = top + bottom
                end = instants - start
                assert np.array_equal(v[start:end, top:bottom, left:right].data, data[start:end, top:bottom, left:right])
    data, v = test_data["apples"], ds.apples
    slice_args = [42, slice(23, 80), slice(None, 20)]
    slice_args += list(itertools.product([42, slice(23, 80)], [9, slice(6, None)]))
    slice_args += list(itertools.product([42, slice(23, 80)], [9, slice(6, 13)], [6, slice(3, 15)]))
    for arg in slice_args:
        expected, got = data.__getitem__(arg), v.__getitem__(arg).data
        if isinstance(expected, (int, float, np.number)):
            assert got == expected
        else:
            assert got.shape == expected.shape and np.array_equal(expected, got)
