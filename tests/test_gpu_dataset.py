"""The reference's own Python test-suite flow (py-dcdf/tests/test_dcdf.py:106-299; Rust twin dataset.rs:1185-1457) on the
real codec: six variables of four dtypes appended in pieces (tail re-encodes), committed, re-loaded, extended; every
get / cell / window / slice permutation read back through the span trees and the device-resident cache."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import span_oracle as so  # noqa: E402

import dataset_flow as flow  # noqa: E402
import oracle_lib as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def test_populate_commit_load_and_query_like_the_reference():
    from dcdf_b200 import Context, Dataset
    from dcdf_b200 import span as sp
    ctx = Context(0)
    store = {}
    ds, test_data, cid = flow.populate(ctx, store, rounds=True)
    flow.check_metadata(ds)
    flow.check_queries(ds, test_data)
    final = Dataset.load(ctx, store, ds.commit())                      # dataset.rs:1338-1339
    assert final.prev == cid and [v.name for v in final.variables] == list(flow.VARIABLES)
    flow.check_metadata(final)
    flow.check_queries(final, test_data)
    for v in final.variables:
        # the slices are the oracle's superchunk nodes, the span tree the reference's walk over their CIDs
        data = test_data[v.name]
        if v.round is None:
            for s in (0, len(v.roots) - 1):
                rn = orc.superchunk_build(np.ascontiguousarray(data[s * 20:(s + 1) * 20]), [2, 2]).save().nodes()
                assert v.roots[s] == rn[-1][0] and store[v.roots[s]] == rn[-1][2], (v.name, s)
        ov = so.OVariable({}, [16, 16], 20, 10, sp.ENCODINGS[v.dtype.name])
        ov.append(list(zip(v.roots, v.instants)), False)
        assert ov.cid == v.cid, v.name
    for v in ds.variables + final.variables:
        v.close()
    ctx.close()
