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
