"""A stand-in for dcdf_b200.api.Superchunk so that the HOST logic above the C-ABI (Variable.append's tail re-encode, the
span tree, the Dataset node, cache groups) runs in the CPU suite: a 'stored superchunk node' here is the real 35-byte
header (superchunk.rs:683-692) followed by the raw raster.  Nothing of the codec is emulated."""
import struct
import types

import numpy as np

from dcdf_b200 import span as sp

_DTYPES = {v: np.dtype(k) for k, v in sp.ENCODINGS.items()}


def fake_superchunk_node(i, instants, rows, cols, bits, enc=32):
    return bytes([0xDC, 0xE0, 0, 0, 0, 1, 2, 5]) + struct.pack(">IIIIBIIBB", instants, rows, cols, 64, 2, 16, 4, bits, enc) + b"%d" % i


class StubSuperchunk:
    def __init__(self, slices):
        self.slices = slices

    @classmethod
    def build(cls, ctx, data, k2_levels, fractional_bits=0, round=False, compute_bits=True, chunk_size=64):
        data = np.ascontiguousarray(data)
        return cls([data[a:a + chunk_size] for a in range(0, data.shape[0], chunk_size)])

    @property
    def n_slices(self):
        return len(self.slices)

    def save(self, s):
        a = self.slices[s]
        node = fake_superchunk_node(0, a.shape[0], a.shape[1], a.shape[2], 3, sp.ENCODINGS[a.dtype.name])[:35] + a.tobytes()
        return [(sp.cid_of(node), 5, node)], None

    def info(self, s):
        return types.SimpleNamespace(shape=list(self.slices[s].shape), fractional_bits=3)

    @classmethod
    def open(cls, ctx, cids, store):
        out = []
        for c in cids:
            n, r, cc = (int.from_bytes(store[c][8 + 4 * i:12 + 4 * i], "big") for i in range(3))
            out.append(np.frombuffer(store[c][35:], _DTYPES[store[c][34]]).reshape(n, r, cc))
        return cls(out)

    def get(self, t, r, c):
        return np.concatenate(self.slices)[t, r, c]

    def cell(self, a, b, r, c):
        return np.concatenate(self.slices)[a:b, r, c]

    def window(self, a, b, t, bm, l, r):
        return np.concatenate(self.slices)[a:b, t:bm, l:r]

    def total_bytes(self):
        return sum(s.nbytes for s in self.slices)

    def close(self):
        pass
