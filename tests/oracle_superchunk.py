"""A stand-in for dcdf_b200.api.Superchunk backed by the CPU ORACLE (tests only): build / save produce the real stored
nodes (superchunk.rs:199-270, 678-710), queries run the oracle's Superchunk::get.  With it the host logic above the C-ABI
-- Variable, the span tree, the Dataset node -- is exercised on the CPU over reference-format bytes; the product itself
never sees the oracle (tests/test_capi_cpu.py keeps it that way)."""
import types

import numpy as np

import oracle_lib as orc

_DTYPES = {4: np.int32, 8: np.int64, 32: np.float32, 64: np.float64}


class OracleSuperchunk:
    built = {}                                   # root CID -> oracle handle (this process's "device memory")

    def __init__(self, handles):
        self.handles = handles

    @classmethod
    def build(cls, ctx, data, k2_levels, fractional_bits=0, round=False, compute_bits=True, chunk_size=64):
        data = np.ascontiguousarray(data)
        return cls([orc.superchunk_build(np.ascontiguousarray(data[a:a + chunk_size]), list(k2_levels), fractional_bits=fractional_bits,
                                         round_=round, compute_bits=compute_bits) for a in range(0, data.shape[0], chunk_size)])

    @property
    def n_slices(self):
        return len(self.handles)

    def save(self, s):
        saved = self.handles[s].save()
        nodes = saved.nodes()
        OracleSuperchunk.built[nodes[-1][0]] = self.handles[s]
        return nodes, saved.stats()

    def info(self, s):
        ni = self.handles[s].node_info(0)
        return types.SimpleNamespace(shape=list(ni.shape), fractional_bits=ni.fractional_bits, encoding=ni.encoding)

    @classmethod
    def open(cls, ctx, cids, store):
        for c in cids:
            assert bytes(store[c])[:8] == bytes([0xDC, 0xE0, 0, 0, 0, 1, 2, 5])    # a stored superchunk node is there
        return cls([cls.built[bytes(c)] for c in cids])

    def window(self, a, b, top, bottom, left, right):
        first = self.handles[0].node_info(0)
        per, dtype = int(first.shape[0]), _DTYPES[first.encoding]
        out = np.empty((b - a, bottom - top, right - left), dtype)
        rr, cc = np.meshgrid(np.arange(top, bottom), np.arange(left, right), indexing="ij")
        for t in range(a, b):
            h = self.handles[t // per]
            irc = np.stack([np.full(rr.size, t % per), rr.ravel(), cc.ravel()], axis=1)
            fixed, bits = h.get_batch(irc)
            if np.dtype(dtype).kind == "i":
                vals = fixed.astype(dtype)
            else:
                vals = np.empty(fixed.shape, dtype)
                for fb in np.unique(bits):
                    m = bits == fb
                    vals[m] = orc.from_fixed_array(fixed[m], int(fb), dtype)
            out[t - a] = vals.reshape(rr.shape)
        return out

    def get(self, t, r, c):
        return self.window(t, t + 1, r, r + 1, c, c + 1)[0, 0, 0]

    def cell(self, a, b, r, c):
        return self.window(a, b, r, r + 1, c, c + 1)[:, 0, 0]

    def total_bytes(self):
        return sum(int(h.node_info(0).bytes0) for h in self.handles)

    def close(self):
        pass
