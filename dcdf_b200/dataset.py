"""The Dataset node above the variables (host metadata; py-dcdf/dcdf/__init__.py:52-232 is the surface mirrored here).

  stored Dataset node    dataset.rs:408-471 behind Resolver::save (resolver.rs:126-138, NODE_DATASET): three coordinates,
                         the variables' records (dataset.rs:1012-1040), raster shape, optional `prev` CID
  coordinates            dataset.rs:676-820 (name + kind), range.rs:17-115, time.rs:9-21
  add_variable / append / commit / prev chain      dataset.rs:108-140, 268-325

A Dataset here is immutable like the reference's: `add_variable` and `append` return a new one, `commit` stores it and a
later `append` links the committed version as `prev`.  The raster arithmetic is all behind `Variable.append`
(one `dcdf_superchunk_build` per call); this module only frames metadata, so that a DAG written here -- dataset ->
spans -> superchunks -> links -> subchunks -- is the reference's, node for node.
"""
import struct

import numpy as np

from . import span as _span
from .api import DcdfError
from .variable import Variable

NODE_DATASET = 0                                                       # node.rs:9
ENC_TIME = 0                                                           # mmstruct.rs:38
_HEAD = _span.MAGIC_AND_VERSION + bytes([NODE_DATASET])
_KINDS = {                                                             # encoding -> (start/step format, numpy dtype)
    ENC_TIME: (">qq", np.int64), 4: (">ii", np.int32), 8: (">qq", np.int64), 32: (">ff", np.float32), 64: (">dd", np.float64)}


class Coordinate:
    """A named axis: an endless time range (seconds since the epoch) or `steps` values start + i * step."""

    def __init__(self, name, encoding, start, step, steps=None):
        if encoding not in _KINDS:
            raise ValueError(f"unsupported encoding for Coordinate {encoding}")
        dt = _KINDS[encoding][1]
        self.name, self.encoding, self.start, self.step = str(name), int(encoding), dt(start), dt(step)
        self.steps = None if encoding == ENC_TIME else int(steps)

    @classmethod
    def time(cls, name, start, step):
        if isinstance(start, np.datetime64):
            start = int((start - np.datetime64(0, "s")) / np.timedelta64(1, "s"))
        if isinstance(step, np.timedelta64):
            step = int(step / np.timedelta64(1, "s"))
        return cls(name, ENC_TIME, start, step)

    @classmethod
    def range(cls, name, start, step, steps, dtype=np.float64):
        enc = _span.ENCODINGS.get(np.dtype(dtype).name)
        if enc is None:
            raise ValueError(f"unsupported dtype for Coordinate {dtype}")
        return cls(name, enc, start, step, steps)

    @property
    def dtype(self):
        return np.datetime64 if self.encoding == ENC_TIME else _KINDS[self.encoding][1]

    def __len__(self):
        if self.steps is None:
            raise ValueError("time is infinite")                       # Error::TimeIsInfinite, dataset.rs:642 (a ValueError in py-dcdf)
        return self.steps

    def _check(self, index):
        if self.steps is not None and not 0 <= index < self.steps:     # range.rs:45-52
            raise DcdfError(5, f"Out of bounds: index {index} is out of bounds for array with length {self.steps}")

    def get(self, index):
        self._check(index)
        dt = _KINDS[self.encoding][1]
        v = dt(index) * self.step + self.start
        return np.datetime64(int(v), "s") if self.encoding == ENC_TIME else v

    def slice(self, start, end):
        self._check(end - 1)
        dt = _KINDS[self.encoding][1]
        if self.encoding in (32, 64):                                   # Array1::range(first, last, step), range.rs:29-35
            first, last = dt(start) * self.step + self.start, dt(end) * self.step + self.start
            n = max(int(np.ceil((last - first) / self.step)), 0)
            return first + self.step * np.arange(n, dtype=dt)
        vals = self.start + self.step * np.arange(start, end, dtype=dt)
        return vals.astype("datetime64[s]") if self.encoding == ENC_TIME else vals

    def __getitem__(self, i):
        if isinstance(i, slice):
            if i.step is not None:
                raise ValueError("step not supported for slice")
            return self.slice(0 if i.start is None else i.start, len(self) if i.stop is None else i.stop)
        return self.get(i)

    # ---- dataset.rs:676-768
    def write_to(self):
        name = self.name.encode()
        if len(name) > 255:
            raise ValueError("name longer than 255 bytes")
        b = bytes([len(name)]) + name + bytes([self.encoding]) + struct.pack(_KINDS[self.encoding][0], self.start, self.step)
        return b if self.steps is None else b + struct.pack(">I", self.steps)

    @classmethod
    def read_from(cls, buf, pos):
        n = buf[pos]
        name, pos = bytes(buf[pos + 1:pos + 1 + n]).decode(), pos + 1 + n
        enc = buf[pos]
        if enc not in _KINDS:
            raise DcdfError(6, f"unknown coordinate encoding {enc}")
        fmt = _KINDS[enc][0]
        start, step = struct.unpack_from(fmt, buf, pos + 1)
        pos += 1 + struct.calcsize(fmt)
        steps = None
        if enc != ENC_TIME:
            (steps,) = struct.unpack_from(">I", buf, pos)
            pos += 4
        return cls(name, enc, start, step, steps), pos


class Dataset:
    """Coordinates + variables of one (time, y, x) grid; `store` maps CID -> stored bytes, `ctx` is the dcdf_b200 Context
    the variables encode and decode on."""

    def __init__(self, ctx, store, coordinates, shape, variables=(), prev=None, cid=None, cache_bytes=1 << 30):
        if len(coordinates) != 3 or len(shape) != 2:
            raise ValueError("a dataset has three coordinates (t, y, x) and a [rows, cols] shape")
        self.ctx, self.store, self.cache_bytes = ctx, store, cache_bytes
        self.coordinates, self.shape = list(coordinates), (int(shape[0]), int(shape[1]))
        self.variables, self.prev, self.cid = list(variables), prev, cid

    @classmethod
    def new(cls, ctx, store, coordinates, shape, cache_bytes=1 << 30):
        return cls(ctx, store, coordinates, shape, cache_bytes=cache_bytes)

    def _with(self, variables):
        prev = self.cid if self.cid is not None else self.prev         # dataset.rs:308-312
        return Dataset(self.ctx, self.store, self.coordinates, self.shape, variables, prev, None, self.cache_bytes)

    def add_variable(self, name, span_size, chunk_size, k2_levels, round=False, fractional_bits=0, dtype=np.float32):
        """dataset.rs:116-158: the variable starts as one empty span of this dataset's raster shape."""
        if any(v.name == name for v in self.variables):
            raise ValueError(f"variable {name!r} exists")
        v = Variable(self.ctx, self.store, k2_levels, chunk_size=chunk_size, round=fractional_bits if round else None,
                     span_size=span_size, dtype=dtype, cache_bytes=self.cache_bytes, name=name)
        v.rows, v.cols = self.shape
        v.tree = _span.SpanTree(self.store, v.rows, v.cols, v.chunk_size, v.span_size, _span.ENCODINGS[v.dtype.name])
        return self._with(self.variables + [v])

    def append(self, name, data):
        """Dataset::append_* (dataset.rs:268-325): the named variable grows, every other one is shared."""
        old = self.get_variable(name)
        if old is None:
            raise KeyError(name)                                        # Error::BadName
        if not hasattr(data, "is_cuda") and np.asarray(data).dtype != old.dtype:
            raise ValueError(f"Unsupported dtype: {np.asarray(data).dtype} for a {old.dtype} variable")
        new = old.fork()
        new.append(data)
        return self._with([new if v is old else v for v in self.variables])

    def commit(self):
        """Resolver::save(dataset) (dataset.rs:104-106): stores the node and returns its CID.  As in the reference the
        object itself keeps `cid` None; only a dataset loaded from the store knows its CID (resolver.rs:147-151)."""
        b = _HEAD + b"".join(c.write_to() for c in self.coordinates) + bytes([len(self.variables)])
        b += b"".join(v.write_to() for v in self.variables)
        b += struct.pack(">II", *self.shape)
        b += b"\x01" + self.prev if self.prev is not None else b"\x00"
        cid = _span.cid_of(b)
        self.store[cid] = b
        return cid

    @classmethod
    def load(cls, ctx, store, cid, cache_bytes=1 << 30):
        """Resolver::get_dataset (resolver.rs:79-87) -> Dataset::load_from (dataset.rs:439-471)."""
        buf = bytes(store[cid])
        if buf[:7] != _HEAD:
            raise DcdfError(6, "not a stored Dataset node")
        try:
            pos, coords = 7, []
            for _ in range(3):
                c, pos = Coordinate.read_from(buf, pos)
                coords.append(c)
            n_vars, pos = buf[pos], pos + 1
            records = []
            for _ in range(n_vars):                                     # a Variable record has no length prefix: walk it
                q = pos + 1 + buf[pos]
                q += 2 if buf[q] == 1 else 1
                q += 8
                q += 1 + 4 * buf[q]
                q += 1 + _span.CID_BYTES
                records.append(buf[pos:q])
                pos = q
            rows, cols = struct.unpack_from(">II", buf, pos)
            pos += 8
            prev = None
            if buf[pos] == 1:
                prev = buf[pos + 1:pos + 1 + _span.CID_BYTES]
                if len(prev) != _span.CID_BYTES:
                    raise IndexError
                pos += _span.CID_BYTES
            if pos + 1 != len(buf):
                raise IndexError
        except (IndexError, struct.error):
            raise DcdfError(6, "truncated or oversized Dataset node") from None
        variables = [Variable.load(ctx, store, r, cache_bytes) for r in records]
        for v in variables:
            if (v.rows, v.cols) != (rows, cols):
                raise DcdfError(6, "a variable does not match the dataset's raster shape")
        return cls(ctx, store, coords, [rows, cols], variables, prev, bytes(cid), cache_bytes)

    def ls(self):
        """Node::ls (dataset.rs:473-483)."""
        out = [(v.name, v.cid) for v in self.variables]
        return out + [("prev", self.prev)] if self.prev is not None else out

    def get_coordinate(self, name):
        return next((c for c in self.coordinates if c.name == name), None)

    def get_variable(self, name):
        return next((v for v in self.variables if v.name == name), None)

    def __getattr__(self, name):                                        # py-dcdf/dcdf/__init__.py:139-148
        for item in self.__dict__.get("coordinates", []) + self.__dict__.get("variables", []):
            if item.name == name:
                return item
        raise AttributeError(name)
