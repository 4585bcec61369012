"""ctypes loader for the CPU oracle (oracle/liboracle.so).

Test infrastructure only: imported by tests/, bench.py's cpu_baseline leg and
__graft_entry__.smoke(); never by the product package dcdf_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

ENC = {np.dtype("int32"): 4, np.dtype("int64"): 8, np.dtype("float32"): 32, np.dtype("float64"): 64}
STATUS = {0: "OK", 1: "NONFINITE", 2: "PRECISION_LOSS", 3: "OVERFLOW", 4: "BAD_LEVELS", 5: "OUT_OF_BOUNDS",
          6: "BAD_FORMAT", 7: "CUDA", 8: "BAD_ARG"}


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_capi.cpp", "dcdf_oracle.hpp")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True, timeout=600)
    return LIB_PATH


class Stats(C.Structure):
    _fields_ = [("size", C.c_uint64), ("elided", C.c_uint32), ("local", C.c_uint32), ("external", C.c_uint32),
                ("snapshots", C.c_uint32), ("logs", C.c_uint32)]


class NodeInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("encoding", C.c_int32), ("fractional_bits", C.c_int32), ("levels", C.c_uint32),
                ("shape", C.c_int64 * 3), ("sidelen", C.c_int64), ("chunks_sidelen", C.c_int64),
                ("subsidelen", C.c_int64), ("n_refs", C.c_uint32), ("bytes0", C.c_uint64), ("bytes1", C.c_uint64),
                ("bytes2", C.c_uint64), ("stats", Stats)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(LIB_PATH)
        _lib.dcdf_oracle_last_error.restype = C.c_char_p
        _lib.dcdf_oracle_from_fixed_f32.restype = C.c_float
        _lib.dcdf_oracle_from_fixed_f32.argtypes = [C.c_int64, C.c_int32]
        _lib.dcdf_oracle_from_fixed_f64.restype = C.c_double
        _lib.dcdf_oracle_from_fixed_f64.argtypes = [C.c_int64, C.c_int32]
        _lib.dcdf_oracle_to_fixed_f32.argtypes = [C.c_float, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        _lib.dcdf_oracle_to_fixed_f64.argtypes = [C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        _lib.dcdf_oracle_dac_get_empty.restype = C.c_int64
        _lib.dcdf_oracle_free.argtypes = [C.c_void_p]
    return _lib


def _check(code):
    if code != 0:
        raise OracleError(code, lib().dcdf_oracle_last_error().decode())


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _i64x3(v):
    return (C.c_int64 * 3)(*[int(x) for x in v])


def _arr_desc(a):
    a = np.asarray(a)
    assert a.ndim == 3 and a.dtype in ENC, (a.ndim, a.dtype)
    strides = [s // a.itemsize for s in a.strides]
    return a, ENC[a.dtype], _i64x3(a.shape), _i64x3(strides)


# ------------------------------------------------------------------ fixed
def to_fixed(n, bits, round_=False, dtype=np.float64):
    out = C.c_int64()
    if np.dtype(dtype) == np.float32:
        _check(lib().dcdf_oracle_to_fixed_f32(C.c_float(n), bits, int(round_), C.byref(out)))
    else:
        _check(lib().dcdf_oracle_to_fixed_f64(C.c_double(n), bits, int(round_), C.byref(out)))
    return out.value


def from_fixed(n, bits, dtype=np.float32):
    if np.dtype(dtype) == np.float32:
        return lib().dcdf_oracle_from_fixed_f32(int(n), bits)
    return lib().dcdf_oracle_from_fixed_f64(int(n), bits)


def from_fixed_array(fixed, bits, dtype=np.float32):
    fixed = np.ascontiguousarray(fixed, dtype=np.int64)
    out = np.empty(fixed.shape, dtype=dtype)
    fn = lib().dcdf_oracle_from_fixed_array_f32 if np.dtype(dtype) == np.float32 else lib().dcdf_oracle_from_fixed_array_f64
    fn(_p(fixed), C.c_uint64(fixed.size), C.c_int32(int(bits)), _p(out))
    return out


def suggest_fraction(a):
    a, enc, shape, strides = _arr_desc(a)
    kind, bits = C.c_int32(), C.c_int32()
    _check(lib().dcdf_oracle_suggest_fraction(enc, _p(a), shape, strides, C.byref(kind), C.byref(bits)))
    return ("Round" if kind.value else "Precise", bits.value)


def min_max(a, bits=0, round_=False):
    a, enc, shape, strides = _arr_desc(a)
    mn = np.empty(a.shape[0], np.int64)
    mx = np.empty(a.shape[0], np.int64)
    _check(lib().dcdf_oracle_min_max(enc, _p(a), shape, strides, bits, int(round_), _p(mn), _p(mx)))
    return mn, mx


# ------------------------------------------------------------------ bitmap / dac
def bitmap_from_bytes(byts, length):
    b = np.asarray(byts, dtype=np.uint8)
    words = np.zeros(len(b) // 4 + 2, np.uint32)
    index = np.zeros(len(b) // 16 + 2, np.uint32)
    nw, ni = C.c_uint32(), C.c_uint32()
    _check(lib().dcdf_oracle_bitmap_from_bytes(_p(b), C.c_uint64(len(b)), C.c_uint64(length), _p(words), _p(index),
                                               C.byref(nw), C.byref(ni)))
    return words[:nw.value].tolist(), index[:ni.value].tolist()


def bitmap_push_rank(bits):
    b = np.asarray(bits, dtype=np.uint8)
    n = len(b)
    get = np.zeros(n, np.uint8)
    rank = np.zeros(n + 1, np.uint32)
    cap = 16 + n
    ser = np.zeros(cap, np.uint8)
    ln = C.c_uint64()
    _check(lib().dcdf_oracle_bitmap_push_rank(_p(b), C.c_uint64(n), _p(get), _p(rank), _p(ser), C.c_uint64(cap), C.byref(ln)))
    return get, rank, bytes(ser[:ln.value])


def dac_roundtrip(values):
    v = np.asarray(values, dtype=np.int64)
    out = np.zeros(len(v), np.int64)
    cap = 64 + 12 * len(v)
    ser = np.zeros(cap, np.uint8)
    ln, nl = C.c_uint64(), C.c_uint32()
    _check(lib().dcdf_oracle_dac_roundtrip(_p(v), C.c_uint64(len(v)), _p(out), _p(ser), C.c_uint64(cap), C.byref(ln), C.byref(nl)))
    return out, bytes(ser[:ln.value]), nl.value


# ------------------------------------------------------------------ handles
class _Handle:
    def __init__(self, ptr):
        self.ptr = C.c_void_p(ptr)

    def __del__(self):
        try:
            if self.ptr:
                lib().dcdf_oracle_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Struct2D(_Handle):
    """A Snapshot, or a Log together with the Snapshot it refers to (testing.rs:341-388 helpers)."""

    def bitmap(self, which=0):
        ln, nw, ni = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().dcdf_oracle_bitmap_info(self.ptr, which, C.byref(ln), C.byref(nw), C.byref(ni)))
        w = np.zeros(max(nw.value, 1), np.uint32)
        ix = np.zeros(max(ni.value, 1), np.uint32)
        _check(lib().dcdf_oracle_bitmap_words(self.ptr, which, _p(w), _p(ix)))
        return ln.value, w[:nw.value].tolist(), ix[:ni.value].tolist()

    def dac(self, which=0):
        n, nl = C.c_uint64(), C.c_uint32()
        _check(lib().dcdf_oracle_dac_len(self.ptr, which, C.byref(n), C.byref(nl)))
        out = np.zeros(max(n.value, 1), np.int64)
        _check(lib().dcdf_oracle_dac_collect(self.ptr, which, _p(out)))
        return out[:n.value].tolist(), nl.value

    def serialize(self, which=0):
        ln = C.c_uint64()
        _check(lib().dcdf_oracle_struct_serialize(self.ptr, which, None, C.c_uint64(0), C.byref(ln)))
        buf = np.zeros(ln.value, np.uint8)
        _check(lib().dcdf_oracle_struct_serialize(self.ptr, which, _p(buf), C.c_uint64(ln.value), C.byref(ln)))
        return bytes(buf)

    def get(self, row, col):
        out = C.c_int64()
        _check(lib().dcdf_oracle_struct_get(self.ptr, C.c_int64(row), C.c_int64(col), C.byref(out)))
        return out.value

    def window(self, top, bottom, left, right):
        out = np.zeros((abs(bottom - top), abs(right - left)), np.int64)
        _check(lib().dcdf_oracle_struct_window(self.ptr, C.c_int64(top), C.c_int64(bottom), C.c_int64(left), C.c_int64(right), _p(out)))
        return out

    def search(self, top, bottom, left, right, lower, upper):
        cap = abs(bottom - top) * abs(right - left) + 1
        out = np.zeros((cap, 2), np.int64)
        n = C.c_uint64()
        _check(lib().dcdf_oracle_struct_search(self.ptr, C.c_int64(top), C.c_int64(bottom), C.c_int64(left), C.c_int64(right),
                                               C.c_int64(lower), C.c_int64(upper), _p(out), C.c_uint64(cap), C.byref(n)))
        return [tuple(x) for x in out[:n.value].tolist()]


def snapshot_build(data, k=2):
    d = np.ascontiguousarray(data, dtype=np.int64)
    h = C.c_void_p()
    _check(lib().dcdf_oracle_snapshot_build_i64(_p(d), C.c_int64(d.shape[0]), C.c_int64(d.shape[1]), k, C.byref(h)))
    return Struct2D(h.value)


def log_build(s, t, k=2):
    s = np.ascontiguousarray(s, dtype=np.int64)
    t = np.ascontiguousarray(t, dtype=np.int64)
    h = C.c_void_p()
    _check(lib().dcdf_oracle_log_build_i64(_p(s), _p(t), C.c_int64(s.shape[0]), C.c_int64(s.shape[1]), k, C.byref(h)))
    return Struct2D(h.value)


class ChunkH(_Handle):
    def __init__(self, ptr, node=-1, owner=None, stats=None):
        super().__init__(ptr)
        self.node = node
        self.owner = owner  # keep the superchunk alive when this is a view into it
        self.stats = stats

    def __del__(self):
        if self.owner is None:
            super().__del__()

    def serialize(self):
        ln = C.c_uint64()
        _check(lib().dcdf_oracle_chunk_serialize(self.ptr, self.node, None, C.c_uint64(0), C.byref(ln)))
        buf = np.zeros(ln.value, np.uint8)
        _check(lib().dcdf_oracle_chunk_serialize(self.ptr, self.node, _p(buf), C.c_uint64(ln.value), C.byref(ln)))
        return bytes(buf)

    def info(self):
        shape = (C.c_int64 * 3)()
        enc, fb, nb = C.c_int32(), C.c_int32(), C.c_uint32()
        _check(lib().dcdf_oracle_chunk_info(self.ptr, self.node, shape, C.byref(enc), C.byref(fb), C.byref(nb)))
        return dict(shape=tuple(shape), encoding=enc.value, fractional_bits=fb.value, n_blocks=nb.value)

    def block_instants(self):
        nb = self.info()["n_blocks"]
        out = np.zeros(nb, np.uint32)
        _check(lib().dcdf_oracle_chunk_block_instants(self.ptr, self.node, _p(out)))
        return out.tolist()

    def get_batch(self, irc):
        irc = np.ascontiguousarray(irc, dtype=np.int64).reshape(-1, 3)
        out = np.zeros(len(irc), np.int64)
        _check(lib().dcdf_oracle_chunk_get_batch(self.ptr, self.node, C.c_uint64(len(irc)), _p(irc), _p(out)))
        return out

    def cell(self, start, end, row, col):
        out = np.zeros(abs(end - start), np.int64)
        _check(lib().dcdf_oracle_chunk_cell(self.ptr, self.node, C.c_int64(start), C.c_int64(end), C.c_int64(row), C.c_int64(col), _p(out)))
        return out

    def window(self, start, end, top, bottom, left, right):
        cube = (C.c_int64 * 6)(start, end, top, bottom, left, right)
        out = np.zeros((abs(end - start), abs(bottom - top), abs(right - left)), np.int64)
        _check(lib().dcdf_oracle_chunk_window(self.ptr, self.node, cube, _p(out)))
        return out

    def search(self, start, end, top, bottom, left, right, lower, upper):
        cube = (C.c_int64 * 6)(start, end, top, bottom, left, right)
        cap = abs(end - start) * abs(bottom - top) * abs(right - left) + 1
        out = np.zeros((cap, 3), np.int64)
        n = C.c_uint64()
        _check(lib().dcdf_oracle_chunk_search(self.ptr, self.node, cube, C.c_int64(lower), C.c_int64(upper), _p(out), C.c_uint64(cap), C.byref(n)))
        return out[:n.value].copy()


def chunk_build(a, k=2, fractional_bits=0, round_=False):
    a, enc, shape, strides = _arr_desc(a)
    h = C.c_void_p()
    st = Stats()
    _check(lib().dcdf_oracle_chunk_build(enc, _p(a), shape, strides, k, fractional_bits, int(round_), C.byref(h), C.byref(st)))
    return ChunkH(h.value, stats=dict(size=st.size, snapshots=st.snapshots, logs=st.logs))


def chunk_open(data):
    b = np.frombuffer(bytes(data), dtype=np.uint8)
    h = C.c_void_p()
    _check(lib().dcdf_oracle_chunk_open(_p(b), C.c_uint64(len(b)), C.byref(h)))
    return ChunkH(h.value)


class SuperH(_Handle):
    def n_nodes(self):
        n = C.c_uint32()
        _check(lib().dcdf_oracle_super_n_nodes(self.ptr, C.byref(n)))
        return n.value

    def node_info(self, node):
        info = NodeInfo()
        _check(lib().dcdf_oracle_super_node_info(self.ptr, node, C.byref(info)))
        return info

    def node_refs(self, node):
        n = self.node_info(node).n_refs
        kinds = np.zeros(n, np.int32)
        child = np.zeros(n, np.int32)
        _check(lib().dcdf_oracle_super_node_refs(self.ptr, node, _p(kinds), _p(child)))
        return kinds, child

    def node_bytes(self, node, which=0):
        ln = C.c_uint64()
        _check(lib().dcdf_oracle_super_node_bytes(self.ptr, node, which, None, C.c_uint64(0), C.byref(ln)))
        buf = np.zeros(ln.value, np.uint8)
        _check(lib().dcdf_oracle_super_node_bytes(self.ptr, node, which, _p(buf), C.c_uint64(ln.value), C.byref(ln)))
        return bytes(buf)

    def chunk(self, node):
        return ChunkH(self.ptr.value, node=node, owner=self)

    def get_batch(self, irc):
        irc = np.ascontiguousarray(irc, dtype=np.int64).reshape(-1, 3)
        out = np.zeros(len(irc), np.int64)
        bits = np.zeros(len(irc), np.int32)
        _check(lib().dcdf_oracle_super_get_batch(self.ptr, C.c_uint64(len(irc)), _p(irc), _p(out), _p(bits)))
        return out, bits

    def save(self):
        h = C.c_void_p()
        _check(lib().dcdf_oracle_super_save(self.ptr, C.byref(h)))
        return SavedH(h.value)

    def window_raw(self, start, end, top, bottom, left, right):
        cube = (C.c_int64 * 6)(start, end, top, bottom, left, right)
        out = np.zeros((abs(end - start), abs(bottom - top), abs(right - left)), np.int64)
        _check(lib().dcdf_oracle_super_window_raw(self.ptr, cube, _p(out)))
        return out

    def cell_batch(self, queries, sectors=False, want_values=True):
        """Superchunk::fill_cell for n x (start, end, row, col) -> (list of raw fixed series, sector count or None)."""
        q = np.ascontiguousarray(queries, dtype=np.int64).reshape(-1, 4)
        lens = np.abs(q[:, 1] - q[:, 0]).astype(np.uint64)
        off = np.zeros(len(q) + 1, np.uint64)
        np.cumsum(lens, out=off[1:])
        out = np.zeros(int(off[-1]) if want_values else 1, np.int64)
        sec = C.c_uint64()
        _check(lib().dcdf_oracle_super_cell_batch(self.ptr, C.c_uint64(len(q)), _p(q), _p(off), _p(out) if want_values else None,
                                                  C.byref(sec) if sectors else None))
        series = [out[int(off[i]):int(off[i + 1])] for i in range(len(q))] if want_values else None
        return series, (sec.value if sectors else None)

    def search_batch(self, cubes, lower, upper, sectors=False, want_cells=True):
        """Superchunk::search per window -> (counts, [n,3] cells or None, sector count or None)."""
        cubes = np.ascontiguousarray(cubes, dtype=np.int64).reshape(-1, 6)
        n = len(cubes)
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lower, np.int64), (n,)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(upper, np.int64), (n,)))
        counts = np.zeros(n, np.uint64)
        found, sec = C.c_uint64(), C.c_uint64()
        fn = lib().dcdf_oracle_super_search_batch
        _check(fn(self.ptr, C.c_uint64(n), _p(cubes), _p(lo), _p(hi), _p(counts), None, C.c_uint64(0), C.byref(found),
                  C.byref(sec) if sectors else None))
        cells = None
        if want_cells:
            cells = np.zeros((found.value, 3), np.int64)
            if found.value:
                _check(fn(self.ptr, C.c_uint64(n), _p(cubes), _p(lo), _p(hi), _p(counts), _p(cells), C.c_uint64(found.value), C.byref(found), None))
        return counts, cells, (sec.value if sectors else None)

    def window_f32(self, start, end, top, bottom, left, right):
        cube = (C.c_int64 * 6)(start, end, top, bottom, left, right)
        out = np.zeros((abs(end - start), abs(bottom - top), abs(right - left)), np.float32)
        _check(lib().dcdf_oracle_super_window_f32(self.ptr, cube, _p(out)))
        return out


class SavedH(_Handle):
    """What Superchunk::build + Resolver::save store (testing.rs MemoryMapper): objects in first-save order."""

    def __del__(self):
        try:
            if self.ptr:
                lib().dcdf_oracle_saved_free(self.ptr)
                self.ptr = None
        except Exception:
            pass

    def nodes(self):
        n = C.c_uint32()
        lib().dcdf_oracle_saved_count(self.ptr, C.byref(n))
        out = []
        for i in range(n.value):
            cid = np.zeros(36, np.uint8)
            t, ln = C.c_int32(), C.c_uint64()
            _check(lib().dcdf_oracle_saved_node(self.ptr, i, _p(cid), C.byref(t), None, C.c_uint64(0), C.byref(ln)))
            buf = np.zeros(max(ln.value, 1), np.uint8)
            _check(lib().dcdf_oracle_saved_node(self.ptr, i, _p(cid), C.byref(t), _p(buf), C.c_uint64(ln.value), C.byref(ln)))
            out.append((bytes(cid), t.value, bytes(buf[:ln.value])))
        return out

    def stats(self):
        size = C.c_uint64()
        el, ex, sn, lg = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().dcdf_oracle_saved_stats(self.ptr, C.byref(size), C.byref(el), C.byref(ex), C.byref(sn), C.byref(lg))
        return dict(size=size.value, elided=el.value, external=ex.value, snapshots=sn.value, logs=lg.value)


def sha256(data):
    b = np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(32, np.uint8)
    lib().dcdf_oracle_sha256(_p(b), C.c_uint64(len(data)), _p(out))
    return bytes(out)


def superchunk_build(a, levels, k=2, fractional_bits=0, round_=False, compute_bits=True):
    a, enc, shape, strides = _arr_desc(a)
    lv = (C.c_uint32 * len(levels))(*levels)
    h = C.c_void_p()
    _check(lib().dcdf_oracle_superchunk_build(enc, _p(a), shape, strides, lv, len(levels), k, fractional_bits,
                                              int(round_), int(compute_bits), C.byref(h)))
    return SuperH(h.value)


def bench_superchunk(a, levels, k=2, fractional_bits=0, round_=False, repeats=1):
    a, enc, shape, strides = _arr_desc(a)
    lv = (C.c_uint32 * len(levels))(*levels)
    sec, nbytes = C.c_double(), C.c_uint64()
    _check(lib().dcdf_oracle_bench_superchunk(enc, _p(a), shape, strides, lv, len(levels), k, fractional_bits,
                                              int(round_), 1, repeats, C.byref(sec), C.byref(nbytes)))
    return sec.value, nbytes.value
