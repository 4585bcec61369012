"""The span tree above the time-slice superchunks of a variable, host metadata only.

  stored Span node      span.rs:283-303 (`Span::save_to`) behind `Resolver::save` (resolver.rs:126-138) and
                        `MMStruct3::save_to` (mmstruct.rs:211-214): magic, version, NODE_MMSTRUCT3, NODE_SPAN, encoding,
                        shape, stride, child CIDs
  CIDs                  testing.rs:170-183: CIDv1, codec 0x12, SHA2-256 multihash of the stored bytes
  growth                dataset.rs:834-957 (`Variable::append`, `create_open_span`, `tail_spans`, `save_spans`),
                        span.rs:50-110 (`Span::append` / `Span::update` and their checks)
  routing               span.rs:274-279 (`find_span`): child = instant / stride

The reference walks the rightmost branch of the tree from the store for every chunk it appends and saves every
intermediate version of every span on the way.  Here the rightmost branch is kept in memory while chunks are appended and
hashed once, bottom-up, in `commit()`: the root CID and every node reachable from it are the reference's (nodes are
content addressed, so the versions nobody links to any more make no difference).  No arithmetic of the hot path lives
here -- the tree only orders the CIDs that `dcdf_superchunk_save` produced on the device.
"""
import hashlib
import struct

MAGIC_AND_VERSION = b"\xDC\xE0\x00\x00\x00\x01"           # MAGIC_NUMBER = 0xDCDF + 1, FORMAT_VERSION = 1 (resolver.rs:24-27)
NODE_MMSTRUCT3, NODE_SPAN, NODE_SUBCHUNK, NODE_SUPERCHUNK = 2, 3, 4, 5   # node.rs:9-15
CID_BYTES = 36
ENCODINGS = {"int32": 4, "int64": 8, "float32": 32, "float64": 64}      # mmstruct.rs:37-43
_SPAN_HEAD = struct.Struct(">6sBBBIIIII")


class SpanError(ValueError):
    """The conditions the reference panics on (span.rs:52-78)."""


def cid_of(stored):
    """CID of a stored node: Cid::new_v1(0x12, Multihash::wrap(0x12, sha256(bytes))) (testing.rs:172-177)."""
    return b"\x01\x12\x12\x20" + hashlib.sha256(stored).digest()


class Span:
    """One node of the tree: `stride` instants per child slot, children in time order."""
    __slots__ = ("encoding", "instants", "rows", "cols", "stride", "children")

    def __init__(self, encoding, rows, cols, stride, instants=0, children=()):
        self.encoding, self.rows, self.cols, self.stride = int(encoding), int(rows), int(cols), int(stride)
        self.instants, self.children = int(instants), list(children)

    def to_bytes(self):
        head = _SPAN_HEAD.pack(MAGIC_AND_VERSION, NODE_MMSTRUCT3, NODE_SPAN, self.encoding, self.instants, self.rows, self.cols,
                               self.stride, len(self.children))
        return head + b"".join(self.children)

    @classmethod
    def from_bytes(cls, stored):
        if len(stored) < _SPAN_HEAD.size:
            raise SpanError("truncated span node")
        magic, outer, inner, enc, instants, rows, cols, stride, n = _SPAN_HEAD.unpack_from(stored)
        if magic != MAGIC_AND_VERSION or outer != NODE_MMSTRUCT3 or inner != NODE_SPAN:
            raise SpanError("not a stored Span node")
        if enc not in ENCODINGS.values():
            raise SpanError(f"unknown encoding {enc}")                      # MMEncoding::try_from, mmstruct.rs:46-60
        body = stored[_SPAN_HEAD.size:]
        if len(body) != n * CID_BYTES:
            raise SpanError("span node length does not match its child count")
        kids = [bytes(body[i * CID_BYTES:(i + 1) * CID_BYTES]) for i in range(n)]
        if any(k[:4] != b"\x01\x12\x12\x20" for k in kids):
            raise SpanError("only CIDv1 with a SHA2-256 multihash is supported")
        return cls(enc, rows, cols, stride, instants, kids)

    @property
    def last_instants(self):
        """Length of the last child; every child before it is full (span.rs:52-62 keeps that invariant)."""
        return self.instants - (len(self.children) - 1) * self.stride if self.children else 0


def node_kind(stored):
    """NODE_SPAN / NODE_SUBCHUNK / NODE_SUPERCHUNK of a stored MMStruct3 node (mmstruct.rs:229-241)."""
    if len(stored) < 8 or stored[:6] != MAGIC_AND_VERSION or stored[6] != NODE_MMSTRUCT3:
        raise SpanError("not a stored MMStruct3 node")
    return stored[7]


def superchunk_header(stored):
    """(instants, rows, cols, fractional_bits, encoding) of a stored superchunk node (superchunk.rs:683-692)."""
    if node_kind(stored) != NODE_SUPERCHUNK or len(stored) < 35:
        raise SpanError("not a stored Superchunk node")
    instants, rows, cols = struct.unpack_from(">III", stored, 8)
    return instants, rows, cols, stored[33], stored[34]


class SpanTree:
    """The tree of one variable.  `store` maps CID -> stored bytes.  `append` takes the CID of a saved time-slice
    superchunk; `commit` writes the changed spans and moves `root`."""

    def __init__(self, store, rows, cols, chunk_size, span_size, encoding, root=None):
        if chunk_size < 1 or span_size < 2:
            raise SpanError("chunk_size must be positive and span_size at least 2")
        self.store, self.chunk_size, self.span_size = store, int(chunk_size), int(span_size)
        self.rows, self.cols, self.encoding = int(rows), int(cols), int(encoding)
        if root is None:                                   # Dataset::add_variable, dataset.rs:127-129: one empty span
            first = Span(self.encoding, rows, cols, chunk_size)
            root = self._save(first)
        self.root = bytes(root)
        self._path = None                                  # rightmost branch, root first; None = as stored
        self._dirty = False

    # ------------------------------------------------------------------ store access
    def _save(self, span):
        stored = span.to_bytes()
        cid = cid_of(stored)
        self.store[cid] = stored
        return cid

    def _load(self, cid):
        stored = self.store.get(cid) if hasattr(self.store, "get") else self.store[cid]
        if stored is None:
            raise KeyError(cid)                            # Error::NotFound, resolver.rs:143
        span = Span.from_bytes(bytes(stored))
        if (span.rows, span.cols) != (self.rows, self.cols):
            raise SpanError("span shape does not match the variable")
        return span

    def _branch(self):
        """Variable::tail_spans (dataset.rs:961-975)."""
        if self._path is None:
            path = [self._load(self.root)]
            while path[-1].stride > self.chunk_size:
                if not path[-1].children:
                    raise SpanError("an upper span without children")
                path.append(self._load(path[-1].children[-1]))
            if path[-1].stride != self.chunk_size:
                raise SpanError("span strides do not end at the chunk size")
            self._path = path
        return self._path

    # ------------------------------------------------------------------ growth
    @property
    def instants(self):
        path = self._branch()
        total = path[-1].instants
        for up in reversed(path[:-1]):                     # what Span::update will make of it (span.rs:98-111)
            total += (len(up.children) - 1) * up.stride
        return total

    def tail(self):
        """(CID, instants) of the last chunk if it is incomplete, else None (Variable::tail_data, dataset.rs:937-957)."""
        bottom = self._branch()[-1]
        if bottom.children and bottom.last_instants < self.chunk_size:
            return bottom.children[-1], bottom.last_instants
        return None

    def append(self, chunk_cid, instants, update=False):
        """One turn of the loop of Variable::append (dataset.rs:853-873): open a new bottom span if the current one is
        full, then Span::append, or Span::update for the re-encoded tail."""
        chunk_cid, instants = bytes(chunk_cid), int(instants)
        if len(chunk_cid) != CID_BYTES:
            raise SpanError("a CID here is 36 bytes (CIDv1, SHA2-256)")
        path = self._branch()
        if path[-1].instants == self.span_size * path[-1].stride:
            self._open_span()
            path = self._path
        bottom = path[-1]
        if update:                                          # Span::update: drop the last child first
            if not bottom.children:
                raise SpanError("nothing to update in an empty span")
            bottom.children.pop()
            bottom.instants = len(bottom.children) * bottom.stride
        if bottom.children and bottom.last_instants != bottom.stride:
            raise SpanError("Can't append to span when last subspan is not full")                   # span.rs:58-60
        if instants > bottom.stride:
            raise SpanError(f"Attempt to add subspan with length ({instants}) greater than stride ({bottom.stride})")  # span.rs:73-78
        bottom.children.append(chunk_cid)
        bottom.instants += instants
        self._dirty = True

    def _open_span(self):
        """Variable::create_open_span (dataset.rs:880-935): a new empty bottom span under the lowest ancestor with a free
        slot, or under a new root one level up."""
        self.commit()                                       # the full branch is final (dataset.rs:858-859)
        path = self._path
        room = len(path) - 2
        while room >= 0 and len(path[room].children) == self.span_size:
            room -= 1
        if room < 0:
            old_root = path[0]
            top = Span(self.encoding, self.rows, self.cols, self.span_size * old_root.stride, old_root.instants, [self.root])
            path = [top]
        else:
            path = path[:room + 1]
        while path[-1].stride > self.chunk_size:
            up = path[-1]
            if up.children and self._child_instants(up) != up.stride:
                raise SpanError("Can't append to span when last subspan is not full")
            up.children.append(None)                        # takes the child's CID in commit()
            path.append(Span(self.encoding, self.rows, self.cols, up.stride // self.span_size))
        self._path = path
        self._dirty = True

    def _child_instants(self, up):
        return up.instants - (len(up.children) - 1) * up.stride

    def commit(self):
        """Variable::save_spans (dataset.rs:977-987): re-link the branch bottom-up and save it; returns the root CID."""
        if self._dirty:
            path = self._path
            cid = self._save(path[-1])
            for k in range(len(path) - 2, -1, -1):
                up, child = path[k], path[k + 1]
                up.children[-1] = cid
                up.instants = (len(up.children) - 1) * up.stride + child.instants
                cid = self._save(up)
            self.root = cid
            self._dirty = False
        return self.root

    # ------------------------------------------------------------------ reading
    def shape(self):
        return [self.instants, self.rows, self.cols]

    def chunks(self):
        """CIDs of the time-slice chunks in time order (depth-first walk of the committed tree)."""
        self.commit()
        out = []

        def walk(cid):
            span = self._load(cid)
            if span.stride == self.chunk_size:
                out.extend(span.children)
            else:
                for c in span.children:
                    walk(c)

        walk(self.root)
        return out

    def locate(self, instant):
        """(chunk CID, instant inside that chunk): Span::get's descent (span.rs:126-139, find_span :274-279)."""
        self.commit()
        span = self._load(self.root)
        if not 0 <= instant < span.instants:
            raise IndexError(instant)
        while True:
            slot, instant = divmod(instant, span.stride)
            cid = span.children[slot]
            if span.stride == self.chunk_size:
                return cid, instant
            span = self._load(cid)
