"""TEST INFRASTRUCTURE -- CPU restatement of the reference's span tree, never imported by dcdf_b200.

Follows the reference step by step, one stored node per `save`, the rightmost branch re-read from the store for every
chunk -- the structure of dataset.rs / span.rs, not of dcdf_b200/span.py (which keeps the branch in memory and hashes once):

  Span, Span::append, Span::update, Span::save_to       span.rs:26-112, 283-303
  Resolver::save (header) / MMStruct3::save_to (tag)    resolver.rs:126-138, mmstruct.rs:205-226
  CID                                                   testing.rs:170-183
  Dataset::add_variable (first, empty span)             dataset.rs:116-140
  Variable::append, create_open_span, tail_data,
  tail_spans, save_spans                                dataset.rs:834-987

Parity pin: the reference's tests hold no byte vectors for span nodes (dataset.rs:1185-1457 checks values read back), so
this restatement is checked against a closed-form description of the tree in tests/test_span_cpu.py; chunks are opaque
here (a CID and a number of instants), exactly what the tree sees of them.
"""
import hashlib
import struct

NODE_MMSTRUCT3, NODE_SPAN = 2, 3


def resolver_save(store, body):
    """resolver.rs:126-138: MAGIC_NUMBER u16, FORMAT_VERSION u32, node type u8, then the node."""
    stored = struct.pack(">HIB", 0xDCDF + 1, 1, NODE_MMSTRUCT3) + body
    digest = hashlib.sha256(stored).digest()
    cid = bytes([1, 0x12, 0x12, 0x20]) + digest             # Cid::new_v1(SHA2_256, wrap(SHA2_256, digest))
    store[cid] = stored
    return cid


class OSpan:
    def __init__(self, shape2, stride, store, encoding):     # Span::new, span.rs:35-48
        self.shape = [0, shape2[0], shape2[1]]
        self.stride = stride
        self.spans = []
        self.store = store
        self.encoding = encoding

    def save(self):                                          # MMStruct3::save_to + Span::save_to
        body = struct.pack(">BBIIIII", NODE_SPAN, self.encoding, self.shape[0], self.shape[1], self.shape[2], self.stride,
                           len(self.spans))
        for cid in self.spans:
            body += cid
        return resolver_save(self.store, body)

    def append(self, child_cid, child_shape, shapes):        # span.rs:50-95; `shapes` stands in for get_mmstruct3(..).shape()
        if len(self.spans) > 0:
            if shapes[self.spans[-1]][0] != self.stride:
                raise RuntimeError("Can't append to span when last subspan is not full")
        if child_shape[1] != self.shape[1] or child_shape[2] != self.shape[2]:
            raise RuntimeError("Shape of subspan doesn't match shape of span")
        if child_shape[0] > self.stride:
            raise RuntimeError("Attempt to add subspan with length greater than stride")
        new = OSpan([self.shape[1], self.shape[2]], self.stride, self.store, self.encoding)
        new.shape = [self.shape[0] + child_shape[0], child_shape[1], child_shape[2]]
        new.spans = list(self.spans) + [child_cid]
        return new

    def update(self, child_cid, child_shape, shapes):        # span.rs:98-111
        spans = list(self.spans)
        spans.pop()
        tmp = OSpan([self.shape[1], self.shape[2]], self.stride, self.store, self.encoding)
        tmp.shape = [len(spans) * self.stride, self.shape[1], self.shape[2]]
        tmp.spans = spans
        return tmp.append(child_cid, child_shape, shapes)


class OVariable:
    """The span-side state of a Variable: root CID, chunk_size, span_size.  `shapes` remembers the shape of every saved
    node so that a loaded span or chunk can report it, as Resolver::get_mmstruct3(..).shape() does."""

    def __init__(self, store, shape2, chunk_size, span_size, encoding):
        self.store, self.shape2, self.chunk_size, self.span_size, self.encoding = store, list(shape2), chunk_size, span_size, encoding
        self.shapes = {}
        self.nodes = {}
        first = OSpan(shape2, chunk_size, store, encoding)   # dataset.rs:127-129
        self.cid = self._save_span(first)

    def _save_span(self, span):
        cid = span.save()
        self.shapes[cid] = list(span.shape)
        self.nodes[cid] = span
        return cid

    def _get(self, cid):
        return self.nodes[cid]

    def tail_spans(self):                                    # dataset.rs:961-975
        ancestors = []
        span = self._get(self.cid)
        while span.stride > self.chunk_size:
            cid = span.spans[-1]
            ancestors.append(span)
            span = self._get(cid)
        ancestors.append(span)
        return ancestors

    def save_spans(self, spans):                             # dataset.rs:977-987
        span = spans.pop()
        while spans:
            last = spans.pop()
            cid = self._save_span(span)                      # Span::append saves its argument (span.rs:81)
            span = last.update(cid, span.shape, self.shapes)
        self.cid = self._save_span(span)

    def create_open_span(self):                              # dataset.rs:880-935
        span = OSpan(self.shape2, self.chunk_size, self.store, self.encoding)
        spans = self.tail_spans()
        left_hand = spans.pop()
        while True:
            if spans:
                parent = spans.pop()
                if len(parent.spans) == self.span_size:
                    new_parent = OSpan(self.shape2, self.span_size * span.stride, self.store, self.encoding)
                    left_hand = parent
                    span = new_parent.append(self._save_span(span), span.shape, self.shapes)
                else:
                    span = parent.append(self._save_span(span), span.shape, self.shapes)
                    break
            else:
                new_root = OSpan(self.shape2, self.span_size * span.stride, self.store, self.encoding)
                right_hand = span
                new_root = new_root.append(self._save_span(left_hand), left_hand.shape, self.shapes)
                span = new_root.append(self._save_span(right_hand), right_hand.shape, self.shapes)
                break
        while spans:
            ancestor = spans.pop()
            span = ancestor.update(self._save_span(span), span.shape, self.shapes)
        self.cid = self._save_span(span)

    def tail_data(self):                                     # dataset.rs:937-957
        tail = self.tail_spans()[-1]
        if len(tail.spans) == 0:
            return None
        cid = tail.spans[-1]
        return cid if self.shapes[cid][0] < self.chunk_size else None

    def append(self, chunks, update):
        """dataset.rs:834-878 with the chunks already built: `chunks` = [(cid, instants)] in time order."""
        spans = self.tail_spans()
        for cid, instants in chunks:
            self.shapes[cid] = [instants, self.shape2[0], self.shape2[1]]
            span = spans.pop()
            if span.shape[0] == self.span_size * span.stride:
                spans.append(span)
                self.save_spans(spans)
                self.create_open_span()
                spans = self.tail_spans()
                span = spans.pop()
                assert len(span.spans) == 0
            if update:
                update = False
                span = span.update(cid, self.shapes[cid], self.shapes)
            else:
                span = span.append(cid, self.shapes[cid], self.shapes)
            spans.append(span)
        self.save_spans(spans)
