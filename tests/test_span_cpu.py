"""Span tree: dcdf_b200.span against the step-by-step restatement of dataset.rs:834-987 / span.rs:50-112 in
oracle/span_oracle.py, and both against a closed-form description of the tree the reference ends up with."""
import hashlib
import os
import random
import struct
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import span_oracle as so  # noqa: E402

from dcdf_b200 import span as sp  # noqa: E402

ROWS, COLS, F32 = 11, 17, 32


def fake_cid(i):
    return b"\x01\x12\x12\x20" + hashlib.sha256(b"chunk %d" % i).digest()


def closed_form(store, lengths, cids, chunk_size, span_size):
    """The tree after all appends: as few levels as hold the chunks, every node filled left to right."""
    def node_bytes(stride, instants, kids):
        return struct.pack(">HIBBBIIIII", 0xDCE0, 1, 2, 3, F32, instants, ROWS, COLS, stride, len(kids)) + b"".join(kids)

    def build(level, lo, hi):
        stride = chunk_size * span_size ** (level - 1)
        if level == 1:
            kids = cids[lo:hi]
        else:
            per = span_size ** (level - 1)
            kids = [build(level - 1, a, min(a + per, hi)) for a in range(lo, hi, per)]
        b = node_bytes(stride, sum(lengths[lo:hi]), kids)
        cid = b"\x01\x12\x12\x20" + hashlib.sha256(b).digest()
        store[cid] = b
        return cid

    levels = 1
    while span_size ** levels < len(cids):
        levels += 1
    return build(levels, 0, len(cids))


def reachable(store, root, chunk_size):
    out, todo = {}, [root]
    while todo:
        cid = todo.pop()
        out[cid] = store[cid]
        s = sp.Span.from_bytes(store[cid])
        if s.stride > chunk_size:
            todo.extend(s.children)
    return out


@pytest.mark.parametrize("span_size", [2, 3, 4, 10])
def test_tree_equals_oracle_and_closed_form_chunk_by_chunk(span_size):
    chunk_size = 5
    n_max = {2: 70, 3: 90, 4: 70, 10: 120}[span_size]
    s_prod, s_orc = {}, {}
    tree = sp.SpanTree(s_prod, ROWS, COLS, chunk_size, span_size, F32)
    orc = so.OVariable(s_orc, [ROWS, COLS], chunk_size, span_size, F32)
    assert tree.root == orc.cid                                  # the empty first span (dataset.rs:127-129)
    assert tree.shape() == [0, ROWS, COLS] and tree.tail() is None
    cids = [fake_cid(i) for i in range(n_max)]
    for n in range(1, n_max + 1):
        tree.append(cids[n - 1], chunk_size)
        root = tree.commit()
        orc.append([(cids[n - 1], chunk_size)], False)
        assert root == orc.cid, n
        s_model = {}
        assert root == closed_form(s_model, [chunk_size] * n, cids[:n], chunk_size, span_size), n
        mine = reachable(s_prod, root, chunk_size)
        assert mine == reachable(s_orc, root, chunk_size) == s_model
        assert tree.shape() == [n * chunk_size, ROWS, COLS]
    assert tree.chunks() == cids


def test_batches_tail_updates_and_reload():
    rng = random.Random(7)
    for trial in range(40):
        chunk_size, span_size = rng.choice([(4, 2), (3, 3), (6, 4)])
        s_prod, s_orc = {}, {}
        tree = sp.SpanTree(s_prod, ROWS, COLS, chunk_size, span_size, F32)
        orc = so.OVariable(s_orc, [ROWS, COLS], chunk_size, span_size, F32)
        lengths, cids, serial = [], [], 0
        for batch in range(rng.randint(1, 9)):
            instants = rng.randint(1, 5 * chunk_size)
            tail = tree.tail()
            assert (tail[0] if tail else None) == orc.tail_data()
            update = tail is not None
            if update:                                           # Dataset::append_*: prepend the incomplete tail (dataset.rs:283-297)
                assert tail == (cids[-1], lengths[-1])
                instants += lengths.pop()
                cids.pop()
            new = []
            for start in range(0, instants, chunk_size):
                serial += 1
                new.append((fake_cid(1000 * trial + serial), min(chunk_size, instants - start)))
            first = True
            for cid, n in new:
                tree.append(cid, n, update=update and first)
                first = False
            orc.append(new, update)
            lengths += [n for _, n in new]
            cids += [c for c, _ in new]
            assert tree.commit() == orc.cid
            s_model = {}
            assert tree.root == closed_form(s_model, lengths, cids, chunk_size, span_size)
            assert reachable(s_prod, tree.root, chunk_size) == s_model
            assert tree.shape()[0] == sum(lengths)
            if rng.random() < 0.5:                               # carry on from the stored root, as a new process would
                tree = sp.SpanTree(s_prod, ROWS, COLS, chunk_size, span_size, F32, root=tree.root)
                assert tree.shape()[0] == sum(lengths)
        assert tree.chunks() == cids
        t = 0
        for cid, n in zip(cids, lengths):
            assert tree.locate(t) == (cid, 0) and tree.locate(t + n - 1) == (cid, n - 1)
            t += n
        with pytest.raises(IndexError):
            tree.locate(t)


def test_span_checks_of_the_reference():
    store = {}
    tree = sp.SpanTree(store, ROWS, COLS, 4, 2, F32)
    with pytest.raises(sp.SpanError):
        tree.append(fake_cid(0), 5)                              # span.rs:73-78: longer than the stride
    with pytest.raises(sp.SpanError):
        tree.append(fake_cid(0), 4, update=True)                 # nothing to replace
    tree.append(fake_cid(0), 3)
    with pytest.raises(sp.SpanError):
        tree.append(fake_cid(1), 4)                              # span.rs:58-60: last subspan is not full
    tree.append(fake_cid(1), 4, update=True)
    tree.append(fake_cid(2), 1)
    assert tree.shape() == [5, ROWS, COLS] and tree.tail() == (fake_cid(2), 1)
    with pytest.raises(sp.SpanError):
        tree.append(b"short", 1)


def test_stored_span_node_layout_and_parser():
    s = sp.Span(F32, ROWS, COLS, 20, 33, [fake_cid(1), fake_cid(2)])
    b = s.to_bytes()
    # resolver.rs:126-138 header, mmstruct.rs:212 tag, span.rs:295-303 body, everything big-endian (extio.rs:196-249)
    assert b[:8] == bytes([0xDC, 0xE0, 0, 0, 0, 1, 2, 3])
    assert b[8] == 32 and struct.unpack(">5I", b[9:29]) == (33, ROWS, COLS, 20, 2)
    assert b[29:] == fake_cid(1) + fake_cid(2) and len(b) == 29 + 72
    assert sp.cid_of(b) == bytes([1, 0x12, 0x12, 0x20]) + hashlib.sha256(b).digest()    # testing.rs:172-177
    back = sp.Span.from_bytes(b)
    assert (back.encoding, back.instants, back.rows, back.cols, back.stride, back.children) == (F32, 33, ROWS, COLS, 20, s.children)
    assert back.last_instants == 13
    for bad in (b[:20], b[:-1], b"\x00" + b[1:], b[:7] + b"\x05" + b[8:], b[:8] + b"\x07" + b[9:]):
        with pytest.raises(sp.SpanError):
            sp.Span.from_bytes(bad)
    assert sp.node_kind(b) == sp.NODE_SPAN
    sc = bytes([0xDC, 0xE0, 0, 0, 0, 1, 2, 5]) + struct.pack(">IIIIBIIBB", 64, 721, 1440, 2048, 5, 32, 64, 12, 32) + b"\x00" * 8
    assert sp.superchunk_header(sc) == (64, 721, 1440, 12, 32)   # superchunk.rs:683-692
    with pytest.raises(sp.SpanError):
        sp.superchunk_header(b)
