"""Pins the CPU oracle against every known-answer / golden vector the reference's own tests hold
for the hot path.  CPU only.  Each test cites the reference test it replays.
"""
import itertools
import math

import numpy as np
import pytest

import fixtures as fx
import oracle_lib as orc


# ----------------------------------------------------------------------------- fixed.rs
def test_to_fixed_known_answers():  # fixed.rs:208-223
    assert orc.to_fixed(1.5, 1) == 7
    assert orc.to_fixed(-1.5, 1) == -5
    assert orc.to_fixed(1.5, 8) == 769
    assert orc.to_fixed(0.0625, 4) == 3
    assert orc.to_fixed(0.0, 16) == 1
    assert orc.to_fixed(-0.0, 16) == 1
    for dt in (np.float32, np.float64):
        assert orc.to_fixed(1.5, 1, dtype=dt) == 7


def test_to_fixed_round():  # fixed.rs:225-246
    assert orc.to_fixed(1.5, 1, True) == 7
    assert orc.to_fixed(1.5, 8, True) == 769
    assert orc.to_fixed(0.0625, 4, True) == 3
    assert orc.to_fixed(0.0625, 3, True) == 3
    assert orc.to_fixed(0.0625, 2, True) == 1
    assert orc.to_fixed(0.1, 16, True) == 6554 * 2 + 1
    assert orc.to_fixed(0.0, 16, True) == 1
    assert orc.to_fixed(-0.0, 16, True) == 1


def test_to_fixed_negative_inexact_truncates():  # SURVEY Appendix B #2 (fixed.rs:47,62-65)
    assert orc.to_fixed(-0.1, 3, True) == 0
    assert orc.to_fixed(-0.1, 3, False) == 0  # not even rejected


def test_from_fixed():  # fixed.rs:248-258
    assert orc.from_fixed(7, 1) == 1.5
    assert orc.from_fixed(-5, 1) == -1.5
    assert orc.from_fixed(769, 8, np.float64) == 1.5
    assert orc.from_fixed(3, 4) == 0.0625
    assert orc.from_fixed(1, 13) == 0.0
    assert abs(orc.from_fixed(6554 * 2 + 1, 16, np.float64) - 0.1) < 1e-5
    assert math.isnan(orc.from_fixed(0, 5))


@pytest.mark.parametrize("n,bits,code", [
    (0.0625, 3, 2), (1.0625, 3, 2),              # fixed.rs:260-272 loss of precision
    (float("inf"), 14, 1), (float("-inf"), 14, 1),  # fixed.rs:280-292
    (1.5e100, 1, 3),                             # fixed.rs:294-299 overflow
])
def test_to_fixed_panics(n, bits, code):
    with pytest.raises(orc.OracleError) as e:
        orc.to_fixed(n, bits, False)
    assert e.value.code == code


def test_to_fixed_nan():  # fixed.rs:274-278
    assert orc.to_fixed(float("nan"), 12) == 0


def test_round_trip_lots_of_fractional_bits():  # fixed.rs:301-309 (issue #5)
    n = np.float32(1024.1)  # `n` is inferred as f32 in the reference test
    assert orc.from_fixed(orc.to_fixed(n, 34, False, dtype=np.float32), 34, np.float32) == n


def test_suggest_fraction():  # fixed.rs:311-401
    assert orc.suggest_fraction(fx.fixed_array()) == ("Precise", 3)
    assert orc.suggest_fraction(np.array([[[16.0, 1.0 / 16.0]]])) == ("Precise", 4)
    assert orc.suggest_fraction(np.array([[[16.0, 0.1]]])) == ("Precise", 55)
    assert orc.suggest_fraction(np.array([[[316.0, 0.1]]])) == ("Round", 53)
    nan = np.float32("nan")
    assert orc.suggest_fraction(np.array([[[nan, 16.0, nan, 1.0 / 16.0]]], dtype=np.float32)) == ("Precise", 4)
    assert orc.suggest_fraction(np.array([[[nan, nan, nan, nan]]], dtype=np.float32)) == ("Precise", 0)


def test_suggest_fraction_strided_view_and_negatives():
    a = fx.farray(16, 6)
    assert orc.suggest_fraction(a[:, 3:11, 2:9]) == ("Precise", 3)
    # max-not-max-abs and saturating cast: large negatives contribute 0 bits (fixed.rs:126,150-152)
    assert orc.suggest_fraction(np.array([[[-100.5, 2.0]]])) == ("Precise", 0)
    assert orc.suggest_fraction(np.array([[[-1.5, 2.0]]])) == ("Precise", 1)


def test_min_max_float_nan_quirk():  # mmbuffer.rs:465-499
    nan = float("nan")
    a = np.array([[[nan, 2.0, 1.0, 3.0]], [[2.0, nan, 1.0, 3.0]], [[nan, nan, nan, nan]]], dtype=np.float32)
    mn, mx = orc.min_max(a, bits=1)
    assert mn.tolist() == [orc.to_fixed(1.0, 1), 0, 0]
    assert mx.tolist() == [orc.to_fixed(3.0, 1), orc.to_fixed(3.0, 1), 0]


# ----------------------------------------------------------------------------- bitmap.rs
def test_bitmap_from_bytes():  # bitmap.rs:261-284
    words, index = orc.bitmap_from_bytes([99, 104, 114, 105, 115], 36)
    assert words == [1667789417, 1929379840] and index == []
    words, index = orc.bitmap_from_bytes([99, 104, 114, 105, 115, 0, 0, 0, 99, 104, 114, 105, 115, 0, 0, 0, 128], 129)
    assert words == [1667789417, 1929379840, 1667789417, 1929379840, 1 << 31]
    assert index == [40]


def test_bitmap_get_rank():  # bitmap.rs:319-347, 386-392 (differential vs naive popcount)
    answers = [1, 0, 1, 0, 1, 0, 1, 0, 0, 0, 1]
    get, rank, ser = orc.bitmap_push_rank(answers)
    assert get.tolist() == answers
    assert rank.tolist() == [0] + np.cumsum(answers).tolist()
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2, 1 << 16).astype(np.uint8)
    get, rank, ser = orc.bitmap_push_rank(bits)
    assert np.array_equal(get, bits)
    assert np.array_equal(rank, np.concatenate([[0], np.cumsum(bits)]))
    n = len(bits)
    assert len(ser) == 8 + 4 * (n // 128) + 4 * ((n + 31) // 32)  # bitmap.rs:166-172, :408-410
    assert ser[:8] == n.to_bytes(4, "big") + (4).to_bytes(4, "big")


# ----------------------------------------------------------------------------- dac.rs
def test_dac_known_answers():  # dac.rs:163-199
    data = [0, 2, -3, -2 ** 9, 2 ** 17 + 1, -2 ** 30 - 42]
    out, ser, n_levels = orc.dac_roundtrip(data)
    assert out.tolist() == data
    assert n_levels == 4
    out, ser, n_levels = orc.dac_roundtrip([-512])
    assert out.tolist() == [-512]
    out, ser, n_levels = orc.dac_roundtrip([])
    assert ser == b"\x00" and n_levels == 0  # dac.rs:124-128
    assert orc.lib().dcdf_oracle_dac_get_empty() == 0


def test_dac_extremes():
    data = [2 ** 62, -2 ** 62, 2 ** 63 - 1, -2 ** 63, 0, 1, -1, 127, -128, 63, 64, -64, -65]
    out, ser, n_levels = orc.dac_roundtrip(data)
    assert out.tolist() == data and n_levels == 8


# ----------------------------------------------------------------------------- snapshot.rs
def test_snapshot_build_golden():  # snapshot.rs:538-559
    s = orc.snapshot_build(fx.array8_3()[0])
    length, words, index = s.bitmap(0)
    assert length == 17 and words == [0b11110101001001011000000000000000]
    mx, _ = s.dac(0)
    assert mx == [9, 0, 3, 4, 5, 0, 2, 3, 3, 0, 3, 3, 3, 0, 0, 1, 0, 0, 1, 2, 2, 0, 0, 1, 1, 0, 1, 0,
                  0, 1, 0, 2, 2, 1, 1, 0, 0, 2, 0, 2, 1]
    mn, _ = s.dac(1)
    assert mn == [2, 3, 0, 1, 2, 0, 0, 0, 0, 0]


def test_snapshot_fill_values_in_padded_quadbox():  # snapshot.rs:561-573
    data = np.full((9, 9), 5, np.int64)
    data[:8, :8] = fx.array8_3()[0]
    s = orc.snapshot_build(data)
    assert s.bitmap(0)[0] == 21
    assert s.get(8, 8) == 5


def test_snapshot_single_node_tree():  # snapshot.rs:586-599
    s = orc.snapshot_build(np.full((16, 16), 42, np.int64))
    length, words, _ = s.bitmap(0)
    assert len(words) == 1
    mx, nl = s.dac(0)
    assert mx == [42]
    assert s.dac(1) == ([], 0)
    assert all(s.get(r, c) == 42 for r in range(16) for c in range(16))
    assert np.array_equal(s.window(3, 11, 2, 16), np.full((8, 14), 42))


@pytest.mark.parametrize("data,k", [(fx.array8_3()[0], 2), (fx.array9_3()[0], 2), (fx.array9_3()[0], 3)])
def test_snapshot_get_window_search_exhaustive(data, k):  # snapshot.rs:575-807
    n = data.shape[0]
    s = orc.snapshot_build(data, k)
    for r in range(n):
        for c in range(n):
            assert s.get(r, c) == data[r, c]
    ser = s.serialize()
    assert ser[0] == k
    for top in range(n):
        for bottom in range(top + 1, n + 1):
            for left in range(n):
                for right in range(left + 1, n + 1):
                    assert np.array_equal(s.window(top, bottom, left, right), data[top:bottom, left:right])
    step = 1 if n == 8 else 2
    for top in range(0, n, step):
        for bottom in range(top + 1, n + 1, step):
            for left in range(0, n, step):
                for right in range(left + 1, n + 1, step):
                    for lower in range(4, 10):
                        for upper in range(lower, 10):
                            got = s.search(top, bottom, left, right, lower, upper)
                            assert len(got) == len(set(got))
                            assert set(got) == fx.brute_search2(data, top, bottom, left, right, lower, upper)


def test_rect_rearranges_bounds():  # geom.rs:13-25, snapshot.rs:810-852
    data = fx.array8_3()[0]
    s = orc.snapshot_build(data)
    assert np.array_equal(s.window(6, 2, 7, 1), data[2:6, 1:7])


# ----------------------------------------------------------------------------- log.rs
def test_log_build_golden():  # log.rs:901-938
    a = fx.array8_3()
    l1 = orc.log_build(a[0], a[1])
    assert l1.bitmap(0)[:2] == (17, [0b10111001000010010000000000000000])
    assert l1.bitmap(1)[:2] == (10, [0b10001010000000000000000000000000])
    assert l1.dac(0)[0] == [0, 0, 1, 0, 1, 1, -1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0]
    assert l1.dac(1)[0] == [0, 0, 0, 0, 0, 1, 0]
    l2 = orc.log_build(a[0], a[2])
    assert l2.bitmap(0)[:2] == (21, [0b11111000010100001001000000000000])
    assert l2.bitmap(1)[:2] == (12, [0b10100010100000000000000000000000])
    assert l2.dac(0)[0] == [0, 0, 2, 0, 2, 0, 0, 1, 0, 2, 2, 1, 1, 0, 0, 0, 0, 0, 0, 2, 0, 2, 1, 1, 1, 1, 0, 1,
                            1, 1, 0, 1, 0, 2, 0, 1, 0]
    assert l2.dac(1)[0] == [1, 1, 1, 0, 0, 1, 0, 1, 0]


def test_log_fill_values_in_padded_quadbox():  # log.rs:940-956
    data = np.full((3, 9, 9), 5, np.int64)
    data[:, :8, :8] = fx.array8_3()
    data[0] = fx.array9_3()[0]
    l = orc.log_build(data[0], data[1])
    assert l.bitmap(0)[0] == 21
    assert l.get(8, 8) == 5


def _sweep_log(s, t, k, search_range=None, step=1):
    n = s.shape[0]
    l = orc.log_build(s, t, k)
    for r in range(n):
        for c in range(n):
            assert l.get(r, c) == t[r, c]
    assert len(l.serialize()) > 13
    for top in range(0, n, step):
        for bottom in range(top + 1, n + 1, step):
            for left in range(0, n, step):
                for right in range(left + 1, n + 1, step):
                    assert np.array_equal(l.window(top, bottom, left, right), t[top:bottom, left:right])
                    if search_range:
                        lo0, hi0 = search_range
                        for lower in range(lo0, hi0 + 1):
                            for upper in range(lower, hi0 + 1):
                                got = l.search(top, bottom, left, right, lower, upper)
                                assert len(got) == len(set(got))
                                assert set(got) == fx.brute_search2(t, top, bottom, left, right, lower, upper)


def test_log_sweeps_array8():  # log.rs:958-976, 1079-1120, 1238-1320
    a = fx.array8_3()
    _sweep_log(a[0], a[1], 2, (4, 9))
    _sweep_log(a[0], a[2], 2, (4, 9))


def test_log_sweeps_array9_and_k3():  # log.rs:1042-1077, 1122-1236, 1322-1480
    a = fx.array9_3()
    _sweep_log(a[0], a[1], 2, (4, 9), step=2)
    _sweep_log(a[0], a[2], 2, (4, 10), step=2)
    _sweep_log(a[0], a[1], 3, (4, 9), step=2)
    _sweep_log(a[0], a[2], 3, None)


def test_log_single_node_cases():  # log.rs:978-1040, 1482-1616
    a = fx.array8_3()
    s20 = np.full((8, 8), 20, np.int64)
    t42 = np.full((8, 8), 42, np.int64)
    _sweep_log(s20, t42, 2, (40, 44))       # both single node
    _sweep_log(s20, a[0], 2, (4, 9))        # single node snapshot
    _sweep_log(a[0], a[0], 2, (4, 9))       # equal snapshot and log
    # single-node (uniform) log over a non-uniform snapshot: get/window are right ...
    l = orc.log_build(a[0], t42)
    assert all(l.get(r, c) == 42 for r in range(8) for c in range(8))
    assert np.array_equal(l.window(0, 8, 0, 8), t42)
    # ... and search agrees with brute force on the ranges the reference tests (log.rs:1520-1557)
    for lower in range(4, 10):
        for upper in range(lower, 10):
            assert l.search(0, 8, 0, 8, lower, upper) == []


def test_log_search_reference_bug_is_reproduced():
    """Log::search_window has no special case for a single-node uniform log over
    a non-uniform snapshot (log.rs:527-548 vs :184-186).  The oracle reproduces the reference's
    (incorrect) answers so that GPU parity means parity with the reference."""
    a = fx.array8_3()
    l = orc.log_build(a[0], np.full((8, 8), 5, np.int64))
    assert len(l.search(0, 8, 0, 8, 5, 5)) == 1    # correct answer would be 64
    assert len(l.search(0, 8, 0, 8, 3, 4)) == 7    # correct answer would be 0


# ----------------------------------------------------------------------------- block.rs / chunk.rs
def test_chunk_build_roundtrip_int():  # chunk.rs:427-565 (through Chunk::build rather than hand-made blocks)
    data = fx.array8()
    ch = orc.chunk_build(data)
    info = ch.info()
    assert info["shape"] == (100, 8, 8) and info["encoding"] == 8 and info["fractional_bits"] == 0
    assert sum(ch.block_instants()) == 100
    assert ch.stats["snapshots"] == info["n_blocks"] and ch.stats["logs"] == 100 - info["n_blocks"]
    ser = ch.serialize()
    assert len(ser) == ch.stats["size"]                       # chunk.rs:572-574
    ch2 = orc.chunk_open(ser)
    assert ch2.serialize() == ser
    irc = np.array(list(itertools.product(range(100), range(8), range(8))))
    assert np.array_equal(ch2.get_batch(irc), data.reshape(-1))
    for row in range(8):                                      # chunk.rs:441-456
        for col in range(8):
            start, end = row * 6, 100 - col * 6
            assert np.array_equal(ch2.cell(start, end, row, col), data[start:end, row, col])
    for top in range(0, 8, 2):                                # chunk.rs:458-482
        for bottom in range(top + 1, 9, 2):
            for left in range(0, 8, 3):
                for right in range(left + 1, 9, 2):
                    start, end = top * left, bottom * right + 36
                    assert np.array_equal(ch2.window(start, end, top, bottom, left, right), data[start:end, top:bottom, left:right])
                    for lower, upper in ((4, 6), (7, 5), (9, 9)):  # swapped bounds chunk.rs:525-565
                        got = ch2.search(start, end, top, bottom, left, right, lower, upper)
                        lo, hi = min(lower, upper), max(lower, upper)
                        assert {tuple(x) for x in got.tolist()} == fx.brute_search3(data, start, end, top, bottom, left, right, lo, hi)


def test_chunk_heuristic_blocks():
    """Chunk::build chunk.rs:55-78: identical instants become tiny 'equal' logs; an unrelated instant
    forces a new snapshot; never more than 254 logs per block."""
    rng = np.random.default_rng(3)
    base = rng.integers(0, 50, (16, 16)).astype(np.int64)
    data = np.stack([base] * 5 + [rng.integers(1000, 2000, (16, 16)).astype(np.int64)] * 3)
    ch = orc.chunk_build(data)
    assert ch.block_instants() == [5, 3]
    data = np.stack([base] * 300)
    ch = orc.chunk_build(data)
    assert ch.block_instants() == [255, 45]                   # 254-log cap, chunk.rs:62 / block.rs:26-32


def test_chunk_float_nan_roundtrip():  # mmarray.rs:1282-1289 style: farray fixtures with 3 fractional bits
    data = fx.farray(16, 12)
    ch = orc.chunk_build(data, fractional_bits=3)
    info = ch.info()
    assert info["encoding"] == 32 and info["fractional_bits"] == 3
    w = ch.window(0, 12, 0, 16, 0, 16)
    out = orc.from_fixed_array(w, 3)
    assert np.array_equal(out, data, equal_nan=True)
    with pytest.raises(orc.OracleError) as e:
        orc.chunk_build(data, fractional_bits=2)              # loss of precision panics (fixed.rs:51)
    assert e.value.code == 2
    ch = orc.chunk_build(data, fractional_bits=2, round_=True)
    assert ch.info()["fractional_bits"] == 2


def test_chunk_bounds_and_format_errors():
    ch = orc.chunk_build(fx.array8(4))
    with pytest.raises(orc.OracleError) as e:
        ch.window(0, 5, 0, 8, 0, 8)
    assert e.value.code == 5
    ser = ch.serialize()
    with pytest.raises(orc.OracleError) as e:
        orc.chunk_open(ser[:-3])
    assert e.value.code == 6
    with pytest.raises(orc.OracleError) as e:
        orc.chunk_open(b"\x07" + ser[1:])
    assert e.value.code == 6


# ----------------------------------------------------------------------------- superchunk.rs
def _super_counts(sc, node=0):
    kinds, child = sc.node_refs(node)
    return int((kinds == 0).sum()), int((kinds == 2).sum())


def test_superchunk_structure_counts():  # superchunk.rs:1005-1019, 1068-1093, 1097-1131, 1135-1173
    # [3,0]: 8x8 with 64 1x1 "subchunks": everything is served from the min/max DAC
    sc = orc.superchunk_build(fx.array8(), [3, 0])
    assert _super_counts(sc) == (64, 0)
    irc = np.array(list(itertools.product(range(0, 100, 7), range(8), range(8))))
    vals, bits = sc.get_batch(irc)
    assert np.array_equal(vals, fx.array8()[irc[:, 0], irc[:, 1], irc[:, 2]])
    # [2,2]: 16x16 of array8 tiles -> 16 stored 4x4 subchunks
    data = fx.array(16)
    sc = orc.superchunk_build(data, [2, 2])
    assert _super_counts(sc) == (0, 16)
    info = sc.node_info(0)
    assert (info.sidelen, info.chunks_sidelen, info.subsidelen, info.levels) == (16, 4, 4, 2)
    kinds, child = sc.node_refs(0)
    distinct = {sc.node_bytes(int(c)) for c in child}
    assert len(distinct) == 4                                  # 4 distinct CIDs in the reference
    # 17x17 with [2,3]: 32-side, 4x4 grid of 8x8 subchunks: 9 in bounds; the reference counts
    # 8 external / 8 elided (the 1x1 corner is uniform per instant -> elided)
    data = fx.array(17)
    sc = orc.superchunk_build(data, [2, 3])
    assert _super_counts(sc) == (8, 8)
    vals, bits = sc.get_batch(np.array([[5, 16, 16], [7, 3, 16], [2, 16, 9], [99, 0, 0]]))
    assert vals.tolist() == [int(data[5, 16, 16]), int(data[7, 3, 16]), int(data[2, 16, 9]), int(data[99, 0, 0])]
    # elide everything
    data = np.zeros((10, 16, 16), np.int64) + 42
    sc = orc.superchunk_build(data, [2, 2])
    assert _super_counts(sc) == (16, 0)
    assert sc.get_batch(np.array([[3, 5, 7]]))[0].tolist() == [42]


def test_superchunk_bad_levels():  # superchunk.rs:105-110
    with pytest.raises(orc.OracleError) as e:
        orc.superchunk_build(fx.array(16), [2, 3])
    assert e.value.code == 4


def test_superchunk_nested_and_window():  # superchunk.rs:1177-1196 ([1,2,2] nesting), :403-457
    data = fx.farray(32, 9)
    sc = orc.superchunk_build(data, [1, 2, 2])
    info = sc.node_info(0)
    assert (info.subsidelen, info.chunks_sidelen) == (2, 16)
    kinds, child = sc.node_refs(0)
    assert kinds.tolist() == [2, 2, 2, 2]
    assert all(sc.node_info(int(c)).kind == 0 for c in child)
    w = sc.window_f32(1, 8, 3, 30, 5, 32)
    assert np.array_equal(w, data[1:8, 3:30, 5:32], equal_nan=True)


def test_cpc_fixture_roundtrip():  # py-dcdf/tests/test_dcdf.py:357-365 (real-world f32 day, 64% NaN)
    import os
    p = os.path.join(os.path.dirname(__file__), "golden", "cpc_day_360x720.npz")
    day = np.load(p)["day"]
    assert day.shape == (360, 720)
    data = np.stack([day, day])
    kind, bits = orc.suggest_fraction(data)
    assert (kind, bits) == ("Precise", 29)                     # SURVEY section 7 "hard parts"
    sc = orc.superchunk_build(data, [4, 6])
    w = sc.window_f32(0, 2, 0, 360, 0, 720)
    assert np.array_equal(w, data, equal_nan=True)


# ----------------------------------------------------------------------------- storage side
def test_sha256_known_answers():
    """FIPS 180-4 / NIST example vectors pin the oracle's own SHA2-256."""
    import hashlib
    ka = {b"abc": "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad",
          b"": "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
          b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq": "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"}
    for m, want in ka.items():
        assert orc.sha256(m).hex() == want
    for n in (55, 56, 63, 64, 65, 119, 120, 1000):
        m = bytes((i * 7 + 3) & 0xff for i in range(n))
        assert orc.sha256(m) == hashlib.sha256(m).digest()


def test_saved_superchunk_objects_follow_the_reference_fixtures():
    """superchunk.rs:1068-1093: [2,2] over array(16) -> 16 External references to 4 distinct subchunks; :1097-1131: 8
    External / 8 Elided; CIDs are CIDv1(0x12, sha2-256) of the stored bytes (testing.rs:172-183); node framing
    resolver.rs:130-132 + mmstruct.rs:209-218."""
    import hashlib
    ref = orc.superchunk_build(fx.array(16, 100), [2, 2])
    sv = ref.save()
    nodes = sv.nodes()
    assert [t for _, t, _ in nodes] == [4, 4, 4, 4, 1, 5] and sv.stats()["external"] == 4 and sv.stats()["elided"] == 0
    for cid, t, b in nodes:
        assert cid == bytes([1, 0x12, 0x12, 0x20]) + hashlib.sha256(b).digest()
        assert b[:6] == bytes([0xDC, 0xE0, 0, 0, 0, 1]) and b[6] == (1 if t == 1 else 2)
        if t != 1:
            assert b[7] == t
    links = nodes[4][2]
    assert int.from_bytes(links[7:11], "big") == 4 and len(links) == 11 + 4 * 36
    root = nodes[-1][2]
    assert int.from_bytes(root[8:12], "big") == 100 and int.from_bytes(root[20:24], "big") == 16       # shape[0], sidelen
    n_refs = int.from_bytes(root[35:39], "big")
    assert n_refs == 16
    refs = [(root[39 + 5 * i], int.from_bytes(root[40 + 5 * i:44 + 5 * i], "big")) for i in range(16)]
    assert all(k == 2 for k, _ in refs) and sorted(set(i for _, i in refs)) == [0, 1, 2, 3]
    assert root[39 + 80:39 + 80 + 36] == nodes[4][0]                                                   # external_cid
    st = orc.superchunk_build(fx.array(17, 100), [2, 3]).save().stats()
    assert (st["external"], st["elided"]) == (3, 8) or st["elided"] == 8
