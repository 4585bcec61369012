"""GPU parity at the BASELINE.json configurations and their stress variants, whole slices against the
oracle -- every chunk of the slice and both superchunk DACs, not a sample -- plus the round-2 boundary fixes:
f64 whole-bit rounding (fixed.rs:126), the arena retry keeping earlier error flags, device-resident queries
validated on the device, and stored bytes validated like Chunk::read_from.

Run on the B200 box with `pytest -m gpu`.  Nothing here reads /root/reference.
"""
import ctypes as C

import numpy as np
import pytest

import fixtures as fx
import oracle_lib as orc
from test_gpu_parity import _check_chunk, _check_superchunk, _diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from dcdf_b200 import Context
    c = Context(0)
    yield c
    c.close()


# ----------------------------------------------------------------------------- whole slices of the configs
def test_c2_whole_slice_every_chunk_and_both_dacs(ctx):
    """configs[1]: one full 64 x 721 x 1440 time slice, k2_levels [5, 6]: all 276 chunks, the Elided / External
    pattern, per-chunk fractional bits, MMStruct3Build counters and the slice's max / min DAC bytes."""
    from dcdf_b200 import synth
    data = synth.raster_slice(64, 128, 721, 1440).numpy()
    got = _check_superchunk(ctx, data, [5, 6])
    assert got.info(0).stats.external == 12 * 23
    got.close()


def test_c3_shaped_whole_slice(ctx):
    """configs[2]: PRISM / CPC-shaped 621 x 1405 daily grid with a NaN ocean (precipitation-like, clamped at 0)."""
    from dcdf_b200 import synth
    data = synth.raster_slice(0, 64, 621, 1405, hourly=False, nan_ocean=True).numpy()
    assert 0.5 < np.isnan(data).mean() < 0.7
    got = _check_superchunk(ctx, data, [5, 6])
    st = got.info(0).stats
    assert st.external + st.elided == 1024 and st.external <= 220           # 10 x 22 in-bounds tiles of 32 x 32 slots
    rng = np.random.default_rng(2)
    q = np.stack([np.zeros(200, np.int64), np.full(200, 64, np.int64), rng.integers(0, 621, 200), rng.integers(0, 1405, 200)], axis=1)
    for qi, series in zip(q, got.cell_batch(q)):
        assert np.array_equal(series, data[:, qi[2], qi[3]], equal_nan=True)
    got.close()


def test_c2_shape_unrounded_real_data_takes_the_int64_path(ctx):
    """Variant v2: full mantissas (field x 0.1f) -> Precise(31), fixed values beyond 32 bits, 5-byte DAC codes."""
    from dcdf_b200 import synth
    data = synth.raster_slice(0, 6, 721, 1440, nan_ocean=True, scale=0.1).numpy()
    kind, bits = orc.suggest_fraction(data)
    assert kind == "Precise" and bits >= 29
    got = _check_superchunk(ctx, data, [5, 6])
    assert got.info(0).fractional_bits == bits
    w = got.window(0, 6, 0, 721, 0, 1440)
    assert np.array_equal(w, data, equal_nan=True)
    got.close()


def test_c2_shape_rounding_branch(ctx):
    """Variant v3: the same data with round = Some(8) (fixed.rs:48-49, mmbuffer.rs:602-603)."""
    from dcdf_b200 import synth
    data = synth.raster_slice(0, 6, 721, 1440, nan_ocean=True, scale=0.1).numpy()
    got = _check_superchunk(ctx, data, [5, 6], fractional_bits=8, round_=True)
    assert got.info(0).fractional_bits == 8
    t = synth.raster_slice(0, 6, 721, 1440, scale=0.1).numpy()             # temperature-like, every tile stored
    _check_superchunk(ctx, t, [5, 6], fractional_bits=6, round_=True).close()
    got.close()


def test_c2_shape_per_cell_noise(ctx):
    """SURVEY 8d's generator jitters EVERY cell; the bench default only every 16th.  Both must be exact."""
    from dcdf_b200 import synth
    data = synth.raster_slice(0, 16, 721, 1440, noise_every=1).numpy()
    _check_superchunk(ctx, data, [5, 6]).close()


def test_c5_shaped_nested_slice_vs_oracle(ctx):
    """configs[4]: 1801 x 3600, k2_levels [2, 4, 6]: the whole node tree of a (short) slice against the oracle."""
    from dcdf_b200 import synth
    data = synth.raster_slice(0, 6, 1801, 3600).numpy()
    _check_superchunk(ctx, data, [2, 4, 6]).close()


# ----------------------------------------------------------------------------- a1 on f64 (fixed.rs:126)
def test_whole_bits_follow_the_host_libm_log2(ctx):
    """log2() of the last doubles below a power of two rounds UP to the integer, so floor(log2(max)) is one more than
    the exponent: 8 - 2^-50 has whole_bits 4, not 3 -> a different max_fraction_bits and different bytes."""
    from dcdf_b200 import Chunk, Superchunk
    below8 = np.nextafter(8.0, 0.0)
    for mx in (below8, np.nextafter(below8, 0.0), np.nextafter(2.0 ** 40, 0.0), 2.0 ** 40 - 22 * 2.0 ** -13, 2.0 ** 40 - 23 * 2.0 ** -13,
               np.nextafter(4.0, 0.0), 8.0, 7.5):
        a = np.array([[[mx, 0.0123], [1.0 / 3.0, -2.5]]], dtype=np.float64)        # 0.0123 needs 59 fractional bits
        assert ctx.suggest_fraction(a) == orc.suggest_fraction(a), mx
    a = np.array([[[below8, 0.0123]]], dtype=np.float64)
    assert ctx.suggest_fraction(a) == ("Round", 58)                        # ilogb would say Precise(59)
    a = np.array([[[np.nextafter(below8, 0.0), 0.0123]]], dtype=np.float64)
    assert ctx.suggest_fraction(a) == ("Precise", 59)
    rng = np.random.default_rng(4)
    data = rng.integers(0, 2 ** 20, (5, 40, 50)).astype(np.float64) / 2.0 ** 20 * 7.0
    data[2, 7, 9] = below8
    data[0, 0, 0] = 0.0123                                                 # 59 fractional bits > 62 - 4: Round(58)
    kind, bits = orc.suggest_fraction(data)
    assert ctx.suggest_fraction(data) == (kind, bits) and kind == "Round"
    _check_chunk(ctx, data, fractional_bits=bits, round_=True)
    _check_superchunk(ctx, data, [1, 5], fractional_bits=60, round_=True).close()     # per-slice and per-subchunk bits


def test_slice_level_fraction_panics_and_exact_pass_through_superchunk_build(ctx):
    """compute_fractional_bits of the whole slice (dataset.rs:842) can fail where no single subchunk does, and its
    saturating-cast corner (fixed.rs:150) needs the exact pass; both must reach Superchunk::build's result."""
    from dcdf_b200 import DcdfError, Superchunk
    rng = np.random.default_rng(12)
    data = 280.0 + rng.integers(0, 64, (4, 128, 128)).astype(np.float64) / 16.0
    data[:, :64, :64] = 0.1 + rng.integers(0, 8, (4, 64, 64)) / 8.0        # this tile alone is Precise(55); the slice is Round(53)
    assert orc.suggest_fraction(np.ascontiguousarray(data[:, :64, :64]))[0] == "Precise" and orc.suggest_fraction(data)[0] == "Round"
    with pytest.raises(orc.OracleError) as eo:
        orc.superchunk_build(data, [1, 6])
    with pytest.raises(DcdfError) as eg:
        Superchunk.build(ctx, data, [1, 6])
    assert eg.value.code == eo.value.code == 2
    _check_superchunk(ctx, data, [1, 6], fractional_bits=20, round_=True).close()
    neg = rng.integers(0, 32, (3, 128, 128)).astype(np.float64) / 16.0
    neg[1, 70, 70] = -100.5                                                # -100.5 * 2^60 saturates `as i64`
    _check_superchunk(ctx, neg, [1, 6]).close()
    _check_superchunk(ctx, neg.astype(np.float32), [1, 6]).close()


# ----------------------------------------------------------------------------- arena retry keeps earlier flags
def test_arena_overflow_retry_is_exact_and_keeps_data_errors():
    from dcdf_b200 import Context, DcdfError, Superchunk, synth
    c = Context(0)
    c.set_option("arena_hint", 4096)                                       # the first guess overflows at once
    data = synth.raster_slice(0, 12, 150, 200, noise_every=1).numpy()
    got = _check_superchunk(c, data, [2, 6], chunk_size=5)
    got.close()
    c.close()
    c = Context(0)
    c.set_option("arena_hint", 4096)
    bad = data.astype(np.float64)
    bad[3, 70, 80] = 316.0
    bad[3, 70, 81] = 0.1                                                   # Round(53) without rounding -> mmbuffer.rs:606 panics
    with pytest.raises(orc.OracleError) as eo:
        orc.superchunk_build(bad, [2, 6])
    with pytest.raises(DcdfError) as eg:
        Superchunk.build(c, bad, [2, 6])
    assert eg.value.code == eo.value.code == 2
    inf = data.copy()
    inf[5, 3, 3] = np.inf
    for cb in (False, True):
        with pytest.raises(orc.OracleError) as eo:
            orc.superchunk_build(inf, [2, 6], fractional_bits=4, compute_bits=cb)
        with pytest.raises(DcdfError) as eg:
            Superchunk.build(c, inf, [2, 6], fractional_bits=4, compute_bits=cb)
        assert eg.value.code == eo.value.code, cb
    c.close()


# ----------------------------------------------------------------------------- device-resident queries
def test_device_queries_are_bounds_checked_on_the_device(ctx):
    import torch
    from dcdf_b200 import Chunk, _ffi
    data = fx.array8(10)
    got = Chunk.build(ctx, data)
    ok = torch.tensor([[0, 0, 0], [9, 7, 7], [4, 3, 2]], dtype=torch.int64, device="cuda")
    out = torch.zeros(3, dtype=torch.int64, device="cuda")
    lib = ctx._lib
    code = lib.dcdf_chunk_get_batch(ctx._h, got._h, 3, C.c_void_p(ok.data_ptr()), C.c_void_p(out.data_ptr()), _ffi.ENC_I64, _ffi.MEM_DEVICE)
    assert code == 0 and out.cpu().tolist() == [int(data[0, 0, 0]), int(data[9, 7, 7]), int(data[4, 3, 2])]
    for bad_q in ([10, 0, 0], [0, 8, 0], [0, 0, -1], [1 << 40, 0, 0]):
        bad = torch.tensor([[1, 1, 1], bad_q], dtype=torch.int64, device="cuda")
        code = lib.dcdf_chunk_get_batch(ctx._h, got._h, 2, C.c_void_p(bad.data_ptr()), C.c_void_p(out.data_ptr()), _ffi.ENC_I64, _ffi.MEM_DEVICE)
        assert code == 5, bad_q                                            # DCDF_ERR_OUT_OF_BOUNDS, mmarray.rs:218-229
    code = lib.dcdf_chunk_get_batch(ctx._h, got._h, 3, C.c_void_p(ok.data_ptr()), C.c_void_p(out.data_ptr()), _ffi.ENC_I64, _ffi.MEM_DEVICE)
    assert code == 0                                                       # the flag does not stick
    got.close()


def test_out_buffers_are_validated_and_ordered_after_torch(ctx):
    import torch
    from dcdf_b200 import Superchunk, synth
    dev = synth.raster_slice(0, 8, 100, 130, device="cuda")
    sc = Superchunk.build(ctx, dev * 1.0, [2, 6])                          # input produced by a kernel still in flight on torch's stream
    host = dev.cpu().numpy()
    assert np.array_equal(sc.window(0, 8, 0, 100, 0, 130), host)
    with pytest.raises(ValueError):
        sc.window(0, 8, 0, 100, 0, 130, out=torch.empty(10, device="cuda"))                      # too small
    with pytest.raises(ValueError):
        sc.window(0, 8, 0, 100, 0, 130, out=torch.empty((8, 100, 130), device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError):
        sc.window(0, 8, 0, 100, 0, 130, out=torch.empty((8, 100, 260), device="cuda")[:, :, ::2])  # not contiguous
    sc.close()


# ----------------------------------------------------------------------------- stored bytes are untrusted
def test_crafted_chunk_bytes_are_rejected(ctx):
    """Chunk::read_from would panic on its bounds checks when the counts of a structure disagree; here the walks index
    with ranks of the nodemap, so dcdf_chunk_open checks those relations (and the rank directories) up front."""
    from dcdf_b200 import Chunk, DcdfError
    rng = np.random.default_rng(8)
    data = rng.integers(0, 1000, (4, 64, 64)).astype(np.int64)
    data[1] = data[0]
    data[1, 5, 5] += 1
    good = Chunk.build(ctx, data).write_to()
    Chunk.read_from(ctx, good).close()
    nm = 6 + 1 + 13                                   # chunk header, block n_instants, snapshot header -> nodemap BitMap
    nm_len = int.from_bytes(good[nm:nm + 4], "big")
    words = nm + 8 + 4 * (nm_len // 128)
    variants = []
    b = bytearray(good); b[words + 8] ^= 0x10; variants.append(("nodemap bit flipped", b))             # popcount != DAC lengths
    b = bytearray(good); b[nm + 8 + 3] ^= 0x01; variants.append(("rank directory entry off by one", b))
    dac = words + 4 * ((nm_len + 31) // 32)           # max Dac: n_levels, then level 0 BitMap
    b = bytearray(good); b[dac + 1 + 3] ^= 0x04; variants.append(("DAC level-0 length changed", b))
    for what, bad in variants:
        with pytest.raises(DcdfError) as e:
            Chunk.read_from(ctx, bytes(bad))
        assert e.value.code == 6, what
