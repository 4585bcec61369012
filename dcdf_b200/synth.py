"""Deterministic synthetic rasters for tests and bench.py.

Integer-only generator (torch int64 ops, identical on CPU and CUDA, no libm): the value of cell
(t, y, x) is q * 2^-b with q = lat(y) + season(t) + diurnal(t) + smooth(t/6, y, x) + noise(t, y, x),
|q| < 2^24 so the float32 value is exact and needs exactly <= b fractional bits.
"""
import torch

_M64 = (1 << 64) - 1


def _i64(v):
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, s):
    return (x >> s) & ((1 << (64 - s)) - 1)


def _splitmix(x):
    x = x + _i64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _hash3(seed, t, y, x):
    h = _splitmix(t + _i64(seed * 0x100000001B3))
    h = _splitmix(h ^ (y * 0x1F123BB5 + 0x7F4A7C15))
    return _splitmix(h ^ (x * 0x2545F491 + 0x165667B1))


def _triangle(v, period, amp):
    """integer triangle wave in [0, amp]"""
    p = v % period
    half = period // 2
    up = torch.where(p < half, p, period - p)
    return (up * amp) // max(half, 1)


def raster_slice(t0, t1, rows, cols, *, seed=0xDCDF0002, frac_bits=4, base=280, hourly=True, nan_ocean=False,
                 noise_every=16, noise_mask=3, scale=None, device="cpu", dtype=torch.float32):
    """float raster [t1-t0, rows, cols] of the synthetic field for instants t0..t1.

    noise_every=1 jitters every cell (SURVEY 8d's per-cell noise; the default touches every 16th cell).
    scale: multiply the finished field by this factor IN THE OUTPUT dtype (one IEEE multiply, identical on CPU and
    CUDA).  A non-dyadic factor such as 0.1 gives what un-rounded real data looks like: full mantissas, ~29
    fractional bits, fixed-point values beyond 32 bits (the int64 encode path) -- SURVEY 8d's variant v2; the same
    data with round=True and a small fractional_bits is variant v3."""
    dev = torch.device(device)
    t = torch.arange(t0, t1, device=dev, dtype=torch.int64).view(-1, 1, 1)
    y = torch.arange(rows, device=dev, dtype=torch.int64).view(1, -1, 1)
    x = torch.arange(cols, device=dev, dtype=torch.int64).view(1, 1, -1)
    unit = 1 << frac_bits
    q = torch.full((1, 1, 1), base * unit, device=dev, dtype=torch.int64)
    q = q + _triangle(y, max(rows, 2) * 2, 40 * unit) - 20 * unit          # latitude gradient
    q = q + _triangle(t, 8766 if hourly else 365, 15 * unit)               # season
    if hourly:
        q = q + _triangle(t, 24, 6 * unit)                                 # diurnal cycle
    # smooth field: bilinear interpolation of a coarse lattice (32 cells, 6 instants), integer weights
    T6 = t // 6
    Y, fy = y // 32, y % 32
    X, fx = x // 32, x % 32

    def lat(dy, dx):
        return _hash3(seed, T6, Y + dy, X + dx) & 0xFF

    s = ((32 - fy) * (32 - fx) * lat(0, 0) + (32 - fy) * fx * lat(0, 1) + fy * (32 - fx) * lat(1, 0) + fy * fx * lat(1, 1)) >> 10
    q = q + (s * unit) // 16
    h = _hash3(seed ^ 0x5555, t, y, x)
    noisy = (h & (noise_every - 1)) == 0
    q = q + torch.where(noisy, _lsr(h, 8) & noise_mask, torch.zeros_like(h))
    out = q.to(torch.float64) / float(unit)
    if nan_ocean:
        # precipitation-like: clamp at 0 and mask ~60% of 16x16 cells as NaN "ocean"
        out = torch.clamp(out - float(base), min=0.0)
        ocean = (_hash3(seed ^ 0xAAAA, torch.zeros_like(t), y // 16, x // 16) & 0xFF) < 154
        out = torch.where(ocean.expand_as(out), torch.full_like(out, float("nan")), out)
    out = out.to(dtype)
    if scale is not None:
        out = out * torch.tensor(scale, dtype=dtype, device=dev)
    return out


def raster(instants, rows, cols, *, slice_instants=64, out=None, t_start=0, **kw):
    """Whole [instants, rows, cols] raster of instants t_start .. t_start + instants, generated slice by slice
    (optionally into `out`)."""
    device = kw.get("device", "cpu")
    dtype = kw.get("dtype", torch.float32)
    if out is None:
        out = torch.empty((instants, rows, cols), device=device, dtype=dtype)
    for t0 in range(0, instants, slice_instants):
        t1 = min(t0 + slice_instants, instants)
        out[t0:t1] = raster_slice(t_start + t0, t_start + t1, rows, cols, **kw)
    return out
