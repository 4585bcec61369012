"""Fast-path encoder alone: a 704 x 1408 raster (11 x 22 full tiles, nothing clipped), T instants (default 704 -> 2662 units =
3 waves of 148 x 6 tiles).  Prints the CUDA-event time of the encode kernels and tile-instants per ms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcdf_b200 import Context, Superchunk, synth, _ffi
T = int(sys.argv[1]) if len(sys.argv) > 1 else 704
kw = {}
opts = []
R, Cc = 704, 1408
for a in sys.argv[2:]:
    k, v = a.split("=")
    if k in ("noise_every", "noise_mask"): kw[k] = int(v)
    elif k == "rows": R = int(v)
    elif k == "cols": Cc = int(v)
    else: opts.append((k, int(v)))
data = synth.raster(T, R, Cc, device="cuda", **kw)
ctx = Context(0)
for k, v in opts: ctx.set_option(k, v)
best = 1e9
for i in range(4):
    sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
    ms = ctx.last_kernel_ms(_ffi.KT_ENCODE)
    best = min(best, ms)
    n = ctx.get_stat("encode_units_fast"), ctx.get_stat("encode_units_general")
    tb = sc.total_bytes()
    if i == 3:
        out = sc.window(0, min(T, 128), 0, R, 0, Cc, out=torch.empty_like(data[:min(T, 128)]))
        ok = bool(torch.equal(out, data[:min(T, 128)]))
    sc.close()
n_tiles = ((R + 63) // 64) * ((Cc + 63) // 64)
print(f"encode ms {best:.3f}  units fast/general {n}  tile-instants/ms {n_tiles * T / best:.0f}  bytes {tb}  ratio {tb / data.numel() / 4:.3f}  round trip {ok}")
