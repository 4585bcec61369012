"""Window decode alone: T instants of the C2 grid (721 x 1440), full-extent window, best of 5; optional per-cell noise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcdf_b200 import Context, Superchunk, synth, _ffi
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    kw[k] = int(v)
data = synth.raster(T, 721, 1440, device="cuda", **kw)
ctx = Context(0)
sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
out = torch.empty_like(data)
best = 1e9
for i in range(5):
    sc.window(0, T, 0, 721, 0, 1440, out=out)
    torch.cuda.synchronize()
    best = min(best, ctx.last_kernel_ms(_ffi.KT_WINDOW))
cells = data.numel()
alg = sc.total_bytes() + 4 * cells
print(f"decode ms {best:.3f}  G cells/s {cells / best / 1e6:.1f}  algorithmic GB/s {alg / best / 1e6:.0f}  frac {alg / best / 1e6 / 6455.6:.3f}  equal {bool(torch.equal(out, data))}")
# a window that starts inside a block and is clipped on every side
o2 = sc.window(37, min(T, 150), 5, 700, 3, 1401)
import numpy as np
print("partial equal", bool(np.array_equal(np.asarray(o2), data[37:min(T, 150), 5:700, 3:1401].cpu().numpy(), equal_nan=True)))
