"""Window decode of unrounded data (20 fractional bits: codes of two and three bytes, 64-bit expansion where needed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcdf_b200 import Context, Superchunk, synth, _ffi
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
data = synth.raster(T, 721, 1440, device="cuda", frac_bits=20)
ctx = Context(0)
sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
print("encode ms", ctx.last_kernel_ms(_ffi.KT_ENCODE), "ratio", sc.total_bytes() / data.numel() / 4)
out = torch.empty_like(data)
best = 1e9
for i in range(4):
    sc.window(0, T, 0, 721, 0, 1440, out=out)
    torch.cuda.synchronize()
    best = min(best, ctx.last_kernel_ms(_ffi.KT_WINDOW))
cells = data.numel()
alg = sc.total_bytes() + 4 * cells
print(f"decode ms {best:.3f}  G cells/s {cells / best / 1e6:.1f}  frac {alg / best / 1e6 / 6455.6:.3f}  equal {bool(torch.equal(out, data))}")
