"""Short encode run used for ncu captures: 2 slices of the C2 grid (552 units), one warm-up + one timed call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcdf_b200 import Context, Superchunk, synth, _ffi
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 721
C = int(sys.argv[3]) if len(sys.argv) > 3 else 1440
data = synth.raster(T, R, C, device="cuda")
ctx = Context(0)
for opt in sys.argv[4:]:          # e.g. no_fast_encode=1
    k, v = opt.split("=")
    ctx.set_option(k, int(v))
for i in range(2):
    sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
    print("encode ms", ctx.last_kernel_ms(_ffi.KT_ENCODE), "stats ms", ctx.last_kernel_ms(_ffi.KT_STATS), "bytes", sc.total_bytes(),
          "fast units", ctx.get_stat("encode_units_fast"), "general", ctx.get_stat("encode_units_general"))
    if i == 0:
        sc.close()
out = torch.empty((T, R, C), device="cuda", dtype=torch.float32)
sc.window(0, T, 0, R, 0, C, out=out)
torch.cuda.synchronize()
print("decode ms", ctx.last_kernel_ms(_ffi.KT_WINDOW), "equal", bool(torch.equal(out, data)))
