"""Host-side phase times of one full C2 build (ctx option `trace`)."""
import sys, time, torch
sys.path.insert(0, ".")
from dcdf_b200 import Context, Superchunk, synth, _ffi
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8760
dev = torch.device("cuda", 0)
ctx = Context(0)
data = synth.raster(T, 721, 1440, device=dev)
torch.cuda.synchronize()
for i in range(3):
    Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64).close()
torch.cuda.synchronize()
t0 = time.perf_counter()
sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
torch.cuda.synchronize()
print("build ms", (time.perf_counter() - t0) * 1e3, "encode kernels", ctx.last_kernel_ms(_ffi.KT_ENCODE), "stats", ctx.last_kernel_ms(_ffi.KT_STATS),
      "gather", ctx.last_kernel_ms(_ffi.KT_GATHER))
sc.close()
ctx.set_option("trace", 1)
Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64).close()
