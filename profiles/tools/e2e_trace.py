"""Per-phase timing of one 512-instant group through pinned host buffers (DCDF_TRACE=1 for the library's own phases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dcdf_b200 import Context, Superchunk, synth
T = 512
data = synth.raster(T, 721, 1440, device="cuda")
host = torch.empty((T, 721, 1440), dtype=torch.float32).pin_memory()
host.copy_(data); torch.cuda.synchronize()
hn = host.numpy()
ctx = Context(0)
out = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for rep in range(3):
    t0 = time.perf_counter()
    sc = Superchunk.build(ctx, hn, [5, 6], compute_bits=True, chunk_size=64)
    t1 = time.perf_counter()
    n = 0
    for s in range(sc.n_slices):
        info = sc.info(s)
        for which, b in ((0, info.chunk_bytes), (1, info.max_dac_bytes), (2, info.min_dac_bytes)):
            ctx.check(ctx._lib.dcdf_superchunk_bytes(ctx._h, sc._h, s, which, out.data_ptr(), b, 0)); n += b
    t2 = time.perf_counter()
    sc.close()
    t3 = time.perf_counter()
    print(f"rep {rep}: build {1e3*(t1-t0):.1f} ms  d2h {1e3*(t2-t1):.1f} ms ({n/1e6:.0f} MB)  close {1e3*(t3-t2):.1f} ms", flush=True)
