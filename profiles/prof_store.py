"""Chunk hashing alone: T instants of the C2 grid, Superchunk.save(0) (hashes every chunk node on the first call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcdf_b200 import Context, Superchunk, synth
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
data = synth.raster(T, 721, 1440, device="cuda")
ctx = Context(0)
sc = Superchunk.build(ctx, data, [5, 6], compute_bits=True, chunk_size=64)
torch.cuda.synchronize()
t0 = time.perf_counter()
nodes, stats = sc.save(0)
dt = time.perf_counter() - t0
t0 = time.perf_counter()
sc.save(1)
dt2 = time.perf_counter() - t0
import ctypes as C
h = C.c_void_p()
t0 = time.perf_counter()
ctx.check(ctx._lib.dcdf_superchunk_save(ctx._h, sc._h, 2, C.byref(h)))
dt3 = time.perf_counter() - t0
ctx._lib.dcdf_saved_free(h)
print(f"dcdf_superchunk_save alone (node assembly, no byte transfer): {dt3 * 1e3:.2f} ms")
print(f"first save {dt * 1e3:.1f} ms (hash of {sc.total_bytes() / 1e9:.2f} GB: {sc.total_bytes() / (dt - dt2) / 1e9:.1f} GB/s), later save {dt2 * 1e3:.1f} ms")
