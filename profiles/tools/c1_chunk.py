import sys, time; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from dcdf_b200 import Context, Chunk, synth, _ffi
ctx = Context(0)
data = synth.raster_slice(0, 100, 256, 256).numpy()
for i in range(3):
    t0=time.perf_counter(); ch = Chunk.build(ctx, data, fractional_bits=4); t1=time.perf_counter()
    w = ch.window(0,100,0,256,0,256); t2=time.perf_counter()
    print("build s", t1-t0, "window s", t2-t1, "kernel ms", ctx.last_kernel_ms(_ffi.KT_WINDOW), "equal", np.array_equal(w, data), "bytes", ch.size())
