"""Timeline of the two-context e2e pipeline of bench.py (who waits for whom)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dcdf_b200 import Context, Superchunk, synth
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 2
G = int(sys.argv[3]) if len(sys.argv) > 3 else 512
data = synth.raster(T, 721, 1440, device="cuda")
host = torch.empty((T, 721, 1440), dtype=torch.float32).pin_memory()
host.copy_(data); torch.cuda.synchronize()
del data
hn = host.numpy()
spans = [(g, min(g + G, T)) for g in range(0, T, G)]
workers = [Context(0) for _ in range(NW)]
outs = [torch.empty(1 << 30, dtype=torch.uint8).pin_memory() for _ in range(NW)]
log = []
def work(w, t00):
    cw = workers[w]
    for gi in range(w, len(spans), NW):
        a0, a1 = spans[gi]
        t0 = time.perf_counter() - t00
        sc = Superchunk.build(cw, hn[a0:a1], [5, 6], compute_bits=True, chunk_size=64)
        t1 = time.perf_counter() - t00
        for s in range(sc.n_slices):
            info = sc.info(s)
            for which, b in ((0, info.chunk_bytes), (1, info.max_dac_bytes), (2, info.min_dac_bytes)):
                cw.check(cw._lib.dcdf_superchunk_bytes(cw._h, sc._h, s, which, outs[w].data_ptr(), b, 0))
        t2 = time.perf_counter() - t00
        sc.close()
        log.append((w, gi, t0, t1, t2))
for rep in range(2):
    log.clear()
    t00 = time.perf_counter()
    th = [threading.Thread(target=work, args=(w, t00)) for w in range(NW)]
    [t.start() for t in th]; [t.join() for t in th]
    tot = time.perf_counter() - t00
    print(f"rep {rep}: {tot*1e3:.1f} ms for {T} instants -> {4*T*721*1440/tot/1e9:.1f} GB/s")
for w, gi, t0, t1, t2 in sorted(log, key=lambda r: r[2]):
    print(f"  worker {w} group {gi}: build {t0*1e3:7.1f} -> {t1*1e3:7.1f}  d2h -> {t2*1e3:7.1f}")
