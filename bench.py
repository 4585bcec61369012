#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

  metric  : encode raw-input GB/s (+ window-decode cells/s in "decode") on N B200s
  workload: configs[1] -- ERA5-shaped 721x1440 grid, 8760 hourly f32 instants, full Superchunk encode
            (k2_levels [5,6], chunk_size 64: 137 time slices x 276 in-bounds 64x64 subchunks), synthetic data
  step    : one pass of the hot path over the whole raster:
            compute_fractional_bits + Superchunk::build for every 64-instant slice (dataset.rs:834-851)

`value` is measured with the raster resident in HBM; `e2e` runs the same call with pinned HOST buffers
(H2D of the raster and D2H of every encoded byte inside the timed region).  N > 1: every rank encodes its
own year of the same grid (independent time spans, no collective; weak scaling).

--impl reference times the CPU oracle (a line-faithful C++ port of the reference's Rust path; the Rust
itself cannot be built in this image) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRID = (721, 1440)
INSTANTS = 8760
LEVELS = [5, 6]
CHUNK_SIZE = 64
METRIC = "encode_raw_input_GBps"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _dist_init(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if args.impl == "ours" else "gloo", rank=rank, world_size=world)
    if args.impl == "ours":
        torch.cuda.set_device(local)
    return rank, world, local


def _bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and so its first-touch pinned buffers) to the CPUs NVML lists as local to the GPU:
    with N ranks on one box the end-to-end leg is bound by host memory / PCIe root-complex traffic, and a rank whose
    staging buffers sit on the other socket pays the inter-socket link twice.  Returns the number of CPUs kept (0 = left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def _max_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(x, world, device):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ------------------------------------------------------------------------------------------ CPU baseline
def _cpu_worker(job):
    import numpy as np
    import oracle_lib as orc
    path, shape = job
    a = np.load(path, mmap_mode="r")
    a = np.ascontiguousarray(a)
    sec, nbytes = orc.bench_superchunk(a, LEVELS, repeats=1)
    return sec, nbytes, a.nbytes


def cpu_baseline(instants, rows, cols, workers, seconds_budget=25.0):
    """Oracle ("port" of the reference's CPU path) on a bounded sample: each worker encodes one
    `instants`-instant slice of a rows x cols sub-grid.  Returns aggregate raw-input GB/s."""
    import tempfile

    import numpy as np
    import torch
    import oracle_lib as orc
    from dcdf_b200 import synth
    orc.build_oracle()
    a = synth.raster_slice(0, instants, rows, cols, device="cpu").numpy()
    t0 = time.perf_counter()
    if workers <= 1:
        sec, nbytes = orc.bench_superchunk(a, LEVELS_FOR(rows, cols), repeats=1)
        wall = time.perf_counter() - t0
        return a.nbytes / wall / 1e9, 1, wall, nbytes
    import multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sample.npy")
        np.save(path, a)
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker_levels, [(path, LEVELS_FOR(rows, cols))] * workers)
            wall = time.perf_counter() - t0
    return sum(r[2] for r in res) / wall / 1e9, workers, wall, res[0][1]


def LEVELS_FOR(rows, cols):
    import math
    total = max(1, math.ceil(math.log2(max(rows, cols))))
    return [max(total - 6, 1), min(6, total - max(total - 6, 1))] if total > 6 else [1, total - 1]


def _cpu_worker_levels(job):
    import numpy as np
    import oracle_lib as orc
    path, levels = job
    a = np.ascontiguousarray(np.load(path))
    sec, nbytes = orc.bench_superchunk(a, levels, repeats=1)
    return sec, nbytes, a.nbytes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the other ranks exit 0 without work
    workers = os.cpu_count() or 1
    rows, cols = GRID
    inst = CHUNK_SIZE
    # bounded sample: one 64-instant slice per worker per step over a band of the grid sized for ~15 s/step
    band_rows = min(rows, args.ref_rows)
    times, gbps = [], []
    for step in range(args.warmup + args.steps):
        v, w, wall, _ = cpu_baseline(inst, band_rows, cols, workers)
        if step >= args.warmup:
            times.append(wall)
            gbps.append(v)
    value = sum(gbps) / len(gbps)
    sample = f"{workers} workers x one {inst}-instant slice of a {band_rows}x{cols} band of the {rows}x{cols} grid per step (levels {LEVELS_FOR(band_rows, cols)})"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i64", "data": "synthetic",
        "config": {"workload": f"ERA5-shaped {GRID[0]}x{GRID[1]} f32, {INSTANTS} hourly instants per GPU, Superchunk encode k2_levels {LEVELS} chunk_size {CHUNK_SIZE} (configs[1])",
                   "sample": sample, "l2": "n/a (CPU)", "parallelism": f"{workers} host threads, one 64-instant slice each"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    from dcdf_b200 import Context, Superchunk, _ffi, synth
    rank, world, local = _dist_init(args)
    numa_cpus = _bind_to_gpu_numa_node(local) if world > 1 else 0
    dev = torch.device("cuda", local)
    rows, cols = GRID
    T = args.instants
    free, total = torch.cuda.mem_get_info(dev)
    need = 4 * T * rows * cols * 2.2
    if need > free:
        T = int(free / (4 * rows * cols * 2.2)) // CHUNK_SIZE * CHUNK_SIZE
    raw_bytes = 4 * T * rows * cols
    # every rank encodes its own year (different seed per rank)
    data = torch.empty((T, rows, cols), device=dev, dtype=torch.float32)
    synth.raster(T, rows, cols, out=data, device=dev, seed=0xDCDF0002 + rank)
    torch.cuda.synchronize()
    ctx = Context(local)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)

    def step(src):
        sc = Superchunk.build(ctx, src, LEVELS, compute_bits=True, chunk_size=CHUNK_SIZE)
        return sc

    peak, peak_kind = _peaks()
    # ---- device-resident throughput
    enc_ms, stat_ms, gather_ms = [], [], []
    s_out = 0
    for _ in range(args.warmup):
        sc = step(data)
        s_out = sc.total_bytes()
        sc.close()
    torch.cuda.synchronize()
    _barrier(world)
    launches0 = ctx.launch_count
    with ClockSampler(local) as clocks:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0.record(stream)
        for i in range(args.steps):
            sc = step(data)
            enc_ms.append(ctx.last_kernel_ms(_ffi.KT_ENCODE))
            stat_ms.append(ctx.last_kernel_ms(_ffi.KT_STATS))
            gather_ms.append(ctx.last_kernel_ms(_ffi.KT_GATHER))
            s_out = sc.total_bytes()
            if i < args.steps - 1:
                sc.close()  # the last result is kept for the decode leg
        ev1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        _barrier(world)
    launches = ctx.launch_count - launches0
    dev_ms = ev0.elapsed_time(ev1)
    step_ms = _max_over_ranks(max(dev_ms, wall * 1e3) / args.steps, world, dev)
    total_raw = _sum_over_ranks(raw_bytes, world, dev)
    value = total_raw / (step_ms * 1e-3) / 1e9

    # ---- window decode of every unit (full extent), device-resident output
    dec = None
    try:
        out = torch.empty((CHUNK_SIZE * 8, rows, cols), device=dev, dtype=torch.float32)
        tspan = min(out.shape[0], T)
        sc.window(0, tspan, 0, rows, 0, cols, out=out[:tspan])
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        reps = 3
        for _ in range(reps):
            sc.window(0, tspan, 0, rows, 0, cols, out=out[:tspan])
        d1.record(stream)
        torch.cuda.synchronize()
        dms = d0.elapsed_time(d1) / reps
        ok = bool(torch.equal(out[:tspan], data[:tspan]))
        cells = tspan * rows * cols
        s_in = s_out * tspan / T
        dtraffic = None
        try:  # DRAM bytes per decoded cell of k_window_tiles4 from the committed ncu capture, scaled to this launch
            dtraffic = json.load(open(os.path.join(ROOT, "profiles", "r1_decode_traffic.json")))["dram_bytes_per_cell"] * cells
        except Exception:
            pass
        dec = {"metric": "window_decode_cells_per_s", "value": _sum_over_ranks(cells, world, dev) / (_max_over_ranks(dms, world, dev) * 1e-3),
               "unit": "cells/s", "ms": dms, "cells": cells, "round_trip_equal": ok,
               "roofline": {"bound": "hbm", "achieved": (s_in + 4 * cells) / (dms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": (s_in + 4 * cells) / (dms * 1e-3) / 1e9 / peak, "traffic": dtraffic,
                            "frac_of_nominal_8TBps": (s_in + 4 * cells) / (dms * 1e-3) / 1e9 / 8000.0,
                            "kernel": "k_window_tiles4<int> (one launch: every 64x64 tile of 8 slices, f32 output resident in HBM)",
                            "algorithmic_bytes": int(s_in + 4 * cells), "kernel_ms": ctx.last_kernel_ms(_ffi.KT_WINDOW)}}
        del out
    except Exception as e:  # decode is the second half of the metric; never hide an encode number behind it
        dec = {"error": str(e)}
    # ---- batched cell time series (configs[2] style) and value-range search (configs[3] style) on the same chunks
    queries = None
    try:
        import numpy as np
        rng = np.random.default_rng(7)
        nq = 4096
        q = np.stack([np.zeros(nq, np.int64), np.full(nq, T, np.int64), rng.integers(0, rows, nq), rng.integers(0, cols, nq)], axis=1)
        sc.cell_batch(q)  # warm-up at full size: the context's staging buffers grow here, not inside the timed call
        tq = time.perf_counter()
        series = sc.cell_batch(q)
        t_cell = time.perf_counter() - tq
        k_cell = ctx.last_kernel_ms(_ffi.KT_CELL)
        okc = bool(np.array_equal(series[5], data[:, int(q[5, 2]), int(q[5, 3])].cpu().numpy()))
        nw = 2048
        side = rng.integers(8, 257, nw)
        top = rng.integers(0, rows - 8, nw); left = rng.integers(0, cols - 8, nw)
        t0s = rng.integers(0, max(T - CHUNK_SIZE, 1), nw)
        cubes = np.stack([t0s, np.minimum(t0s + CHUNK_SIZE, T), top, np.minimum(top + side, rows), left, np.minimum(left + side, cols)], axis=1)
        lo_v = rng.integers(270 * 32, 300 * 32, nw)      # fixed point with 4 fractional bits: value * 32 + 1
        sc.search_batch(cubes, lo_v, lo_v + 48)  # warm-up at full size (result / cache buffers are allocated here)
        ts = time.perf_counter()
        counts, cells = sc.search_batch(cubes, lo_v, lo_v + 48)
        t_search = time.perf_counter() - ts
        k_search = ctx.last_kernel_ms(_ffi.KT_SEARCH)
        vol = int(((cubes[:, 1] - cubes[:, 0]) * (cubes[:, 3] - cubes[:, 2]) * (cubes[:, 5] - cubes[:, 4])).sum())
        queries = {"cell_series": {"series": nq, "cells": int(nq * T), "cells_per_s_kernel": nq * T / (k_cell * 1e-3),
                                   "cells_per_s_e2e": nq * T / t_cell, "kernel_ms": k_cell, "matches_input": okc},
                   "search": {"windows": nw, "cells_scanned": vol, "matches": int(counts.sum()), "windows_per_s_e2e": nw / t_search,
                              "cells_scanned_per_s_kernel": vol / (k_search * 1e-3), "kernel_ms": k_search}}
    except Exception as e:
        queries = {"error": str(e)}
    sc.close()

    # ---- end to end through host buffers (pinned): H2D raster + encode + D2H of all encoded bytes
    e2e = None
    if not args.no_e2e:
        Te = min(T, args.e2e_instants)
        try:
            host = torch.empty((Te, rows, cols), dtype=torch.float32).pin_memory()
            host.copy_(data[:Te])
            torch.cuda.synchronize()
            host_np = host.numpy()
            # Two contexts on two host threads work on alternate groups of slices, so the H2D copy of one group
            # overlaps the kernels and the D2H copies of the other (the public API, used the way a caller would).
            group = 8 * CHUNK_SIZE
            spans = [(g, min(g + group, Te)) for g in range(0, Te, group)]
            workers = [Context(local) for _ in range(2)]
            out_bufs = [None, None]

            def work(w):
                cw = workers[w]
                nbytes = 0
                for gi in range(w, len(spans), 2):
                    a0, a1 = spans[gi]
                    sc_ = Superchunk.build(cw, host_np[a0:a1], LEVELS, compute_bits=True, chunk_size=CHUNK_SIZE)
                    for s in range(sc_.n_slices):
                        info = sc_.info(s)
                        for which, n in ((0, info.chunk_bytes), (1, info.max_dac_bytes), (2, info.min_dac_bytes)):
                            if out_bufs[w] is None or out_bufs[w].numel() < n:
                                out_bufs[w] = torch.empty(int(n * 1.5) + 1024, dtype=torch.uint8).pin_memory()
                            cw.check(cw._lib.dcdf_superchunk_bytes(cw._h, sc_._h, s, which, out_bufs[w].data_ptr(), n, 0))
                            nbytes += n
                    sc_.close()
                done[w] = nbytes

            done = [0, 0]

            def e2e_step():
                th = [threading.Thread(target=work, args=(w,)) for w in range(2)]
                for t_ in th:
                    t_.start()
                for t_ in th:
                    t_.join()
                return done[0] + done[1]

            e2e_step()
            torch.cuda.synchronize()
            _barrier(world)
            t0 = time.perf_counter()
            n_e2e = max(1, min(args.steps, 2))
            d2h = 0
            for _ in range(n_e2e):
                d2h = e2e_step()
            torch.cuda.synchronize()
            t_e2e = _max_over_ranks((time.perf_counter() - t0) / n_e2e, world, dev)
            raw_e = 4 * Te * rows * cols
            e2e = {"value": _sum_over_ranks(raw_e, world, dev) / t_e2e / 1e9, "unit": "GB/s", "h2d_bytes_per_step": raw_e,
                   "d2h_bytes_per_step": int(d2h), "instants": Te, "ms_per_step": t_e2e * 1e3,
                   "how": "pinned host raster -> Superchunk.build on 2 contexts / host threads over alternate 512-instant groups -> every encoded byte copied back to pinned host memory"}
            for cw in workers:
                cw.close()
            del host
        except Exception as e:
            e2e = {"error": str(e)}

    # ---- CPU baseline beside it (rank 0, N == 1 only, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            v, cores, wall_c, _ = cpu_baseline(CHUNK_SIZE, args.cpu_rows, cols, 1)
            cpu = {"value": v, "unit": "GB/s", "cores": cores, "kind": "port",
                   "sample": f"one {CHUNK_SIZE}-instant slice of a {args.cpu_rows}x{cols} band, single thread ({wall_c:.1f} s)"}
        except Exception as e:
            cpu = {"error": str(e)}

    if rank == 0:
        k_ms = sum(enc_ms) / len(enc_ms)
        algo = raw_bytes + s_out
        # DRAM traffic of the dominant kernel from the committed ncu capture (bytes per (full 64x64 unit, instant)),
        # scaled to the full units of this launch; the clipped ring (k_encode_tiles, 8 % of the cells) is not in it.
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            full_units = (rows // 64) * (cols // 64) * ((T + CHUNK_SIZE - 1) // CHUNK_SIZE)
            traffic = tj["dram_bytes_per_unit_instant"] * full_units * CHUNK_SIZE * (T / (((T + CHUNK_SIZE - 1) // CHUNK_SIZE) * CHUNK_SIZE))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i64",
            "data": "synthetic",
            "config": {"workload": f"ERA5-shaped {rows}x{cols} f32, {T} hourly instants per GPU, Superchunk encode k2_levels {LEVELS} chunk_size {CHUNK_SIZE} (configs[1])",
                       "l2": "inputs larger than L2 (no flush needed)", "encoded_bytes": int(s_out), "ratio": s_out / raw_bytes,
                       "parallelism": f"{world} independent time spans",
                       "host_binding": f"rank threads pinned to the {numa_cpus} CPUs local to their GPU (NVML)" if numa_cpus else "none"},
            "roofline": {"bound": "hbm", "achieved": algo / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (k_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "frac_of_nominal_8TBps": algo / (k_ms * 1e-3) / 1e9 / 8000.0,
                         "kernel": "k_encode_v4 (64-side tiles: full on the main stream, clipped on a second stream; k_encode_tiles only for the corner tile with a 32-side tree); one timed region",
                         "traffic_note": "ncu dram bytes of k_encode_v4 per (unit, instant) x full units of this launch (profiles/r1_traffic.json)",
                         "kernel_ms": k_ms, "stats_kernel_ms": sum(stat_ms) / len(stat_ms),
                         "gather_ms": sum(gather_ms) / len(gather_ms), "algorithmic_bytes": int(algo)},
            "cpu_baseline": cpu, "e2e": e2e, "decode": dec, "queries": queries, "gpu_launches": int(launches), "clocks": clocks.summary(),
            "wall_ms_per_step": wall * 1e3 / args.steps, "device_ms_per_step": dev_ms / args.steps,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instants", type=int, default=INSTANTS)
    ap.add_argument("--e2e-instants", type=int, default=INSTANTS)
    ap.add_argument("--cpu-rows", type=int, default=721)
    ap.add_argument("--ref-rows", type=int, default=256)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
