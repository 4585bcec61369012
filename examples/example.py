"""The reference's examples/example.py on dcdf_b200: the same two dataset layouts (CPC precipitation, ERA5-Land), the same
init / copy / query life cycle with a HEAD file, every raster encoded and decoded on the B200.

    python examples/example.py init  cpc_precip
    python examples/example.py copy  cpc_precip --instants 400      # appends in chunk_size steps, commits every 10
    python examples/example.py query cpc_precip

Differences from the reference script: the source is a synthetic field (dClimate / IPFS are not reachable from this image), and
objects go to a directory (`dcdf_b200.DirStore`) instead of a local IPFS node.  Names, parameters and the order of calls are
the reference's (examples/example.py:33-132 CpcPrecip.factory, :181-215 Era5LandPrecip.factory, :221-325 the commands).
tests/test_example_cpu.py runs the life cycle on the host logic (stand-in codec, reduced grid); the calls underneath are
the ones tests/test_gpu_dataset.py and tests/test_gpu_variable.py run on the B200.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class CpcPrecip:
    name = "cpc_precip_global-daily"
    variable = "precip"
    shape = (360, 720)

    @staticmethod
    def factory(ctx, store):
        from dcdf_b200 import Coordinate, Dataset
        t = Coordinate.time("time", np.datetime64("1979-01-01"), np.timedelta64(1, "D"))
        lat = Coordinate.range("latitude", -89.75, 0.5, 360, np.float32)
        lon = Coordinate.range("longitude", -179.75, 0.5, 720, np.float32)
        dataset = Dataset.new(ctx, store, [t, lat, lon], CpcPrecip.shape)
        # 360x720 pads to 1024x1024 = 10 quadtree levels: 4 levels of superchunk over 64x64 (6-level) subchunks, 64 instants
        # per superchunk (1 MiB of f32 per subchunk), 20000 chunk CIDs per span -- example.py:64-132
        return dataset.add_variable("precip", span_size=20000, chunk_size=64, k2_levels=[4, 6])

    @staticmethod
    def source(start, stop, device="cpu"):
        """A daily, non-negative, mostly-zero field with 1/8 mm resolution (what CPC precipitation looks like)."""
        from dcdf_b200 import synth
        f = synth.raster_slice(start, stop, *CpcPrecip.shape, hourly=False, base=0, device=device)
        return (f.clamp_(min=0) if hasattr(f, "clamp_") else np.maximum(f, 0))


class Era5LandPrecip:
    name = "era5_land_precip-hourly"
    variable = "tp"
    shape = (1801, 3600)

    @staticmethod
    def factory(ctx, store):
        from dcdf_b200 import Coordinate, Dataset
        t = Coordinate.time("time", np.datetime64("1981-01-01"), np.timedelta64(1, "h"))
        lat = Coordinate.range("latitude", -90.0, 0.1, 1801, np.float64)
        lon = Coordinate.range("longitude", -180.0, 0.1, 3600, np.float64)
        dataset = Dataset.new(ctx, store, [t, lat, lon], Era5LandPrecip.shape)
        # 1801x3600 pads to 4096x4096 = 12 levels: [2, 4, 6] -- example.py:199-210
        return dataset.add_variable("tp", span_size=20000, chunk_size=64, k2_levels=[2, 4, 6])

    @staticmethod
    def source(start, stop, device="cpu"):
        from dcdf_b200 import synth
        return synth.raster_slice(start, stop, *Era5LandPrecip.shape, device=device)


DATASETS = {"cpc_precip": CpcPrecip, "era5_land_precip": Era5LandPrecip}


def head_file(kind, root):
    return os.path.join(root, f".{kind.name}_head")


def save_head(path, cid, message="Success."):
    with open(path, "w") as out:
        print(cid.hex(), file=out)
    print(f"{message} New head saved to {path}.")


def initialize_dataset(kind, ctx, store, root):
    """example.py:221-247: an empty dataset with its one variable, committed; the CID goes to the HEAD file."""
    path = head_file(kind, root)
    if os.path.exists(path):
        raise SystemExit(f"Dataset already initialized. HEAD is stored at {path}")
    dataset = kind.factory(ctx, store)
    save_head(path, dataset.commit())
    return dataset


def load_head(kind, ctx, store, root):
    from dcdf_b200 import Dataset
    path = head_file(kind, root)
    if not os.path.exists(path):
        raise SystemExit(f"Dataset doesn't exist. Have you initalized it? HEAD should be stored at {path}")
    return Dataset.load(ctx, store, bytes.fromhex(open(path).read().strip()))


def copy_data(kind, ctx, store, root, n_instants, commit_every=10, device="cpu"):
    """example.py:257-325: append one chunk width at a time, commit every `commit_every` appends and at the end.  Every
    append returns a new Dataset (the data is immutable); an incomplete last chunk is re-encoded by the next append."""
    dataset = load_head(kind, ctx, store, root)
    dst = dataset.variables[0]
    written = dst.shape[0]
    commit_count = commit_every
    for index in range(written, written + n_instants, dst.chunk_size):
        stop = min(index + dst.chunk_size, written + n_instants)
        data = kind.source(index, stop, device)
        dataset = dataset.append(dst.name, data if data.is_cuda else data.numpy())    # CUDA tensors are taken zero-copy
        if commit_count == 1:
            save_head(head_file(kind, root), dataset.commit(), "Incremental progress saved.")
            commit_count = commit_every
        else:
            commit_count -= 1
        dst = dataset.variables[0]
        print(f"Copied {dst.shape[0]}/{written + n_instants}")
    save_head(head_file(kind, root), dataset.commit())
    return dataset


def query(kind, ctx, store, root, with_search=True):
    """A cell series, one day's window and a value-range search through the stored dataset, checked against the source."""
    dataset = load_head(kind, ctx, store, root)
    var = getattr(dataset, kind.variable)
    instants, rows, cols = var.shape
    print(f"{var.name}: shape {var.shape}, k2_levels {var.k2_levels}, chunk_size {var.chunk_size}, prev {dataset.prev and dataset.prev.hex()[:16]}")
    if instants == 0:
        return dataset
    src = np.asarray(kind.source(0, instants))
    r, c = rows // 3, cols // 2
    series = var[:, r, c].data
    assert np.array_equal(series, src[:, r, c])
    print(f"cell ({r}, {c}): {instants} instants from {dataset.coordinates[0][0]}, max {series.max():.3f} at "
          f"{dataset.coordinates[0][int(series.argmax())]}")
    t = instants // 2
    r0, r1, c0, c1 = rows // 4, rows // 2, cols // 3, 2 * cols // 3
    window = var[t, r0:r1, c0:c1].data
    assert np.array_equal(window, src[t, r0:r1, c0:c1])
    lat, lon = dataset.coordinates[1][r0:r1], dataset.coordinates[2][c0:c1]
    print(f"window at {dataset.coordinates[0][t]}: lat {lat[0]:.2f}..{lat[-1]:.2f}, lon {lon[0]:.2f}..{lon[-1]:.2f}, mean {window.mean():.4f}")
    lo, hi = max(float(np.quantile(src[t], 0.98)), 0.125), float(src[t].max())
    if with_search and lo <= hi:
        hits = var.search(t, t + 1, 0, rows, 0, cols, lo, hi)
        want = np.argwhere((src[t] >= lo) & (src[t] <= hi))
        assert len(hits) == len(want) and {tuple(h[1:]) for h in hits.tolist()} == {tuple(w) for w in want.tolist()}
        print(f"search [{lo:.3f}, {hi:.3f}] at instant {t}: {len(hits)} cells")
    return dataset


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("command", choices=["init", "copy", "query"])
    ap.add_argument("dataset", choices=sorted(DATASETS))
    ap.add_argument("--root", default="./dcdf_store", help="directory of the object store and the HEAD files")
    ap.add_argument("--instants", type=int, default=128)
    ap.add_argument("--commit-every", type=int, default=10)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    from dcdf_b200 import Context, DirStore
    kind = DATASETS[args.dataset]
    store = DirStore(os.path.join(args.root, "objects"))
    ctx = Context(args.device)            # fails loudly without a CUDA device: there is no CPU codec in dcdf_b200
    try:
        if args.command == "init":
            initialize_dataset(kind, ctx, store, args.root)
        elif args.command == "copy":
            copy_data(kind, ctx, store, args.root, args.instants, args.commit_every, device=f"cuda:{args.device}")
        else:
            query(kind, ctx, store, args.root)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
