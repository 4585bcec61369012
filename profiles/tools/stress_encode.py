"""Randomised byte-for-byte cross-check of Superchunk::build against the oracle on rasters with uniform instants, constant
offsets between instants (single-node and `equal` Logs), NaNs, clipped tiles and 64-bit values -- the temporal patterns the
fixture-based parity tests are thin on.  f32 cases go through the fast-path encoder (k_encode_v5) where eligible."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "profiles", "tools"))
import numpy as np
from stress_batches import raster

def run(seed, n_cases, fast_only=False):
    import test_gpu_parity as tp
    from dcdf_b200 import Context
    rng = np.random.default_rng(seed)
    ctx = Context(0)
    for case in range(n_cases):
        kind = "f32" if fast_only else ("f32", "i32", "i64", "f32")[case % 4]
        levels = [[1, 6], [2, 6]][int(rng.integers(0, 2))] if fast_only else [[1, 6], [2, 6], [2, 5], [1, 4], [1, 1, 6]][int(rng.integers(0, 5))]
        side = 2 ** sum(levels)
        R, C = int(rng.integers(side // 2 + 1, side + 1)), int(rng.integers(side // 2 + 1, side + 1))
        T, cs = (int(rng.integers(20, 90)), int(rng.integers(8, 65))) if fast_only else (int(rng.integers(5, 30)), int(rng.integers(3, 17)))
        data = raster(rng, T, R, C, kind)
        try:
            got = tp._check_superchunk(ctx, data, levels, chunk_size=cs, compute_bits=(kind == "f32"))
        except AssertionError as ex:
            print(f"case {case}: {kind} levels {levels} {T}x{R}x{C} chunk_size {cs}: MISMATCH {str(ex)[:300]}")
            return False
        except Exception as ex:
            print(f"case {case}: {kind} levels {levels} {T}x{R}x{C}: skipped ({str(ex)[:70]})")
            continue
        fast = ctx.get_stat("encode_units_fast")
        print(f"case {case}: {kind} levels {levels} {T}x{R}x{C} chunk_size {cs}: bytes equal ({got.total_bytes()} B, fast units of the last slice call: {fast})", flush=True)
        got.close()
    ctx.close()
    print("all cases agree")
    return True

if __name__ == "__main__":
    ok = run(int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 10, len(sys.argv) > 3 and sys.argv[3] == "fast")
    sys.exit(0 if ok else 1)
