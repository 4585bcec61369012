"""Randomised cross-check of the batch kernels (profiles/tools/stress_batches.py): search counts from the shared-decode
pass against the per-window kernel (which the other tests pin to the oracle), cell series from the tile decoder against the
per-cell walks and the input -- on rasters with uniform instants, constant offsets between instants (single-node and `equal`
Logs, Logs over single-node Snapshots: the reference's root-test rule, log.rs:527-548), NaNs, clipped tiles, nested trees and
64-bit values.  Seed 2 holds the case that exposed the rule for Logs over a single-node Snapshot.

Run on the B200 box with `pytest -m gpu`.
"""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_batch_kernels_agree_with_the_per_query_kernels(seed):
    spec = importlib.util.spec_from_file_location("stress_batches", os.path.join(ROOT, "profiles", "tools", "stress_batches.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(seed, 9)


@pytest.mark.parametrize("seed,fast", [(1, False), (11, True)])
def test_encoder_bytes_on_temporal_patterns(seed, fast):
    """profiles/tools/stress_encode.py: Superchunk::build byte for byte against the oracle on rasters with uniform
    instants, constant offsets between instants, NaNs, clipped tiles, nested trees and 64-bit values; fast = f32 rasters with
    64-side leaves and long slices, most of whose units take the fast-path encoder."""
    spec = importlib.util.spec_from_file_location("stress_encode", os.path.join(ROOT, "profiles", "tools", "stress_encode.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(seed, 6, fast)
