"""The synthetic generator is integer-only, so CPU and GPU produce identical rasters; pin a checksum."""
import numpy as np
import torch

import oracle_lib as orc
from dcdf_b200 import synth


def test_generator_is_deterministic_and_exact():
    a = synth.raster_slice(5, 9, 37, 53).numpy()
    b = synth.raster(9, 37, 53, slice_instants=4)[5:9].numpy()
    assert np.array_equal(a, b)
    assert a.dtype == np.float32 and np.isfinite(a).all()
    assert np.array_equal(a * 16, np.round(a * 16))          # exactly 4 fractional bits
    assert orc.suggest_fraction(a)[0] == "Precise" and orc.suggest_fraction(a)[1] <= 4
    n = synth.raster_slice(0, 3, 40, 40, nan_ocean=True, hourly=False).numpy()
    assert 0.2 < np.isnan(n).mean() < 0.9
    assert int(torch.from_numpy(a.view(np.int32).astype(np.int64)).sum()) == int(a.view(np.int32).astype(np.int64).sum())
