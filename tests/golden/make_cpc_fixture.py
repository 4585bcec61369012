"""Converts the reference's real-world test raster (py-dcdf/tests/testdata.txt: one 360x720 float32
CPC-precip day, read by py-dcdf/tests/test_dcdf.py:344-355) into tests/golden/cpc_day_360x720.npz.
Run in the build container only (the GPU box has no /root/reference)."""
import numpy as np
vals = np.loadtxt("/root/reference/py-dcdf/tests/testdata.txt", dtype=np.float32)
np.savez_compressed("tests/golden/cpc_day_360x720.npz", day=vals.reshape(360, 720))
print(vals.shape, np.isnan(vals).mean())
