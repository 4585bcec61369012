"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dcdf_cuda.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    hdr = open(os.path.join(ROOT, "include", "dcdf_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dcdf_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    from dcdf_b200 import _ffi
    path = _ffi.build_library()
    assert os.path.exists(path)
    out = subprocess.check_output(["nm", "-D", "--defined-only", path]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = _header_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/dcdf_cuda.h but not exported: {missing}"
    lib = _ffi.lib()                      # ctypes resolves every declared symbol with its signature
    assert set(_ffi.DECLARED) == set(declared)
    assert lib.dcdf_abi_version() == 1


def test_library_is_sm100a_and_has_no_oracle_dependency():
    from dcdf_b200 import _ffi
    path = _ffi.build_library()
    elf = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True)
    if elf.returncode == 0:
        assert "sm_100a" in elf.stdout
    needed = subprocess.check_output(["objdump", "-p", path]).decode()
    assert "liboracle" not in needed and "torch" not in needed
    # the product sources never reference the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dcdf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "synth.py", f"{f} mentions the oracle"


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_device():
    from dcdf_b200 import Context, DcdfError
    with pytest.raises(DcdfError) as e:
        Context(0)
    assert e.value.code == 7          # DCDF_ERR_CUDA


def test_null_and_bad_arguments_do_not_crash():
    from dcdf_b200 import _ffi
    lib = _ffi.lib()
    assert lib.dcdf_ctx_create(0, None) == 8
    assert lib.dcdf_ctx_destroy(None) == 0
    assert lib.dcdf_chunk_free(None) == 0
    assert lib.dcdf_superchunk_free(None) == 0
    assert lib.dcdf_chunk_size(None, None) == 8
    assert lib.dcdf_ctx_launch_count(None) == 0
    assert lib.dcdf_last_error(None) == b"null context"
    a = _ffi.Array3()
    kind, bits = C.c_int32(), C.c_int32()
    assert lib.dcdf_suggest_fraction(None, C.byref(a), C.byref(kind), C.byref(bits)) == 8


def test_array_descriptor_matches_ndarray_views():
    from dcdf_b200 import api
    a = np.zeros((5, 12, 20), np.float32)[1:4, 2:9, ::2]
    d, _ = api._describe(a)
    assert tuple(d.shape) == (3, 7, 10) and tuple(d.strides) == (240, 20, 2)
    assert d.encoding == 32 and d.mem == 0
    assert d.base == a.ctypes.data
    with pytest.raises(TypeError):
        api._describe(np.zeros((1, 2, 2), np.int16))
    with pytest.raises(ValueError):
        api._describe(np.zeros((2, 2), np.float32))
