"""The Rust `-sys` crate a dcdf maintainer would add (bindings/rust/dcdf-cuda-sys) is generated from include/dcdf_cuda.h.
No Rust toolchain exists in this image, so it is checked from the other three sides: against the generator, against the
symbols the shared library really exports, and against the ctypes declarations every GPU test calls through."""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bindings"))
import gen_rust_sys as gen  # noqa: E402

from dcdf_b200 import _ffi  # noqa: E402


def test_committed_binding_is_what_the_generator_emits():
    src, _ = gen.generate()
    assert open(gen.OUT).read() == src, "run python bindings/gen_rust_sys.py"


def test_every_exported_symbol_is_declared_once():
    _ffi.build_library()
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T dcdf_" in ln}
    declared = re.findall(r"pub fn (dcdf_\w+)\(", open(gen.OUT).read())
    assert len(declared) == len(set(declared)) and set(declared) == exported


def test_struct_layouts_equal_the_ctypes_ones():
    _, (_, _, structs, _, _) = gen.generate()
    mirror = {"dcdf_array3": _ffi.Array3, "dcdf_build_stats": _ffi.BuildStats, "dcdf_cube": _ffi.Cube,
              "dcdf_superchunk_info": _ffi.SuperchunkInfo}
    layouts = {}
    for name, fields in structs:
        layouts[name] = gen.struct_layout(fields, layouts)
        ct = mirror[name]
        assert layouts[name] == (C.sizeof(ct), C.alignment(ct)), name
        assert [f for f, _, _ in fields] == [f[0] for f in ct._fields_], name
    assert set(layouts) == set(mirror)


def test_prototypes_agree_with_the_ctypes_signatures():
    """Same arity, and pointer / 32-bit / 64-bit / float class per argument, as dcdf_b200/_ffi.py declares them."""
    _, (_, opaque, structs, fn_types, protos) = gen.generate()
    known = set(opaque) | {s for s, _ in structs} | {f for f, _, _ in fn_types}
    lib = _ffi.lib()

    def klass(rust):
        if rust.startswith("*") or rust.startswith("Option<") or rust in known:
            return "ptr"
        return {"i32": "32", "u32": "32", "i64": "64", "u64": "64", "f32": "f32"}[rust]

    def cklass(ct):
        if ct in (C.c_int32, C.c_uint32):
            return "32"
        if ct in (C.c_int64, C.c_uint64):
            return "64"
        if ct is C.c_float:
            return "f32"
        return "ptr"                                        # c_void_p, c_char_p, POINTER(..), CFUNCTYPE

    assert len(protos) >= 48
    for name, ret, ptext in protos:
        fn = getattr(lib, name)
        ps = gen.params(ptext, known)
        assert fn.argtypes is not None and len(fn.argtypes) == len(ps), name
        assert [klass(t) for _, t in ps] == [cklass(a) for a in fn.argtypes], name
        r = gen.rust_type(ret, known)
        assert klass(r) == cklass(fn.restype), name
