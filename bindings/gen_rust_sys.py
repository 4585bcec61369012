"""Generates bindings/rust/dcdf-cuda-sys/src/lib.rs from include/dcdf_cuda.h.

The reference is a Rust crate and this image has no rustc / bindgen, so the `-sys` crate a maintainer would add under
dcdf/ is produced by this small C-declaration reader instead (the header uses a regular subset of C: anonymous enums,
integer #defines, opaque and plain structs, one function-pointer typedef, prototypes).  tests/test_bindings_cpu.py checks
that the committed lib.rs is what this script emits, that it declares every symbol libdcdf_cuda.so exports, and that the
struct layouts agree with the ctypes declarations the GPU tests call through.

    python bindings/gen_rust_sys.py            # rewrites lib.rs
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dcdf_cuda.h")
OUT = os.path.join(ROOT, "bindings", "rust", "dcdf-cuda-sys", "src", "lib.rs")

SCALARS = {"int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8", "float": "f32", "double": "f64",
           "char": "c_char", "void": "c_void"}
SIZES = {"i32": 4, "u32": 4, "i64": 8, "u64": 8, "u8": 1, "f32": 4, "f64": 8}


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def rust_type(c, known):
    """`const uint8_t**` -> `*mut *const u8`; names of header structs / typedefs pass through."""
    c = c.strip()
    stars = c.count("*")
    c = c.replace("*", " ")
    words = c.split()
    const = "const" in words
    base = [w for w in words if w not in ("const", "struct")]
    assert len(base) == 1, c
    name = base[0]
    if name in SCALARS:
        t = SCALARS[name]
    elif name in known:
        t = name
    else:
        raise ValueError(f"unknown C type {name!r}")
    if stars == 0:
        if t == "c_void":
            return "()"
        return t
    # only the innermost pointee's constness is spelled in the header (`const T*`, `const T**`, `T**`)
    out = ("*const " if const else "*mut ") + t
    for _ in range(stars - 1):
        out = "*mut " + out
    return out


def parse(text):
    text = strip_comments(text)
    consts, opaque, structs, fn_types, protos = [], [], [], [], []
    for m in re.finditer(r"#define\s+(DCDF_[A-Z0-9_]+)\s+(-?\d+)\s*$", text, flags=re.M):
        consts.append((m.group(1), int(m.group(2))))
    for m in re.finditer(r"enum\s*\{(.*?)\}\s*;", text, flags=re.S):
        value = -1
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = (x.strip() for x in item.split("="))
                value = int(v, 0)
            else:
                name, value = item, value + 1
            consts.append((name, value))
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s+(\w+)\s*;", text):
        opaque.append(m.group(2))
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = [d.strip() for d in decl.split(",")]
            mm = re.match(r"(.*?)(\w+)\s*(\[\d+\])?$", first)
            ctype = mm.group(1)
            for piece in [first[len(ctype):]] + rest:
                pm = re.match(r"(\w+)\s*(?:\[(\d+)\])?$", piece.strip())
                fields.append((pm.group(1), ctype.strip(), int(pm.group(2)) if pm.group(2) else None))
        structs.append((m.group(3), fields))
    for m in re.finditer(r"typedef\s+(\w+)\s*\(\s*\*\s*(\w+)\s*\)\s*\((.*?)\)\s*;", text, flags=re.S):
        fn_types.append((m.group(2), m.group(1), m.group(3)))
    body = re.sub(r"typedef[^;{]*\{.*?\}[^;]*;", " ", text, flags=re.S)
    body = re.sub(r"typedef[^;]*;", " ", body)
    for m in re.finditer(r"((?:const\s+)?\w+\s*\*?)\s*(dcdf_\w+)\s*\(([^;{]*?)\)\s*;", body, flags=re.S):
        protos.append((m.group(2), m.group(1), m.group(3)))
    return consts, opaque, structs, fn_types, protos


def params(text, known):
    text = " ".join(text.split())
    if text in ("", "void"):
        return []
    out = []
    for p in text.split(","):
        p = p.strip()
        m = re.match(r"(.*?)(\w+)\s*(\[\d+\])?$", p)
        ctype, name, arr = m.group(1), m.group(2), m.group(3)
        t = rust_type(ctype, known)
        if arr:                                     # `int64_t shape[3]` in a parameter list is a pointer
            t = "*mut " + t
        out.append((name if name not in ("in", "type", "ref", "box", "use") else name + "_", t))
    return out


def struct_layout(fields, layouts):
    """(size, align) under the C / #[repr(C)] rules."""
    off, align = 0, 1
    for _, ctype, n in fields:
        t = rust_type(ctype, layouts)
        if t.startswith("*"):
            size, a = 8, 8
        elif t in SIZES:
            size = a = SIZES[t]
        else:
            size, a = layouts[t]
        off = (off + a - 1) // a * a + size * (n or 1)
        align = max(align, a)
    return (off + align - 1) // align * align, align


def generate():
    consts, opaque, structs, fn_types, protos = parse(open(HEADER).read())
    known = set(opaque) | {s for s, _ in structs} | {f for f, _, _ in fn_types}
    L = ["// dcdf-cuda-sys: raw FFI of libdcdf_cuda.so (include/dcdf_cuda.h).  GENERATED by bindings/gen_rust_sys.py -- do not edit.",
         "// Every item mirrors the header declaration of the same name; the header cites the dcdf source line each entry replaces.",
         "#![allow(non_camel_case_types, non_upper_case_globals)]",
         "use std::os::raw::{c_char, c_void};", ""]
    for name, value in consts:
        L.append(f"pub const {name}: i32 = {value};")
    L.append("")
    for name in opaque:
        L += ["#[repr(C)]", f"pub struct {name} {{ _private: [u8; 0] }}"]
    L.append("")
    for name, fields in structs:
        L += ["#[repr(C)]", "#[derive(Clone, Copy, Debug)]", f"pub struct {name} {{"]
        for fname, ctype, n in fields:
            t = rust_type(ctype, known)
            L.append(f"    pub {fname}: {f'[{t}; {n}]' if n else t},")
        L += ["}", ""]
    for name, ret, ptext in fn_types:
        ps = ", ".join(f"{n}: {t}" for n, t in params(ptext, known))
        L += [f"pub type {name} = Option<unsafe extern \"C\" fn({ps}) -> {rust_type(ret, known)}>;", ""]
    L += ['#[link(name = "dcdf_cuda")]', 'extern "C" {']
    for name, ret, ptext in protos:
        ps = ", ".join(f"{n}: {t}" for n, t in params(ptext, known))
        r = rust_type(ret, known)
        L.append(f"    pub fn {name}({ps}){'' if r == '()' else ' -> ' + r};")
    L += ["}", ""]
    return "\n".join(L), (consts, opaque, structs, fn_types, protos)


if __name__ == "__main__":
    src, _ = generate()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write(src)
    print("wrote", OUT)
