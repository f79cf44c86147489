"""The Rust FFI crate (rust/film_grain_cuda) cannot be compiled in this image (no rustc), so its declarations are checked
against include/fg.h textually: every exported function is declared with the same number of parameters, the two
#[repr(C)] structs list the header's fields in the header's order with matching widths, and the error codes agree."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "fg.h")).read()
RS = open(os.path.join(ROOT, "rust", "film_grain_cuda", "src", "lib.rs")).read()


def _strip_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def _c_functions():
    src = _strip_comments(HDR)
    out = {}
    for m in re.finditer(r"\b(?:int|void|uint64_t|const char\*)\s+(fg_\w+)\s*\(([^)]*)\)\s*;", src):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def _rs_functions():
    out = {}
    for m in re.finditer(r"pub fn (fg_\w+)\s*\(([^)]*)\)", RS, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return out


def test_every_header_function_is_declared_in_rust_with_the_same_arity():
    c, r = _c_functions(), _rs_functions()
    assert len(c) >= 20
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name in c:
        assert c[name] == r[name], (name, c[name], r[name])


def _c_struct(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), _strip_comments(HDR), flags=re.S).group(1)
    fields = []
    for line in body.split(";"):
        line = line.strip()
        if not line:
            continue
        ty, names = line.rsplit(" ", 1)[0], line.split(" ", 1)[1]
        ty = line.split()[0]
        for n in line[len(ty):].split(","):
            fields.append((n.strip(), ty))
    return fields


def _rs_struct(name):
    body = re.search(r"pub struct %s \{(.*?)\n    \}" % name, RS, flags=re.S).group(1)
    return [(m.group(1), m.group(2)) for m in re.finditer(r"pub (\w+): (\w+),", body)]


def test_repr_c_structs_match_the_header_field_for_field():
    width = {"uint32_t": "u32", "uint64_t": "u64", "float": "f32", "double": "f64"}
    for c_name, rs_name in (("fg_params", "FgParams"), ("fg_stats", "FgStats")):
        c, r = _c_struct(c_name), _rs_struct(rs_name)
        assert [n for n, _ in c] == [n for n, _ in r], (c_name, c, r)
        assert [width[t] for _, t in c] == [t for _, t in r], c_name


def test_error_codes_agree():
    for m in re.finditer(r"(FG_(?:OK|ERR_\w+)) = (-?\d+)", HDR):
        rs = re.search(r"pub const %s: c_int = (-?\d+);" % m.group(1), RS)
        assert rs and rs.group(1) == m.group(2), m.group(1)
