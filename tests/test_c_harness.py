"""CPU tests of the drop-in boundary from the two languages that bind it without ctypes:

* a plain-C caller (tests/c_harness/abi_harness.c) compiled as C99 -pedantic -Werror against include/movfem_b200.h and
  linked with the shared library: the struct layouts it reports are compared with the ctypes mirror, and without a
  CUDA device the create call must fail loudly with MOVFEM_E_NOGPU;
* the Fortran ISO_C_BINDING shim (movfem_b200/fortran/movfem_cuda.f90, not compilable here: no Fortran compiler in
  the image) is checked textually against the header: every bind(C) type has the header's fields in the header's
  order with interoperable kinds, every interface binds an exported symbol with the prototype's argument count, and
  by-value / by-reference passing matches the C declaration.
"""
import ctypes as C
import os
import re
import subprocess

import pytest

from movfem_b200 import abi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "movfem_b200.h")
SHIM = os.path.join(ROOT, "movfem_b200", "fortran", "movfem_cuda.f90")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = host.build()
    exe = str(tmp_path_factory.mktemp("c_harness") / "abi_harness")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_harness", "abi_harness.c"), "-o", exe,
           "-L", os.path.dirname(so), "-l:" + os.path.basename(so), "-Wl,-rpath," + os.path.dirname(so)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def _run(exe, *args):
    out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
    return out.returncode, out.stdout


def test_header_is_c99_and_layouts_match_the_ctypes_mirror(harness):
    rc, txt = _run(harness, "layout")
    assert rc == 0, txt
    sizes = {m.group(1): int(m.group(2)) for m in re.finditer(r"^sizeof (\w+) (\d+)$", txt, flags=re.M)}
    offs = {(m.group(1), m.group(2)): int(m.group(3)) for m in re.finditer(r"^offset (\w+)\.(\w+) (\d+)$", txt, flags=re.M)}
    mirror = {"movfem_desc": abi.MovfemDesc, "movfem_stats": abi.MovfemStats, "movfem_geomodel": abi.MovfemGeomodel}
    assert set(sizes) == set(mirror)
    for name, cls in mirror.items():
        assert sizes[name] == C.sizeof(cls), name
    assert len(offs) >= 20
    for (name, fld), off in offs.items():
        assert getattr(mirror[name], fld).offset == off, (name, fld)
    assert "sm_100a" in txt


def test_c_caller_fails_loudly_without_a_gpu(harness):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the C caller's device pass belongs to the gpu suite")
    rc, txt = _run(harness)
    assert rc == 0, txt
    assert f"create {abi.MOVFEM_E_NOGPU}" in txt and "assemble" not in txt


# ---- Fortran shim vs header -------------------------------------------------------------------------------------

_F_KIND = {"integer(c_int32_t)": "int32_t", "integer(c_int64_t)": "int64_t", "real(c_double)": "double", "type(c_ptr)": "ptr",
           "integer(c_int)": "int"}


def _c_struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), open(HEADER).read(), flags=re.S).group(1)
    out = []
    for decl in _strip_c_comments(body).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const double \*|int32_t|int64_t|double)\s*(.*)", decl)
        base, rest = m.group(1), m.group(2)
        for item in rest.split(","):
            item = item.strip()
            is_ptr = base.endswith("*") or item.startswith("*")
            dims = [int(x) for x in re.findall(r"\[(\d+)\]", item)]
            nm = re.sub(r"\[\d+\]", "", item).lstrip("*").strip()
            out.append((nm, "ptr" if is_ptr else base, dims))
    return out


def _fortran_type_fields(text, name):
    body = re.search(r"type, bind\(C\) :: %s\b(.*?)end type" % name, text, flags=re.S | re.I).group(1)
    out = []
    for line in body.splitlines():
        line = line.split("!")[0].strip()
        if not line:
            continue
        kind, names = [x.strip() for x in line.split("::")]
        for item in re.findall(r"\w+(?:\([\d,]+\))?", names):
            m = re.match(r"(\w+)(?:\(([\d,]+)\))?", item)
            dims = [int(x) for x in m.group(2).split(",")] if m.group(2) else []
            out.append((m.group(1), _F_KIND[kind.replace(" ", "")], dims))
    return out


def test_fortran_bind_c_types_mirror_the_header_field_for_field():
    text = open(SHIM).read()
    c_desc = _c_struct_fields("movfem_desc")
    for fname in ("movfem_desc", "movfem_desc_geo"):
        assert _fortran_type_fields(text, fname) == c_desc, fname
    c_gm, f_gm = _c_struct_fields("movfem_geomodel"), _fortran_type_fields(text, "movfem_geomodel")
    assert [(n, k) for n, k, _ in f_gm] == [(n, k) for n, k, _ in c_gm]
    # Fortran is column-major: C int32_t ijsigma[9][2] is ijsigma(2,9) on the Fortran side
    assert [list(reversed(d)) for _, _, d in f_gm] == [d for _, _, d in c_gm]


def _c_prototypes():
    txt = _strip_c_comments(open(HEADER).read())
    txt = txt[txt.index("typedef struct movfem_handle"):]
    protos = {}
    for m in re.finditer(r"\b(?:int|void|const char \*)\s*(movfem_[a-z0-9_]+)\s*\((.*?)\)\s*;", txt, flags=re.S):
        args = [" ".join(a.split()) for a in m.group(2).split(",")]
        protos[m.group(1)] = [] if args == ["void"] else args
    return protos


def test_fortran_interfaces_bind_exported_symbols_with_matching_arguments():
    protos = _c_prototypes()
    assert {"movfem_create", "movfem_assemble", "movfem_sizes", "movfem_get_gne", "movfem_destroy"} <= set(protos)
    text = open(SHIM).read()
    text = re.sub(r"&\s*\n\s*", " ", text)       # join continuation lines
    lib = C.CDLL(host.build())
    bound = []
    for m in re.finditer(r"^\s*(?:integer\(c_int\) function|subroutine|function)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C, name='(\w+)'\)(.*?)"
                         r"end (?:function|subroutine)", text, flags=re.S | re.M | re.I):
        fname, fargs, cname, body = m.group(1), [a.strip() for a in m.group(2).split(",")], m.group(3), m.group(4)
        assert fname == cname and cname in protos, cname
        assert hasattr(lib, cname), cname
        cargs = protos[cname]
        assert len(fargs) == len(cargs), (cname, fargs, cargs)
        # every dummy argument is declared; scalars the C side takes by value carry VALUE, pointers do not
        decls = {}
        for line in body.splitlines():
            line = line.split("!")[0]
            if "::" not in line or line.strip().startswith("import"):
                continue
            attrs, names = line.split("::")
            for nm in re.findall(r"\b(\w+)(?:\([^)]*\))?", names):
                decls[nm] = attrs.lower()
        for fa, ca in zip(fargs, cargs):
            assert fa in decls, (cname, fa)
            by_value_c = "*" not in ca
            if "movfem_handle **" in ca:
                assert "type(c_ptr)" in decls[fa] and "value" not in decls[fa], (cname, fa)  # receives the handle
            elif ca.replace("const ", "").startswith("movfem_handle *") or ca.startswith("void *"):
                assert "type(c_ptr)" in decls[fa] and "value" in decls[fa], (cname, fa)      # opaque handle passed by value
            elif fa == "ms_device":
                assert "type(c_ptr)" in decls[fa] and "value" in decls[fa]                     # optional double*: c_null_ptr
            else:
                assert ("value" in decls[fa]) == by_value_c, (cname, fa, ca, decls[fa])
        bound.append(cname)
    assert {"movfem_create", "movfem_destroy", "movfem_sizes", "movfem_get_gne", "movfem_assemble", "movfem_last_error",
            "movfem_geo_innermodel"} <= set(bound)
