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


# ---- module dependency graph: shim + reference must compile in SOME order ------------------------------------------------

def _module_uses(text):
    """{module: set(used modules)} of a Fortran source (intrinsic modules ignored)."""
    code = "\n".join(line.split("!")[0] for line in text.splitlines())
    out, cur = {}, None
    for line in code.splitlines():
        m = re.match(r"^\s*(module|program)\s+(\w+)\s*$", line, flags=re.I)
        if m and m.group(2).lower() != "procedure":
            cur = m.group(2).lower()
            out[cur] = set()
            continue
        m = re.match(r"^\s*use\s*(,\s*intrinsic\s*::)?\s*(\w+)", line, flags=re.I)
        if m and cur and not m.group(1):
            out[cur].add(m.group(2).lower())
    return out


def test_module_use_graph_of_shim_and_reference_is_acyclic():
    """The Fortran shim is called from INSIDE module global_assembly (ga_init) and module geometry (innermodel_gqg), so it
    must not use either -- a circular `use` does not compile.  Sorts the `use` graph of the reference's sources (fixture
    tests/golden/ref_module_uses.json, written from /root/reference by tests/golden/make_use_graph.py and re-derived live when
    the reference is present) + the shim's modules + the edits of INTEGRATION.md topologically."""
    import json
    fix = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_module_uses.json")))
    ref_src = "/root/reference/MoVFEM_3DMT/src"
    if os.path.isdir(ref_src):       # this container only; the fixture travels
        for fname, rec in fix.items():
            live = _module_uses(open(os.path.join(ref_src, fname), errors="replace").read())
            assert sorted(live) == sorted(rec["defines"]), fname
            assert sorted(set().union(*live.values())) == rec["uses"], fname
    graph = {}
    for rec in fix.values():
        for mod in rec["defines"]:
            graph.setdefault(mod, set()).update(u for u in rec["uses"] if u != mod)
    shim = _module_uses(open(SHIM).read())
    assert set(shim) == {"movfem_cuda", "movfem_cuda_geo"}
    assert "global_assembly" not in shim["movfem_cuda"] and "geometry" not in shim["movfem_cuda_geo"]
    graph.update({k: set(v) for k, v in shim.items()})
    # the edits of INTEGRATION.md section 2
    graph["global_assembly"].add("movfem_cuda")      # ga_init calls movfem_cuda_init
    graph["geometry"].add("movfem_cuda_geo")         # grid_3d calls movfem_cuda_innermodel (optional)
    graph["movfem_3dmt"].add("movfem_cuda")          # the frequency loop calls movfem_cuda_assemble
    for deps in graph.values():
        assert deps <= set(graph), deps - set(graph)
    order, done = [], set()
    while len(done) < len(graph):
        ready = sorted(m for m in graph if m not in done and graph[m] <= done)
        assert ready, "circular module dependency among " + ", ".join(sorted(set(graph) - done))
        order += ready
        done |= set(ready)
    assert order.index("movfem_cuda") < order.index("global_assembly") and order.index("movfem_cuda_geo") < order.index("geometry")
    # movfem_cuda_init receives global_assembly's variables as arguments
    assert re.search(r"subroutine movfem_cuda_init\(sym, nne, nnze, gne\)", open(SHIM).read())


@pytest.mark.gpu
def test_c_caller_runs_the_device_pass(harness):
    """The plain-C caller with malloc()ed (pageable) arrays on a real device: create -> sizes -> get_gne -> assemble -> destroy,
    the order the Fortran wrappers use; its own checks (sorted upper triangle, finite values, nz <= capacity) must pass."""
    rc, txt = _run(harness)
    assert rc == 0, txt
    assert "create 0" in txt and "assemble 0" in txt and "OK" in txt, txt
