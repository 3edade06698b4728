"""CPU-side checks: the C-ABI library loads and exports every symbol include/itn_b200.h declares,
and fails loudly without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

import itn_b200
from itn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "itn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(itn_[a-z0-9_]+)\s*\(", txt)))


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported():
    l = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/itn_b200.h but not exported"


def test_binding_covers_header():
    assert sorted(_lib.EXPORTED_SYMBOLS) == _header_symbols()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(itn_b200.ITNError) as ei:
        itn_b200.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "itensornetworks.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                assert "oracle" not in open(os.path.join(dp, f), errors="ignore").read().replace("second opinion", ""), f


def test_host_schedules():
    g = itn_b200.named_grid((4, 4))
    seq = itn_b200.default_edge_sequence(g)
    assert len(seq) == 2 * g.ne == len(set(seq))
    par = itn_b200.parallel_edge_sequence(g)
    assert len(par) == 2 * g.ne and all(len(p) == 1 for p in par)
    hh = itn_b200.heavy_hex_eagle()
    assert hh.nv == 127 and hh.ne == 144
    cols = itn_b200.edge_coloring(hh)
    assert 3 <= len(cols) <= 4 and sum(len(c) for c in cols) == 144
    for c in cols:
        vs = [x for e in c for x in hh.edges[e]]
        assert len(vs) == len(set(vs))
    g3 = itn_b200.named_grid((3, 3, 3))
    assert g3.ne == 54 and max(g3.degree(v) for v in range(g3.nv)) == 6


def test_binding_arity_matches_the_prototypes():
    # every ctypes signature of the host mirror has as many arguments as the C prototype it binds, and pointer / scalar
    # kinds agree position by position (a drifted signature would corrupt the stack silently)
    txt = open(os.path.join(ROOT, "include", "itn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = dict(re.findall(r"\b(itn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt))
    assert set(protos) == set(_lib.EXPORTED_SYMBOLS)
    for name, (_, argtypes) in _lib._SIGS.items():
        params = [p.strip() for p in protos[name].split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, binding {len(argtypes)}"
        for p, a in zip(params, argtypes):
            is_ptr_c = "*" in p or "[" in p
            is_ptr_py = a in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(a, "contents") or a.__name__.startswith("LP_")
            assert is_ptr_c == is_ptr_py, f"{name}: parameter `{p}` vs binding {a}"


def test_julia_shim_ccalls_match_the_prototypes():
    # julia/ITNB200.jl cannot be executed here (no Julia runtime); at least every ccall in it names an exported symbol and
    # passes as many argument types as the C prototype has parameters, pointers where the prototype has pointers
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "itn_b200.h")).read(), flags=re.S)
    protos = dict(re.findall(r"\b(itn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr))
    jl = open(os.path.join(ROOT, "julia", "ITNB200.jl")).read()
    calls = re.findall(r"ccall\(\(:(itn_[a-z0-9_]+),\s*LIB\),\s*(\w+),\s*\(([^)]*)\)", jl, flags=re.S)
    assert len(calls) >= 15
    for name, ret, args in calls:
        assert name in protos, f"{name} is not declared in include/itn_b200.h"
        params = [p.strip() for p in protos[name].split(",") if p.strip() and p.strip() != "void"]
        types = [t.strip() for t in args.split(",") if t.strip()]
        assert len(types) == len(params), f"{name}: ccall passes {len(types)} arguments, the prototype takes {len(params)}"
        assert ret == ("Cstring" if name == "itn_last_error" else "Cint")
        for p, t in zip(params, types):
            assert ("*" in p or "[" in p) == t.startswith("Ptr"), f"{name}: `{p}` vs {t}"
