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


def test_julia_shim_implements_the_abstract_cache_interface():
    # The reference's extension point is dispatch on AbstractBeliefPropagationCache{V, PV}
    # (src/caches/abstractbeliefpropagationcache.jl:17,44-69) plus the algorithm tag that initialize_cache / update
    # dispatch on (src/initialize_cache.jl:10-29, abstract :313-337).  Julia is not available here, so this is a static
    # check that julia/ITNB200.jl defines a method on B200BeliefPropagationCache for every function a new cache type must
    # provide, and that the front-ends' entry points carry the "bp_b200" tag.
    jl = open(os.path.join(ROOT, "julia", "ITNB200.jl")).read()
    m = re.search(r"mutable struct B200BeliefPropagationCache\{([^}]*)\}\s*<:\s*AbstractBeliefPropagationCache\{V,\s*PV\}", jl, flags=re.S)
    assert m, "the cache must subtype AbstractBeliefPropagationCache{V, PV} with BOTH parameters"
    assert m.group(1).replace(" ", "").startswith("V,PV")
    required = ["partitioned_tensornetwork", "messages", "default_update_alg", "default_message_update_alg", "default_bp_maxiter",
                "default_edge_sequence", "default_bp_edge_sequence", "environment", "region_scalar", "partitions", "rescale",
                "rescale_messages", "rescale_partitions", "update_factors", "update_factor", "message", "set_message!",
                "set_messages!", "updated_message", "logscalar", "vertex_scalars", "edge_scalars", "set_default_kwargs", "update",
                "tensornetwork"]
    for name in required:
        pat = r"ITensorNetworks\." + re.escape(name) + r"\((?:[^()]|\([^()]*\))*B200BeliefPropagationCache"
        assert re.search(pat, jl), f"no B200BeliefPropagationCache method for ITensorNetworks.{name}"
    for pat in (r"Base\.copy\(bpc::B200BeliefPropagationCache", r"PartitionedGraphs\.quotientedges\(bpc::B200BeliefPropagationCache",
                r"PartitionedGraphs\.partitioned_vertices\(bpc::B200BeliefPropagationCache",
                r"DataGraphs\.set_vertex_data!\(bpc::B200BeliefPropagationCache",
                r"Adapt\.adapt_structure\(to::B200Device, bpc::BeliefPropagationCache\)",
                r"Adapt\.adapt_structure\(to::Type\{<:Array\}, bpc::B200BeliefPropagationCache\)"):
        assert re.search(pat, jl), pat
    # region_scalar for both region kinds
    assert re.search(r"region_scalar\(bpc::B200BeliefPropagationCache, pv::QuotientVertex", jl)
    assert re.search(r"region_scalar\(bpc::B200BeliefPropagationCache, pe::QuotientEdge", jl)
    # the algorithm tag the front-ends dispatch on
    for fn in ("initialize_cache", "update", "set_default_kwargs", "expect"):
        assert re.search(r"ITensorNetworks\." + fn + r"\(alg::Algorithm\"bp_b200\"", jl), fn
    assert re.search(r"ITensors\.apply\(o::ITensor, ψ::AbstractITensorNetwork; envs::B200Environment", jl)
    # every name imported from ITensorNetworks exists in the reference sources (a typo would fail at `using` time)
    ref = os.environ.get("ITN_REFERENCE_SRC", "/root/reference/src")
    if os.path.isdir(ref):
        src = "\n".join(open(os.path.join(dp, f), errors="ignore").read() for dp, _, fs in os.walk(ref) for f in fs if f.endswith(".jl"))
        imp = re.search(r"using ITensorNetworks: (.*?)\nusing", jl, flags=re.S).group(1)
        for name in [x.strip() for x in imp.replace("\n", " ").split(",")]:
            if name and name != "ITensorNetworks":
                assert re.search(r"\b" + re.escape(name) + r"\b", src), f"{name} is not defined in the reference"
        for name in set(re.findall(r"ITensorNetworks\.([a-z_!]+)\(", jl)):
            assert re.search(r"(function\s+|\.|^|\s)" + re.escape(name) + r"\s*\(", src, flags=re.M), \
                f"ITensorNetworks.{name} is not a function of the reference"
