"""B1 drop-in proof: the Backend::Cuda adapter - miplib_b200/miplib/cuda/*.{hpp,cpp},
byte for byte as shipped - compiles against the REFERENCE's own headers once the edits
INTEGRATION.md section 2 lists are applied to them.

The test works on a temporary copy (nothing of the reference enters the repository): it copies
/root/reference/miplib/*.hpp, appends `Cuda` to `Solver::Backend` (miplib/solver.hpp:20) and adds the
friend declarations (miplib/solver.hpp:109-120, miplib/constr.hpp:52-56), supplies a 12-line stand-in
for <boost/container_hash/hash.hpp> (Boost is not installed here; the reference only needs
boost::hash of a pair, miplib/util.hpp:78), and compiles every adapter source to an object file
with g++ -std=c++17.  /root/reference exists only in the build container, so the test is skipped
on the GPU box.
"""
from __future__ import annotations

import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/miplib")
ADAPTER = ROOT / "miplib_b200" / "miplib" / "cuda"

BOOST_SHIM = """#pragma once
#include <cstddef>
#include <functional>
#include <utility>
namespace boost {
template<class T> struct hash;
template<class T> inline void hash_combine(std::size_t& seed, T const& v);
template<class A, class B> struct hash<std::pair<A, B>> {
  std::size_t operator()(std::pair<A, B> const& p) const noexcept
  { std::size_t s = 0; hash_combine(s, p.first); hash_combine(s, p.second); return s; }
};
template<class T> inline void hash_combine(std::size_t& seed, T const& v)
{ seed ^= hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2); }
}
"""


def _patch(path: Path, old: str, new: str):
    s = path.read_text()
    assert s.count(old) >= 1, f"{path.name}: the reference no longer contains {old!r}"
    path.write_text(s.replace(old, new, 1))


@pytest.fixture(scope="module")
def ref_tree(tmp_path_factory):
    if not REF.is_dir():
        pytest.skip("/root/reference is not present on this machine")
    top = tmp_path_factory.mktemp("dropin")
    inc = top / "miplib"
    inc.mkdir()
    for h in REF.glob("*.hpp"):
        shutil.copy(h, inc / h.name)
    shutil.copytree(REF / "util", inc / "util")
    (top / "boost" / "container_hash").mkdir(parents=True)
    (top / "boost" / "container_hash" / "hash.hpp").write_text(BOOST_SHIM)
    # --- the edits of INTEGRATION.md section 2 -------------------------------------------------
    _patch(inc / "solver.hpp", "BestAtCompileTime, BestAtRunTime }", "BestAtCompileTime, BestAtRunTime, Cuda }")
    _patch(inc / "solver.hpp", "  friend struct LpsolveVar;\n",
           "  friend struct LpsolveVar;\n  friend struct CudaSolver;\n  friend struct CudaVar;\n")
    _patch(inc / "constr.hpp", "  friend struct LpsolveSolver;\n",
           "  friend struct LpsolveSolver;\n  friend struct CudaSolver;\n  friend struct CudaBranchAndBound;\n")
    # --- the adapter, unchanged ---------------------------------------------------------------
    (inc / "cuda").mkdir()
    for f in ADAPTER.iterdir():
        if f.suffix in (".hpp", ".cpp"):
            shutil.copy(f, inc / "cuda" / f.name)
    return top


def _compile(top: Path, src: Path, obj: Path):
    # the reference's headers lean on transitive includes of its vendor headers (assert, std::abs,
    # std::string): the three -include flags stand in for those, nothing else is added
    cmd = ["g++", "-std=c++17", "-O0", "-c", "-DWITH_CUDA", "-include", "cmath", "-include", "cassert",
           "-include", "string", f"-I{top}", f"-I{ROOT / 'include'}", str(src), "-o", str(obj)]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.mark.parametrize("unit", ["solver.cpp", "var.cpp", "bnb.cpp", "mps.cpp"])
def test_adapter_compiles_against_the_reference_headers(ref_tree, unit):
    src = ref_tree / "miplib" / "cuda" / unit
    p = _compile(ref_tree, src, ref_tree / (unit + ".o"))
    assert p.returncode == 0, p.stderr[-4000:]
    # the object really contains the ISolver implementation (not an empty translation unit)
    if unit == "solver.cpp":
        nm = subprocess.run(["nm", "-C", str(ref_tree / (unit + ".o"))], capture_output=True, text=True).stdout
        for sym in ("miplib::CudaSolver::solve()", "miplib::CudaSolver::add(miplib::Constr const&)",
                    "miplib::CudaSolver::create_var(", "miplib::CudaSolver::set_objective("):
            assert sym in nm, sym


def test_adapter_overrides_every_pure_virtual_of_the_reference_isolver(ref_tree):
    """`new CudaSolver(true)` against the reference's detail::ISolver (miplib/solver.hpp:125-188):
    compiles only if no pure virtual is left unimplemented and every `override` matches."""
    probe = ref_tree / "probe.cpp"
    probe.write_text(
        '#include <miplib/cuda/solver.hpp>\n#include <miplib/cuda/var.hpp>\n'
        'std::shared_ptr<miplib::detail::ISolver> make() { return std::make_shared<miplib::CudaSolver>(true); }\n'
        'miplib::Solver::Backend cuda_enum() { return miplib::Solver::Backend::Cuda; }\n')
    p = _compile(ref_tree, probe, ref_tree / "probe.o")
    assert p.returncode == 0, p.stderr[-4000:]


def test_mirror_headers_keep_the_reference_visibility():
    """The repo's own front-end mirror must not be laxer than the reference at the seam: Constr::p_impl
    and Solver::p_impl private, LazyConstrHandler without extra members."""
    constr = (ROOT / "miplib_b200" / "miplib" / "constr.hpp").read_text()
    body = constr[constr.index("struct Constr"):constr.index("std::ostream& operator<<(std::ostream& os, Constr const& c);\n\n")]
    assert re.search(r"private:[^}]*std::shared_ptr<detail::IConstr> p_impl;", body, re.S)
    lazy = (ROOT / "miplib_b200" / "miplib" / "lazy.hpp").read_text()
    assert "valid()" not in lazy
    adapter = "".join(f.read_text() for f in ADAPTER.glob("*.cpp"))
    assert ".valid()" not in adapter
