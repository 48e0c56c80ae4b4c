"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/spl.h declares; the host mirror fails loudly without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import spalinalg_b200 as sp
from spalinalg_b200 import _capi, build as spl_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    spl_build.build()
    return _capi.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spl.h but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_capi.SIGNATURES) == names


def test_no_oracle_in_product_path():
    """The product path must not import or call the oracle (test infrastructure only)."""
    pkg = os.path.join(ROOT, "spalinalg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "orc_" not in src, f


def test_fails_loudly_without_cuda(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    h = ctypes.c_void_p()
    assert lib.spl_ctx_create(0, None, ctypes.byref(h)) == _capi.SPL_ERR_CUDA
    with pytest.raises(sp.DeviceError):
        sp.CsrMatrix.eye(2)
    with pytest.raises(sp.DeviceError):
        sp.CsrMatrix.from_coo(sp.CooMatrix.with_entries(2, 2, [(0, 0, 1.0)]))
    with pytest.raises(sp.DeviceError):                       # the streamed CooMatrix storage has no host-only form
        sp.PinnedCooMatrix.new(2, 2)
    assert lib.spl_coo_free(None) == _capi.SPL_ERR_ARG and lib.spl_coo_len(None) == 0
    with pytest.raises(sp.Panic):                             # the reference's assert comes before any device work
        sp.PinnedCooMatrix.new(0, 2)


def test_coo_shell_mirrors_reference_panics():
    # src/coo.rs tests: new_invalid_nrows / push_invalid_row / with_triplets_* (should_panic)
    with pytest.raises(sp.Panic):
        sp.CooMatrix.new(0, 1)
    with pytest.raises(sp.Panic):
        sp.CooMatrix.new(1, 0)
    coo = sp.CooMatrix.new(1, 2)
    with pytest.raises(sp.Panic):
        coo.push(1, 0, 1.0)
    with pytest.raises(sp.Panic):
        coo.push(0, 2, 1.0)
    with pytest.raises(sp.Panic):
        sp.CooMatrix.with_triplets(2, 2, [0, 1], [0], [1.0, 2.0])
    with pytest.raises(sp.Panic):
        sp.CooMatrix.with_entries(2, 2, [(2, 0, 1.0)])
    coo = sp.CooMatrix.with_entries(2, 3, [(0, 1, 1.0), (1, 2, 2.0)])
    assert coo.length() == 2 and coo.shape() == (2, 3) and coo.get(1) == (1, 2, 2.0) and coo.get(2) is None
    assert coo.pop() == (1, 2, 2.0) and coo.length() == 1
    t = coo.transpose()
    assert t.shape() == (3, 2) and t.get(0) == (1, 0, 1.0)
    s = (coo + coo)
    assert s.length() == 2
    assert (-coo).get(0) == (0, 1, -1.0)
    assert list((coo - coo).iter()) == [(0, 1, 1.0), (0, 1, -1.0)]
    with pytest.raises(sp.Panic):
        coo + sp.CooMatrix.new(3, 3)


def test_dok_shell():
    dok = sp.DokMatrix.new(2, 2)
    assert dok.insert(0, 0, 1.0) is None
    assert dok.insert(0, 0, 2.0) == 1.0
    assert dok.get(0, 0) == 2.0 and dok.contains(0, 0) and not dok.contains(1, 1) and dok.length() == 1
    with pytest.raises(sp.Panic):
        dok.insert(2, 0, 1.0)


def test_dok_arithmetic_and_from_coo():
    """src/dok.rs:1081-1111 (add, sub, neg goldens) and From<&CooMatrix> (src/dok.rs:640-668)."""
    lhs = sp.DokMatrix.with_entries(1, 1, [(0, 0, 1.0)])
    rhs = sp.DokMatrix.with_entries(1, 1, [(0, 0, 2.0)])
    assert list((lhs + rhs).iter()) == [(0, 0, 3.0)]
    assert list((lhs - rhs).iter()) == [(0, 0, -1.0)]
    assert list((-lhs).iter()) == [(0, 0, -1.0)]
    only_rhs = sp.DokMatrix.with_entries(2, 2, [(1, 1, 5.0)])
    assert (sp.DokMatrix.new(2, 2) - only_rhs).get(1, 1) == -5.0           # 0.0 - v
    coo = sp.CooMatrix.with_entries(2, 3, [(0, 0, 1.0), (1, 1, 2.0), (0, 0, 0.5), (1, 2, -0.0)])
    dok = sp.DokMatrix.from_coo(coo)
    assert dok.length() == 3 and dok.get(0, 0) == 1.5 and dok.get(1, 1) == 2.0
    import numpy as np
    assert not np.signbit(dok.get(1, 2))                                    # 0.0 + -0.0 = +0.0, entry kept


def test_synthetic_generators():
    from spalinalg_b200 import synthetic as syn
    r, c, v = syn.laplacian_2d(8)
    assert len(v) == 5 * 64 - 4 * 8 and v.sum() == 4 * 8       # row sums: boundary excess
    r, c, v = syn.stencil_27(5)
    assert len(v) == (3 * 5 - 2) ** 3
    r, c, v = syn.banded(100, range(-4, 5))
    assert len(v) == 9 * 100 - 20


def _c_declarations():
    import re
    h = open(os.path.join(ROOT, "include", "spl.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    h = re.sub(r"//[^\n]*", "", h)
    out = {}
    for name, args in re.findall(r"\b(?:int|uint64_t|const char \*)\s*(spl_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        out[name] = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
    return out


def test_rust_and_python_bindings_follow_the_header():
    """rust/src/ffi.rs (not compiled in this image) and spalinalg_b200/_capi.py declare every entry point of
    include/spl.h with the same number of arguments, pointers where the header has pointers and integers of
    the header's width where it has integers."""
    import re
    from spalinalg_b200 import _capi
    c = _c_declarations()
    assert len(c) >= 60
    r = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    r = re.sub(r"//[^\n]*", "", r)
    rust = {m.group(1): [a.strip() for a in m.group(2).split(",") if a.strip()]
            for m in re.finditer(r"pub fn (spl_[a-z_0-9]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", r, flags=re.S)}
    assert set(rust) == set(c), (sorted(set(c) - set(rust)), sorted(set(rust) - set(c)))

    def kind_c(a):
        if "*" in a:
            return "ptr"
        t = a.rsplit(" ", 1)[0].replace("const ", "").strip()
        return {"int": "i32", "uint32_t": "u32", "uint64_t": "u64", "double": "f64", "float": "f32"}[t]

    def kind_rust(a):
        t = a.split(":", 1)[1].strip()
        if t.startswith("*"):
            return "ptr"
        return {"c_int": "i32", "u32": "u32", "u64": "u64", "f64": "f64", "f32": "f32"}[t]

    for name, args in c.items():
        assert [kind_c(a) for a in args] == [kind_rust(a) for a in rust[name]], name
    sigs = {k: v for k, v in vars(_capi).items() if isinstance(v, dict) and "spl_ctx_create" in v}
    assert len(sigs) == 1
    table = next(iter(sigs.values()))
    assert set(table) == set(c)
    for name, (_res, argtypes) in table.items():
        assert len(argtypes) == len(c[name]), name
