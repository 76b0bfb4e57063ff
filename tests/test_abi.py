"""CPU-side checks of the drop-in boundary: libdind.so loads without a GPU, exports every
symbol include/dind.h declares, refuses to evaluate without a device (no CPU fallback), and
the host-side logic that needs no device works."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dind_b200 as D
from dind_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "dind.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dind_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libdind.so does not export {n}"
    assert sorted(s[0] for s in _lib.SYMBOLS) == names, "python binding and header disagree"
    assert lib.dind_version() == 100


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.dind_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.dind_create(C.byref(h), 2, _lib.F64, 0)
    assert rc == _lib.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.dind_last_error()
    with pytest.raises(D.DindError):
        D.NDInterpolation(np.zeros((2, 2)), (D.LinearInterpolationDimension([0.0, 1.0]),) * 2)


def test_product_package_never_touches_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "datainterpolationsnd.jl_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|libdind_oracle|dindo_|oracle/|oracle\.py")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), (dirpath, f)


def test_shard_range_partitions():
    for n, world in ((10, 3), (0, 4), (7, 8), (4_000_000_000, 8), (512, 1)):
        parts = [D.shard_range(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and b >= a
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(D.DindError):
        D.shard_range(10, 3, 3)


def test_host_side_dimension_validation_matches_reference_asserts():
    # validate_t (src/interpolation_utils.jl:34-37)
    with pytest.raises(AssertionError):
        D.LinearInterpolationDimension([0.0, 0.0, 1.0])
    with pytest.raises(AssertionError):
        D.ConstantInterpolationDimension([1.0, 0.5])
    # multiplicity asserts (src/interpolation_dimensions.jl:174-175)
    with pytest.raises(AssertionError):
        D.BSplineInterpolationDimension([0.0, 1.0, 2.0], 2, multiplicities=[3, 1])
    with pytest.raises(AssertionError):
        D.BSplineInterpolationDimension([0.0, 1.0, 2.0], 2, multiplicities=[3, 4, 3])
    d = D.BSplineInterpolationDimension([0.0, 1.0, 2.0], 2)
    assert d.multiplicities.tolist() == [3, 1, 3]            # default clamped (:168-173)
    assert d.knots_all.tolist() == [0, 0, 0, 1, 2, 2, 2]     # expand_knot_vector_kernel
    assert d.n_u() == 4 and len(d) == 3                      # get_n_basis_functions, Base.length
    c = D.ConstantInterpolationDimension([0.0, 1.0], left=False)
    assert c.left is False                                   # stored, never read (dead field in the reference)
