"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/feddb200.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "feddb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(feddb200_[a-z0-9_A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from feddlib_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/feddb200.h but not exported"
    # and the Python binding types exactly the declared set
    assert sorted(_lib.SIGNATURES) == syms


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "feddb200.h")).read()
    assert "torch" not in src.lower() and "at::" not in src and "#include <cuda" not in src


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "feddlib_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "fedd_oracle" not in txt, f


def test_create_fails_loudly_without_gpu():
    from feddlib_b200 import _lib
    L = _lib.load()
    if L.feddb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    rc = L.feddb200_create(ctypes.byref(h), 0)
    assert rc == _lib.ERUNTIME
    assert b"no CUDA device" in L.feddb200_last_error()
    from feddlib_b200 import Context, EngineRuntimeError
    with pytest.raises(EngineRuntimeError):
        Context(0, use_torch_stream=False)


def test_product_mesh_generator_matches_oracle_generator():
    from feddlib_b200 import mesh as PM
    from oracle import mesh as OM
    for dim in (2, 3):
        for fe in ("P1", "P2"):
            for N, M, rank in ((1, 3, 0), (2, 2, 1), (2, 2, 3), (2, 3, 2 ** dim - 1)):
                a = OM.structured(dim, fe, N, M, rank)
                b = PM.build_structured(dim, fe, N, M, rank)
                for x, y in zip(a, b):
                    assert np.array_equal(x, y), (dim, fe, N, M, rank)
    with pytest.raises(ValueError):
        PM.build_structured(3, "P1", 1, 0)
    with pytest.raises(ValueError):
        PM.build_structured(3, "Q2", 1, 2)


def test_host_mirror_maps_and_domain():
    from feddlib_b200 import Domain, LogicError, Map
    d = Domain.buildMesh(3, "P2", 1, 2)
    assert d.getApproxEntriesPerRow() == 80 and d.getDimension() == 3 and d.getFEType() == "P2"
    assert d.getMapUnique().getNodeNumElements() == 125
    v = d.getMapVecFieldUnique()
    assert v.getNodeNumElements() == 375 and v.getGlobalElement(7) == 3 * d.getMapUnique().getGlobalElement(2) + 1
    m = Map([5, 9, 2])
    assert m.getLocalElement(9) == 1 and m.getLocalElement(4) == -1 and m.getMaxAllGlobalIndex() == 9
    with pytest.raises(LogicError):
        m.buildVecFieldMap(2, "DimensionWise")
