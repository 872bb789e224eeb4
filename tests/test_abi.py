"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly what
include/trixi_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import trixi_b200 as T
from trixi_b200 import _abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "trixi_b200.h")).read()
    return sorted(set(re.findall(r"TRIXI_B200_API [\w \*]*?(trixi_b200_\w+)\(", hdr)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 20
    l = lib.load_library()
    for name in names:
        assert hasattr(l, name), f"{name} declared in include/trixi_b200.h but not exported"
    assert sorted(lib.EXPORTS) == names
    assert l.trixi_b200_abi_version() == _abi.ABI_VERSION


def test_desc_struct_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "trixi_b200.h")).read()
    body = hdr[hdr.index("typedef struct trixi_b200_desc {"):hdr.index("} trixi_b200_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).replace("typedef struct trixi_b200_desc {", "")
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt or stmt.startswith("typedef"):
            continue
        stmt = re.sub(r"^(const\s+)?(int32_t|int64_t|double)\s*", "", stmt)
        for name in stmt.split(","):
            fields.append(re.sub(r"[\*\s]|\[\d+\]", "", name))
    assert fields == [f[0] for f in _abi.Desc._fields_]


def _small_semi():
    eq = T.CompressibleEulerEquations3D(1.4)
    solver = T.DGSEM(polydeg=3, surface_flux=T.flux_ranocha,
                     volume_integral=T.VolumeIntegralFluxDifferencing(T.flux_ranocha))
    mesh = T.TreeMesh((-2.0,) * 3, (2.0,) * 3, initial_refinement_level=1, periodicity=True)
    return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_weak_blast_wave, solver)


def _have_gpu():
    l = lib.load_library()
    semi = _small_semi()
    h = C.c_void_p()
    rc = l.trixi_b200_create(semi.descriptor().byref(), C.byref(h))
    if rc == 0:
        l.trixi_b200_destroy(h)
    return rc == 0


def test_create_fails_loudly_without_gpu():
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    semi = _small_semi()
    with pytest.raises(lib.TrixiB200Error, match="no CUDA device|CPU fallback"):
        semi.backend()
    ode = T.semidiscretize(semi, (0.0, 0.1))
    du = np.empty_like(ode.u0)
    with pytest.raises(lib.TrixiB200Error):
        T.rhs_hyperbolic(du, ode.u0, semi, 0.0)


def test_create_rejects_bad_descriptors():
    l = lib.load_library()
    semi = _small_semi()
    holder = semi.descriptor()
    h = C.c_void_p()
    holder.desc.abi_version = 99
    assert l.trixi_b200_create(holder.byref(), C.byref(h)) == -1
    assert b"ABI version" in l.trixi_b200_last_error(None)
    holder.desc.abi_version = _abi.ABI_VERSION
    holder.desc.nnodes = 11
    assert l.trixi_b200_create(holder.byref(), C.byref(h)) == -1
    holder.desc.nnodes = 4
    holder.desc.equation = 77
    assert l.trixi_b200_create(holder.byref(), C.byref(h)) == -1
    holder.desc.equation = semi.equations.eq_id
    holder.desc.nmortars = 3
    assert l.trixi_b200_create(holder.byref(), C.byref(h)) == -1
    assert l.trixi_b200_create(None, C.byref(h)) == -1
