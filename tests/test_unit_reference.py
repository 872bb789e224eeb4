"""Known-answer unit tests the reference's own test-suite holds for pieces of the hot path (test/test_unit.jl),
restated against the oracle and the host-side basis: they pin the oracle below the level of whole elixir runs."""
import ctypes as C
import math

import numpy as np
import pytest
import trixi_b200 as T
from trixi_b200 import basis as B

FLUX = {"central": 0, "ranocha": 1, "llf": 2, "llf_naive": 3, "hll_davis": 4, "hll_naive": 5, "shima_etal": 6,
        "kennedy_gruber": 7, "chandrashekar": 8, "hlle": 17, "hllc": 18}


def _semi(ndims):
    eq = T.CompressibleEulerEquations3D(1.4) if ndims == 3 else T.CompressibleEulerEquations2D(1.4)
    solver = T.DGSEM(polydeg=3, surface_flux=T.flux_ranocha)
    mesh = T.TreeMesh((-1.0,) * ndims, (1.0,) * ndims, initial_refinement_level=1, periodicity=True)
    return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_constant, solver)


class _Oracle:
    def __init__(self, oracle_module, ndims):
        self.lib = oracle_module.load()
        self.holder = _semi(ndims).descriptor()
        self.nv = ndims + 2
        self.lib.oracle_ln_mean.restype = C.c_double
        self.lib.oracle_inv_ln_mean.restype = C.c_double

    def _p(self, a):
        return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))

    def numflux(self, flux, ul, ur, orientation):
        f = np.zeros(self.nv)
        self.lib.oracle_numflux(self.holder.byref(), C.c_int(FLUX[flux]), self._p(ul), self._p(ur), C.c_int(orientation),
                                f.ctypes.data_as(C.POINTER(C.c_double)))
        return f

    def numflux_normal(self, flux, ul, ur, normal):
        f = np.zeros(self.nv)
        self.lib.oracle_numflux_normal(self.holder.byref(), C.c_int(FLUX[flux]), self._p(ul), self._p(ur), self._p(normal),
                                       f.ctypes.data_as(C.POINTER(C.c_double)))
        return f

    def flux(self, u, orientation):
        f = np.zeros(self.nv)
        self.lib.oracle_flux(self.holder.byref(), self._p(u), C.c_int(orientation), f.ctypes.data_as(C.POINTER(C.c_double)))
        return f

    def flux_normal(self, u, normal):
        f = np.zeros(self.nv)
        self.lib.oracle_flux_normal(self.holder.byref(), self._p(u), self._p(normal),
                                    f.ctypes.data_as(C.POINTER(C.c_double)))
        return f


def test_nodes_and_weights():
    """test/test_unit.jl:353-361."""
    n, w = B.gauss_nodes_weights(1)
    assert list(n) == [0.0] and list(w) == [2.0]
    n, w = B.gauss_nodes_weights(2)
    np.testing.assert_allclose(n, [-1 / math.sqrt(3), 1 / math.sqrt(3)])
    assert list(w) == [1.0, 1.0]
    n, w = B.gauss_nodes_weights(3)
    np.testing.assert_allclose(n, [-math.sqrt(3 / 5), 0.0, math.sqrt(3 / 5)], atol=1e-16)
    np.testing.assert_allclose(w, [5 / 9, 8 / 9, 5 / 9])


def test_boundary_interpolation():
    """test/test_unit.jl:363-377: inverse_weights[1] == Lhat(-1)[1] and the mirror image, polydeg 1..7 (the surface
    integral uses inverse_weights[1] where the strong form would use Lhat, dg_3d.jl:1349)."""
    for p in range(1, 8):
        basis = T.LobattoLegendreBasis(p)
        wbary = B.barycentric_weights(basis.nodes)
        l_minus = B.lagrange_interpolating_polynomials(-1.0, basis.nodes, wbary)
        l_plus = B.lagrange_interpolating_polynomials(1.0, basis.nodes, wbary)
        assert basis.inverse_weights[0] == (l_minus / basis.weights)[0]
        assert basis.inverse_weights[p] == (l_plus / basis.weights)[p]


@pytest.mark.parametrize("ndims", [2, 3])
def test_hll_consistency(ndims, oracle_module):
    """test/test_unit.jl:1604-1640: flux_hll(u, u) == flux(u) along axes and along general normals."""
    o = _Oracle(oracle_module, ndims)
    u = [1.1, -0.5, 2.34, 5.5] if ndims == 2 else [1.1, -0.5, 2.34, 2.4, 5.5]
    for orientation in range(1, ndims + 1):
        np.testing.assert_allclose(o.numflux("hll_davis", u, u, orientation), o.flux(u, orientation), rtol=1e-14)
    normals = ([(1.0, 0.0), (0.0, 1.0), (0.5, -0.5), (-1.2, 0.3)] if ndims == 2 else
               [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.5, -0.5, 0.2), (-1.2, 0.3, 1.4)])
    for n in normals:
        np.testing.assert_allclose(o.numflux_normal("hll_davis", u, u, n), o.flux_normal(u, n), rtol=1e-13)
    # test/test_unit.jl (Consistency check for HLLC flux: CEE): the same for flux_hllc
    for orientation in range(1, ndims + 1):
        np.testing.assert_allclose(o.numflux("hllc", u, u, orientation), o.flux(u, orientation), rtol=1e-14)
    for n in normals:
        np.testing.assert_allclose(o.numflux_normal("hllc", u, u, n), o.flux_normal(u, n), rtol=1e-13)
    # test/test_unit.jl:1803-1849 (Consistency check for HLLE flux: CEE): the same for flux_hlle
    for orientation in range(1, ndims + 1):
        np.testing.assert_allclose(o.numflux("hlle", u, u, orientation), o.flux(u, orientation), rtol=1e-14)
    for n in normals:
        np.testing.assert_allclose(o.numflux_normal("hlle", u, u, n), o.flux_normal(u, n), rtol=1e-13)


def test_rotated_fluxes_3d(oracle_module):
    """test/test_unit.jl:2286-2309 (FluxRotated): along a coordinate axis the normal-direction form of every two-point
    flux equals its orientation form."""
    o = _Oracle(oracle_module, 3)
    u_values = [(1.0, 0.5, -0.7, 0.1, 1.0), (1.5, -0.2, 0.1, 0.2, 5.0)]
    for flux in ["central", "ranocha", "shima_etal", "kennedy_gruber", "hll_davis", "hlle", "hllc", "chandrashekar", "llf",
                 "llf_naive"]:
        for ul in u_values:
            for ur in u_values:
                for d in range(3):
                    n = [0.0, 0.0, 0.0]
                    n[d] = 1.0
                    np.testing.assert_allclose(o.numflux_normal(flux, ul, ur, n), o.numflux(flux, ul, ur, d + 1),
                                               rtol=1e-13, atol=1e-15)
                # two-point fluxes are consistent with the physical flux
                for d in range(3):
                    np.testing.assert_allclose(o.numflux(flux, ul, ul, d + 1), o.flux(ul, d + 1), rtol=1e-13, atol=1e-15)


def test_max_abs_speed_equal_ratio(oracle_module):
    """test/test_unit.jl:2654-2689: with equal p / rho on both sides max_abs_speed_naive == max_abs_speed, so both
    Lax-Friedrichs variants give the same flux."""
    o = _Oracle(oracle_module, 3)
    eq = T.CompressibleEulerEquations3D(1.4)
    ul = eq.prim2cons((np.array(1.0), np.array(0.1), np.array(0.4), np.array(0.9), np.array(11.0))).ravel()
    ur = eq.prim2cons((np.array(2.0), np.array(0.2), np.array(0.3), np.array(0.8), np.array(22.0))).ravel()
    for d in range(1, 4):
        np.testing.assert_allclose(o.numflux("llf", ul, ur, d), o.numflux("llf_naive", ul, ur, d), rtol=1e-14)


def test_ln_mean(oracle_module):
    """ln_mean / inv_ln_mean (math.jl:198-250): both branches against the definition, continuity at the switch."""
    lib = oracle_module.load()
    lib.oracle_ln_mean.restype = C.c_double
    lib.oracle_inv_ln_mean.restype = C.c_double
    for x, y in [(1.0, 2.0), (0.3, 7.5), (1.0, 1.0 + 1e-3), (2.0, 2.0), (5.0, 5.0 * (1 + 2.1e-2)), (5.0, 5.0 * (1 + 1.9e-2))]:
        got = lib.oracle_ln_mean(C.c_double(x), C.c_double(y))
        want = x if x == y else (y - x) / math.log(y / x)
        assert got == pytest.approx(want, rel=1e-12)
        assert lib.oracle_inv_ln_mean(C.c_double(x), C.c_double(y)) == pytest.approx(1 / want, rel=1e-12)


def test_nested_refinement_patches_balance():
    """examples/tree_3d_dgsem/elixir_advection_mortar.jl: two nested box patches on a level-2 TreeMesh.  The second
    patch puts level-4 leaves next to level-2 leaves across x = 0; refine! (abstract_tree.jl:367-403) restores the
    2:1 balance by refining exactly the four coarse neighbours."""
    patches = ({"type": "box", "coordinates_min": (0.0, -1.0, -1.0), "coordinates_max": (1.0, 1.0, 1.0)},
               {"type": "box", "coordinates_min": (0.0, -0.5, -0.5), "coordinates_max": (0.5, 0.5, 0.5)})
    mesh = T.TreeMesh((-1.0,) * 3, (1.0,) * 3, initial_refinement_level=2, refinement_patches=patches, periodicity=True)
    assert list(np.bincount(mesh.levels)) == [0, 0, 28, 256, 256]
    assert not mesh._unbalanced_mask().any()
