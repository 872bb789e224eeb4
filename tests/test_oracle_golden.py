"""Pins the CPU oracle (oracle/trixi_oracle.c) against the reference's own golden vectors: every
elixir below is run to its final time through the oracle RHS + the 2N integrator and must reproduce
the L2/Linf errors hard-coded in the reference's test-suite (tests/elixirs.py cites file:line)."""
import numpy as np
import pytest

from elixirs import ELIXIRS


@pytest.mark.parametrize("name", sorted(ELIXIRS))
def test_oracle_reproduces_reference_golden(name, oracle_module):
    ex = ELIXIRS[name]
    semi = ex.semi()
    semi.set_backend(oracle_module.OracleBackend(semi))
    sol, l2, linf = ex.run(semi)
    ex.check(l2, linf)


def test_p4est_curved_equals_structured(oracle_module):
    """A conforming P4estMesh whose trees carry the same polynomial geometry as a StructuredMesh gives the same
    right-hand side (different face bookkeeping: node_indices + outward normals vs left_neighbors + signed
    contravariant vectors)."""
    import numpy as np
    import trixi_b200 as T
    from elixirs import EXTRA
    a, b = EXTRA["p4est_3d_curved_ec"].semi(), EXTRA["structured_3d_like_p4est_curved"].semi()
    u = T.compute_coefficients(0.0, a)
    rng = np.random.default_rng(0)
    u = np.asfortranarray(u * (1 + 0.1 * rng.uniform(-1, 1, u.shape)))
    da, db = np.empty_like(u), np.empty_like(u)
    oracle_module.OracleBackend(a).rhs_host(da, u, 0.1)
    oracle_module.OracleBackend(b).rhs_host(db, u, 0.1)
    assert np.abs(da - db).max() <= 1e-12 * np.abs(db).max()


def test_p4est_brick_equals_treemesh(oracle_module):
    """examples/p4est_3d_dgsem/elixir_euler_source_terms.jl is the TreeMesh elixir on a 4^3-tree forest at
    level 1: same elements, different order."""
    import numpy as np
    import trixi_b200 as T
    from elixirs import ELIXIRS, EXTRA
    a = EXTRA["p4est_3d_periodic_source_terms"].semi()
    b = ELIXIRS["tree_3d_euler_source_terms"].semi(level=3)
    ua, ub = T.compute_coefficients(0.0, a), T.compute_coefficients(0.0, b)
    da, db = np.empty_like(ua), np.empty_like(ub)
    oracle_module.OracleBackend(a).rhs_host(da, ua, 0.2)
    oracle_module.OracleBackend(b).rhs_host(db, ub, 0.2)
    # match elements by their first node's coordinates
    ka = np.round(a.cache.elements.node_coordinates[:, 0, 0, 0, :] * 1e6).astype(np.int64)
    kb = np.round(b.cache.elements.node_coordinates[:, 0, 0, 0, :] * 1e6).astype(np.int64)
    oa, ob = np.lexsort(ka), np.lexsort(kb)
    assert np.array_equal(ka[:, oa], kb[:, ob])
    # the forest's metric terms go through the curl-invariant form (containers_3d.jl:125-285): 4e-13 relative
    # round-off on Ja, amplified by inverse_jacobian = 512 and |flux| ~ 10 against the TreeMesh's exact 2/dx
    assert np.abs(da[..., oa] - db[..., ob]).max() <= 1e-9 * np.abs(db).max()


def test_turbo_kernel_matches_generic(oracle_module):
    """The reference's performance specialization for flux_ranocha_turbo (dg_3d_compressible_euler.jl:265-617, hoisted
    logarithms) against the generic flux_differencing_kernel! -- its own parity test is
    test/test_performance_specializations_3d.jl:49-89 (isapprox of du)."""
    import trixi_b200 as T

    def semi(flux):
        eq = T.CompressibleEulerEquations3D(1.4)
        solver = T.DGSEM(polydeg=3, surface_flux=T.flux_ranocha, volume_integral=T.VolumeIntegralFluxDifferencing(flux))
        mesh = T.TreeMesh((-2.0,) * 3, (2.0,) * 3, initial_refinement_level=2, periodicity=True)
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_weak_blast_wave, solver)

    a, b = semi(T.flux_ranocha), semi(T.flux_ranocha_turbo)
    u = T.compute_coefficients(0.0, a)
    rng = np.random.default_rng(0)
    for amplitude in (0.2, 1e-3):  # regular and series branch of the logarithmic means
        v = np.asfortranarray(u * (1 + amplitude * rng.uniform(-1, 1, (1,) + u.shape[1:])))
        da, db = np.empty_like(v), np.empty_like(v)
        oracle_module.OracleBackend(a).rhs_host(da, v, 0.0)
        oracle_module.OracleBackend(b).rhs_host(db, v, 0.0)
        assert np.abs(da - db).max() <= 1e-13 * np.abs(da).max()


def test_p4est_nonconforming_free_stream(oracle_module):
    """examples/p4est_3d_dgsem/elixir_euler_free_stream.jl (test/test_p4est_3d.jl:164-184) on a programmatic forest:
    with a mesh polydeg of half the solver polydeg a constant state stays constant across curved hanging faces
    (mortar normals taken from the small elements, the large side scaled by 4) -- and does not with a full-degree
    mesh, which is what the reference's comment in that elixir says."""
    import trixi_b200 as T
    from elixirs import EXTRA
    ex = EXTRA["p4est_3d_free_stream_nonconforming"]
    semi = ex.semi()
    assert semi.cache.mortars.nmortars > 0 and semi.cache.boundaries.nboundaries > 0
    u = T.compute_coefficients(0.0, semi)
    du = np.empty_like(u)
    oracle_module.OracleBackend(semi).rhs_host(du, u, 0.0)
    assert np.abs(du).max() < 5e-10
    # the reference's run: t = 0.03, errors of 1e-14 .. 1e-11
    semi.set_backend(oracle_module.OracleBackend(semi))
    ode = T.semidiscretize(semi, (0.0, 0.03))
    analysis = T.AnalysisCallback(semi, interval=100)
    sol = T.solve(ode, T.CarpenterKennedy2N54(), dt=1.0,
                  callback=T.CallbackSet(analysis, T.StepsizeCallback(cfl=1.2)))
    l2, linf = analysis(sol)
    assert np.all(l2 < 1e-12) and np.all(linf < 5e-11)


def test_p4est_refine_and_balance():
    """refine_p4est! + balance! on a periodic 3 x 2 forest (elixir_advection_nonconforming_flag.jl): the corner
    quadrant of every tree goes from level 1 to 4, the 2:1 face balance ripples outwards and across the periodic
    boundary; leaves tile the forest exactly and stay in Morton order per tree."""
    from elixirs import ELIXIRS
    semi = ELIXIRS["p4est_2d_advection_nonconforming_flag"].semi()
    mesh = semi.mesh
    assert mesh.ncells == 168 and int(mesh.levels.max()) == 4 and int(mesh.levels.min()) >= 1
    area = np.sum(0.25 ** mesh.levels.astype(float))
    assert abs(area - 6.0) < 1e-14
    assert np.all(np.diff(mesh.tree_of_element) >= 0)
    # every hanging face has exactly one level of difference
    mo = semi.cache.mortars
    lv = mesh.levels
    assert np.all(lv[mo.neighbor_ids[:-1] - 1] == lv[mo.neighbor_ids[-1] - 1][None, :] + 1)
    # faces are counted once: 4 per element = 2 interfaces + (1 + 2) mortar sides
    assert 2 * semi.cache.interfaces.ninterfaces + 3 * mo.nmortars == 4 * mesh.ncells


def test_goldens_are_the_reference_test_suite_values():
    """tests/golden/reference_goldens.json was extracted from the reference's test/*.jl by
    tests/golden/extract_reference_goldens.py; the values the elixirs are checked against are those, bit for bit."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")) as f:
        golden = json.load(f)
    assert sorted(golden) == sorted(ELIXIRS)
    for name, ex in ELIXIRS.items():
        assert golden[name]["l2"] == list(ex.l2) and golden[name]["linf"] == list(ex.linf), name
        assert golden[name]["source"] == ex.source
