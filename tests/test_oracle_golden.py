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
