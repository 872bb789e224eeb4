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
