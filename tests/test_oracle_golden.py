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
