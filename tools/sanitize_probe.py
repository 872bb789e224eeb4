"""Small runs of every tuned kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import trixi_b200 as T  # noqa: E402
from elixirs import ELIXIRS, EXTRA  # noqa: E402

CASES = ["tree_3d_euler_ec", "tree_3d_euler_source_terms", "tree_3d_euler_ec_shima_etal", "tree_3d_mhd_ec",
         "tree_3d_euler_shockcapturing", "structured_3d_euler_source_terms", "p4est_3d_euler_source_terms_nonperiodic",
         "tree_3d_euler_mortar", "tree_2d_euler_ec"]


def main():
    alg = T.CarpenterKennedy2N54()
    for name in CASES:
        ex = (ELIXIRS.get(name) or EXTRA[name])
        semi = ex.semi()
        gpu = semi.backend()
        gpu.set_option(gpu.OPT_FUSED_CFL, 1)
        u = T.compute_coefficients(0.0, semi)
        du = np.empty_like(u)
        T.rhs_hyperbolic(du, u, semi, 0.1)
        gpu.upload(0, u)
        dt = 0.5 * gpu.max_dt()
        for k in range(2):
            gpu.step_2n(k * dt, dt, alg.a, alg.b, alg.c)
            dt = 0.5 * gpu.max_dt()
        ok = bool(np.isfinite(gpu.download(0)).all() and np.isfinite(du).all())
        print(name, "finite" if ok else "NOT FINITE", flush=True)
        gpu.close()


if __name__ == "__main__":
    main()
