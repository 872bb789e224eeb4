"""Small runs of every tuned kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import trixi_b200 as T  # noqa: E402
from elixirs import ELIXIRS, EXTRA  # noqa: E402

# round 2 adds: the rebuilt headline kernel in its streamed (reduce-add, single-copy face fluxes) and resident forms
# (tree_3d_euler_ec runs both: the fused-CFL last stage keeps a resident u tile), with boundary faces
# (tree_3d_euler_slip_wall_mixed), the curved flux-differencing kernels at p = 3 and p = 5, and the staged halo kernels
# (two in-process ranks)
CASES = ["tree_3d_euler_ec", "tree_3d_euler_source_terms", "tree_3d_euler_ec_shima_etal", "tree_3d_mhd_ec",
         "tree_3d_euler_shockcapturing", "structured_3d_euler_source_terms", "p4est_3d_euler_source_terms_nonperiodic",
         "tree_3d_euler_mortar", "tree_2d_euler_ec", "tree_3d_euler_slip_wall_mixed", "p4est_3d_curved_ec",
         "structured_3d_euler_ec", "p4est_3d_curved_p5",
         # round 2, second half: P4est mortars (2D curved + 3D), mortars with nonconservative terms, GLM-MHD node records,
         # shock capturing and GLM-MHD on curved meshes
         "p4est_2d_advection_nonconforming_flag", "p4est_3d_nonconforming_curved_ec", "tree_3d_mhd_alfven_wave_mortar",
         "structured_3d_euler_sedov", "p4est_2d_euler_sedov", "structured_3d_mhd_ec",
         "p4est_3d_mhd_alfven_wave_nonperiodic",
         # round 2, last part: shock capturing for GLM-MHD (TreeMesh and curved), flux_hlle of the Euler equations
         "tree_3d_mhd_ec_shockcapturing", "structured_3d_mhd_ec_shockcapturing", "p4est_3d_euler_sedov_hlle"]
# the 3S* and SSP stage updates fused into the element kernels (modes 2, 3): one case per kernel family
STAGE_CASES = ["tree_3d_euler_ec", "tree_3d_euler_source_terms", "tree_3d_euler_ec_shima_etal", "tree_3d_mhd_ec",
               "tree_3d_euler_shockcapturing", "structured_3d_euler_ec", "p4est_3d_curved_p5", "tree_2d_euler_ec"]


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
    from trixi_b200 import time_integration as ti
    for name in STAGE_CASES:
        ex = (ELIXIRS.get(name) or EXTRA[name])
        semi = ex.semi()
        gpu = semi.backend()
        gpu.set_option(gpu.OPT_FUSED_CFL, 1)
        gpu.upload(0, T.compute_coefficients(0.0, semi))
        dt = 0.5 * gpu.max_dt()
        for alg_name in ("ParsaniKetchesonDeconinck3Sstar32", "SimpleSSPRK33"):
            ti._stage_loop(gpu, getattr(T, alg_name)(), 0.0, dt)
            dt = 0.5 * gpu.max_dt()
        ok = bool(np.isfinite(gpu.download(0)).all())
        print(name, "fused 3S*/SSP stages", "finite" if ok else "NOT FINITE", flush=True)
        gpu.close()
    if "--halo" in sys.argv:  # (not under compute-sanitizer: it serialises kernels, and a wait kernel spinning for a
        halo_probe()          # pack kernel that cannot start would hang)


def halo_probe():
    """Two ranks in one process: staged pack, signal / wait, staged MPI interface flux."""
    from test_gpu_parity import _ranked_semis
    alg = T.CarpenterKennedy2N54()
    base, semis = _ranked_semis("tree_3d_euler_ec", 2)
    backends = [s.backend() for s in semis]
    blobs = [b.comm_info() for b in backends]
    for b in backends:
        b.comm_connect(blobs)
    u = T.compute_coefficients(0.0, base)
    for b, s in zip(backends, semis):
        a, z = s.cache.first_element, s.cache.last_element
        b.upload(0, np.asfortranarray(u[..., a:z]))
    dt = 0.4 * min(b.max_dt() for b in backends)
    for k in range(2):
        for b in backends:
            b.step_2n(k * dt, dt, alg.a, alg.b, alg.c)
    ok = all(bool(np.isfinite(b.download(0)).all()) for b in backends)
    print("halo (2 in-process ranks)", "finite" if ok else "NOT FINITE", flush=True)


if __name__ == "__main__":
    main()
