"""3S* / SSP steps with the stage update fused into the element kernel against rhs! + pointwise stage kernel
(TRIXI_B200_OPT_FUSED_STAGE 1 / 0): ms per step on the headline configuration (TreeMesh, 3D Euler EC, polydeg 3).
Device-timed with the handle's CUDA events; one JSON line per (integrator, variant)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trixi_b200 as T  # noqa: E402
from trixi_b200 import time_integration as ti  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    eq = T.CompressibleEulerEquations3D(1.4)
    solver = T.DGSEM(polydeg=3, surface_flux=T.flux_ranocha,
                     volume_integral=T.VolumeIntegralFluxDifferencing(T.flux_ranocha))
    mesh = T.TreeMesh((-2.0,) * 3, (2.0,) * 3, initial_refinement_level=args.level, periodicity=True)
    semi = T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_weak_blast_wave, solver, device=0)
    u = T.compute_coefficients(0.0, semi).ravel(order="F")
    gpu = semi.backend()
    ndofs = u.size // 5
    for name in ("CarpenterKennedy2N54", "ParsaniKetchesonDeconinck3Sstar94", "ParsaniKetchesonDeconinck3Sstar32",
                 "SimpleSSPRK33"):
        alg = getattr(T, name)()
        for fused in (1, 0):
            if name == "CarpenterKennedy2N54" and not fused:
                continue
            gpu.set_option(gpu.OPT_FUSED_STAGE, fused)
            gpu.upload(0, u)
            dt = 0.5 * gpu.max_dt()
            for _ in range(3):
                ti._stage_loop(gpu, alg, 0.0, dt)
            gpu.synchronize()
            n0 = gpu.launch_count()
            gpu.timer_start()
            for _ in range(args.steps):
                ti._stage_loop(gpu, alg, 0.0, dt)
            ms = gpu.timer_stop() / args.steps
            nl = (gpu.launch_count() - n0) / args.steps
            print(json.dumps({"integrator": name, "fused_stage": bool(fused), "stages": len(alg.c), "ms_per_step": ms,
                              "ms_per_stage": ms / len(alg.c), "launches_per_step": nl,
                              "dof_updates_per_s": ndofs / (ms * 1e-3), "ndofs": ndofs,
                              "finite": bool(np.isfinite(gpu.download(0)).all())}), flush=True)


if __name__ == "__main__":
    main()
