#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): 3D compressible Euler,
entropy-conservative flux-differencing DGSEM (flux_ranocha volume + surface flux), polydeg 3, periodic
TreeMesh; metric = DOF-updates/s (DOF x RHS evaluations per second; PID = ns per DOF per RHS is
reported beside it).

A "step" is one CarpenterKennedy2N54 time step = 5 RHS evaluations with the stage update fused in,
plus the CFL max_dt reduction (StepsizeCallback interval 1).  One JSON line on stdout (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--level L] [--impl b200|reference]

`--impl reference` times the CPU restatement of the reference's rhs! (oracle/, OpenMP, all host
cores) on a bounded sample of the same workload; Julia/Trixi itself cannot run here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dof_updates_per_s"
UNIT = "DOF-updates/s"
ALGO_FLOP_PER_DOF_VOLUME = 312.5  # 4.5 pairs x 67 + 11 (p = 3, Euler 3D, flux_ranocha)


CELLS = None  # --cells: elements per direction and rank instead of 2^level (e.g. 100 -> 64 M DOF, BASELINE config 3)


def cells_per_direction(level):
    return CELLS if CELLS else 1 << level


def box_cells(level, world):
    """Weak scaling: every rank keeps a (2^level)^3-element block; the global box doubles in x, y, z in turn."""
    n = [cells_per_direction(level)] * 3
    k, d = world, 0
    while k > 1:
        if k % 2:
            raise ValueError("--gpus must be a power of two")
        n[d % 3] *= 2
        k //= 2
        d += 1
    return tuple(n)


def _warped_mapping_3d(xi_, eta_, zeta_):
    # examples/structured_3d_dgsem/elixir_euler_free_stream.jl:17-41
    pi = np.pi
    xi, eta, zeta = 1.5 * xi_ + 1.5, 1.5 * eta_ + 1.5, 1.5 * zeta_ + 1.5
    y = eta + 3 / 8 * (np.cos(1.5 * pi * (2 * xi - 3) / 3) * np.cos(0.5 * pi * (2 * eta - 3) / 3)
                       * np.cos(0.5 * pi * (2 * zeta - 3) / 3))
    x = xi + 3 / 8 * (np.cos(0.5 * pi * (2 * xi - 3) / 3) * np.cos(2 * pi * (2 * y - 3) / 3)
                      * np.cos(0.5 * pi * (2 * zeta - 3) / 3))
    z = zeta + 3 / 8 * (np.cos(0.5 * pi * (2 * x - 3) / 3) * np.cos(pi * (2 * y - 3) / 3)
                        * np.cos(0.5 * pi * (2 * zeta - 3) / 3))
    return x, y, z


# Secondary workloads (SURVEY.md §8d C2/C4 + the MHD row); `euler_ec` is the headline the driver runs.
# bytes = algorithmic HBM bytes per DOF of the element kernel in RK mode averaged over the 5 CK54 stages
# (stage 1 does not read u_tmp): u + sfv + 0.8 u_tmp + geometry + sources' x, plus the u, u_tmp writes.
WORKLOADS = {
    "euler_ec": {"nvars": 5, "bytes": 212.0, "flop": 312.5 + 45.0,
                 "kernel": "k_element_euler3d_ranocha_p3 (volume+surface+jacobian+2N stage, TMA tiles)"},
    "tgv": {"nvars": 5, "bytes": 212.0, "flop": 312.5 + 45.0,
            "kernel": "k_element_euler3d_ranocha_p3 (volume+surface+jacobian+2N stage, TMA tiles)"},
    "euler_sc": {"nvars": 5, "bytes": 212.0, "flop": None,
                 "kernel": "k_element_fd3d_p3<Euler3D, SC> (blended flux differencing + subcell FV, line sweeps, TMA "
                           "tiles); the indicator kernels are counted under max_dt_ms = 0 / not separately"},
    "euler_shima": {"nvars": 5, "bytes": 212.0, "flop": None,
                    "kernel": "k_element_fd3d_p3<Euler3D> (flux_shima_etal on hoisted node records, line sweeps, TMA tiles)"},
    "euler_weak": {"nvars": 5, "bytes": 212.0 + 24.0, "flop": 149.0 + 45.0 + 25.0,
                   "kernel": "k_element_euler3d_weak_p3 (weak form+surface+jacobian+source+2N stage, TMA tiles)"},
    "structured_curved": {"nvars": 5, "bytes": 212.0 + 72.0 + 8.0, "flop": 149.0 + 45.0 + 45.0,
                          "kernel": "k_element_euler3d_weak_p3<curved> (contravariant fluxes, nodal Jacobian, TMA tiles)"},
    "p4est_curved": {"nvars": 5, "bytes": 212.0 + 72.0 + 8.0, "flop": 149.0 + 45.0 + 45.0,
                     "kernel": "k_element_euler3d_weak_p3<curved> (P4estMesh: + surface integral, TMA tiles)"},
    # curved flux differencing: u + sfv + 0.8 u_tmp + contravariant vectors (72) + nodal
    # Jacobian (8) read, u_tmp + u written
    "structured_ec": {"nvars": 5, "bytes": 212.0 + 72.0 + 8.0, "flop": None,
                      "kernel": "k_element_euler3d_ranocha_curved_p3 (flux_ranocha along averaged contravariant vectors, "
                                "line per thread, TMA tiles)"},
    "p4est_ec": {"nvars": 5, "bytes": 212.0 + 72.0 + 8.0, "flop": None,
                 "kernel": "k_element_euler3d_ranocha_curved_p3 (P4estMesh: + surface integral)"},
    # the reference's own GPU benchmark (benchmark/CUDA/elixir_euler_taylor_green_vortex.jl + run.jl): P4estMesh
    # 4^3 trees, polydeg 5, flux_ranocha volume + flux_lax_friedrichs surface, CarpenterKennedy2N54
    "p4est_tgv_p5": {"nvars": 5, "bytes": 212.0 + 72.0 + 8.0, "flop": None,
                     "kernel": "k_element_euler3d_ranocha_curved_pn<6> (flux_ranocha along averaged contravariant vectors at "
                               "polydeg 5, line per thread, TMA / cp.async tiles)"},
    "mhd_ec": {"nvars": 9, "bytes": 9 * 8 * (1 + 1.5 + 0.8 + 2), "flop": None,
               "kernel": "k_element_fd3d_p3<Mhd3D> (line sweeps, Hindenlang-Gassner + Powell nonconservative, TMA tiles)"},
}


def make_semi(level, device=-1, rank=0, world=1, comm=None, workload="euler_ec"):
    import trixi_b200 as T
    kw = dict(device=device, rank=rank, world_size=world, comm=comm)
    n = cells_per_direction(level)
    if workload in ("euler_ec", "tgv"):
        # examples/tree_3d_dgsem/elixir_euler_ec.jl at a larger refinement level; "tgv" (BASELINE config 5):
        # elixir_euler_taylor_green_vortex.jl = the same volume integral with FluxLaxFriedrichs(max_abs_speed_naive)
        # surface fluxes on [-pi, pi]^3 and the Mach-0.1 vortex as initial condition
        eq = T.CompressibleEulerEquations3D(1.4)
        tgv = workload == "tgv"
        solver = T.DGSEM(polydeg=3,
                         surface_flux=T.FluxLaxFriedrichs(T.max_abs_speed_naive) if tgv else T.flux_ranocha,
                         volume_integral=T.VolumeIntegralFluxDifferencing(T.flux_ranocha))
        half = np.pi if tgv else 2.0
        if world == 1 and not CELLS:
            mesh = T.TreeMesh((-half,) * 3, (half,) * 3, initial_refinement_level=level, periodicity=True)
        else:
            # same cells (size 2 half / 2^level) and the same Morton element order as the TreeMesh, on a box that
            # grows with the number of ranks; contiguous chunks of the order are (2^level)^3 blocks
            mesh = T.CartesianBoxMesh((-half,) * 3, 2 * half / n, box_cells(level, world), periodicity=True)
        ic = T.initial_condition_taylor_green_vortex if tgv else T.initial_condition_weak_blast_wave
        return T.SemidiscretizationHyperbolic(mesh, eq, ic, solver, **kw)
    if world != 1 and workload != "p4est_curved":
        raise ValueError(f"workload {workload} is single-rank")
    if workload == "euler_shima":
        # examples/tree_3d_dgsem/elixir_euler_ec.jl with the kinetic-energy preserving flux_shima_etal (the reference's
        # other SIMD specialization, flux_shima_etal_turbo) as volume flux and LLF surface fluxes
        eq = T.CompressibleEulerEquations3D(1.4)
        solver = T.DGSEM(polydeg=3, surface_flux=T.flux_lax_friedrichs,
                         volume_integral=T.VolumeIntegralFluxDifferencing(T.flux_shima_etal))
        mesh = T.TreeMesh((-2.0,) * 3, (2.0,) * 3, initial_refinement_level=level, periodicity=True)
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_weak_blast_wave, solver, **kw)
    if workload == "euler_weak":
        # examples/tree_3d_dgsem/elixir_euler_source_terms.jl (C2)
        eq = T.CompressibleEulerEquations3D(1.4)
        solver = T.DGSEM(polydeg=3, surface_flux=T.FluxLaxFriedrichs(T.max_abs_speed_naive),
                         volume_integral=T.VolumeIntegralWeakForm())
        mesh = T.TreeMesh((0.0,) * 3, (2.0,) * 3, initial_refinement_level=level, periodicity=True)
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_convergence_test, solver,
                                              source_terms=T.source_terms_convergence_test, **kw)
    if workload in ("structured_curved", "p4est_curved"):
        # examples/structured_3d_dgsem/elixir_euler_free_stream.jl on n^3 cells (C4)
        eq = T.CompressibleEulerEquations3D(1.4)
        solver = T.DGSEM(polydeg=3, surface_flux=T.FluxLaxFriedrichs(T.max_abs_speed_naive),
                         volume_integral=T.VolumeIntegralWeakForm())
        if workload == "structured_curved":
            mesh = T.StructuredMesh((n, n, n), _warped_mapping_3d, periodicity=True)
        else:
            trees = min(n, 4)
            mesh = T.P4estMesh((trees,) * 3, polydeg=3, mapping=_warped_mapping_3d, periodicity=True,
                               initial_refinement_level=int(np.log2(n // trees)))
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_constant, solver, **kw)
    if workload in ("structured_ec", "p4est_ec"):
        # examples/structured_3d_dgsem/elixir_euler_ec.jl / p4est_3d_dgsem/elixir_euler_ec.jl at polydeg 3: entropy
        # conservative flux differencing on the warped mapping (the P4estMesh elixir reads a mesh file; here the same
        # mapping on a programmatic forest)
        eq = T.CompressibleEulerEquations3D(5 / 3)
        solver = T.DGSEM(polydeg=3, surface_flux=T.flux_ranocha,
                         volume_integral=T.VolumeIntegralFluxDifferencing(T.flux_ranocha))
        if workload == "structured_ec":
            mesh = T.StructuredMesh((n, n, n), _warped_mapping_3d, periodicity=True)
        else:
            trees = min(n, 4)
            mesh = T.P4estMesh((trees,) * 3, polydeg=3, mapping=_warped_mapping_3d, periodicity=True,
                               initial_refinement_level=int(np.log2(n // trees)))
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_weak_blast_wave, solver, **kw)
    if workload == "p4est_tgv_p5":
        # benchmark/CUDA/elixir_euler_taylor_green_vortex.jl:29-44 (run.jl: initial_refinement_level = 3 -> 32^3
        # elements = 7.1 M DOF, which is --level 5 here)
        eq = T.CompressibleEulerEquations3D(1.4)
        solver = T.DGSEM(polydeg=5, surface_flux=T.flux_lax_friedrichs,
                         volume_integral=T.VolumeIntegralFluxDifferencing(T.flux_ranocha))
        trees = min(n, 4)
        mesh = T.P4estMesh((trees,) * 3, polydeg=1, coordinates_min=(-np.pi,) * 3, coordinates_max=(np.pi,) * 3,
                           periodicity=True, initial_refinement_level=int(np.log2(n // trees)))
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_taylor_green_vortex, solver, **kw)
    if workload == "euler_sc":
        # examples/tree_3d_dgsem/elixir_euler_shockcapturing.jl
        eq = T.CompressibleEulerEquations3D(1.4)
        basis = T.LobattoLegendreBasis(3)
        indicator = T.IndicatorHennemannGassner(eq, basis, alpha_max=0.5, alpha_min=0.001, alpha_smooth=True,
                                                variable=T.density_pressure)
        volint = T.VolumeIntegralShockCapturingHG(indicator, volume_flux_dg=T.flux_ranocha, volume_flux_fv=T.flux_ranocha)
        solver = T.DGSEM(basis=basis, surface_flux=T.flux_ranocha, volume_integral=volint)
        mesh = T.TreeMesh((-2.0,) * 3, (2.0,) * 3, initial_refinement_level=level, periodicity=True)
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_weak_blast_wave, solver, **kw)
    if workload == "mhd_ec":
        # examples/tree_3d_dgsem/elixir_mhd_ec.jl
        eq = T.IdealGlmMhdEquations3D(1.4)
        eq.c_h = 1.0  # GlmSpeedCallback value of a fixed-dt run; any finite value exercises the same code
        flux = (T.flux_hindenlang_gassner, T.flux_nonconservative_powell)
        solver = T.DGSEM(polydeg=3, surface_flux=flux, volume_integral=T.VolumeIntegralFluxDifferencing(flux))
        mesh = T.TreeMesh((-2.0,) * 3, (2.0,) * 3, initial_refinement_level=level, periodicity=True)
        return T.SemidiscretizationHyperbolic(mesh, eq, T.initial_condition_weak_blast_wave, solver, **kw)
    raise ValueError(f"unknown workload {workload}")


def workload_name(level, world=1, workload="euler_ec"):
    n = cells_per_direction(level)
    if workload == "tgv":
        cells = f"{n}^3 elements per rank" if world > 1 else f"TreeMesh level {level} ({n}^3 elements)"
        return ("tree_3d_dgsem/elixir_euler_taylor_green_vortex.jl: 3D Euler flux differencing (flux_ranocha) + "
                f"LLF(naive) surface flux, polydeg=3, {cells}, {64 * n**3 * world / 1e6:.1f} M DOF total, periodic "
                "[-pi, pi]^3 (Morton-ordered Cartesian box for more than one rank)")
    if workload != "euler_ec":
        desc = {"euler_sc": "tree_3d_dgsem/elixir_euler_shockcapturing.jl: 3D Euler VolumeIntegralShockCapturingHG "
                            "(IndicatorHennemannGassner, flux_ranocha DG/FV/surface), TreeMesh",
                "euler_shima": "tree_3d_dgsem/elixir_euler_ec.jl with flux_shima_etal as volume flux and "
                               "flux_lax_friedrichs surface fluxes, TreeMesh",
                "euler_weak": "tree_3d_dgsem/elixir_euler_source_terms.jl: 3D Euler weak form + LLF(naive) + "
                              "convergence-test sources, TreeMesh",
                "structured_curved": "structured_3d_dgsem/elixir_euler_free_stream.jl: 3D Euler weak form + "
                                     "LLF(naive), warped StructuredMesh",
                "p4est_curved": "the same warped mapping on a conforming P4estMesh (4^3 trees), 3D Euler weak "
                                "form + LLF(naive)",
                "mhd_ec": "tree_3d_dgsem/elixir_mhd_ec.jl: ideal GLM-MHD, flux differencing with "
                          "flux_hindenlang_gassner + flux_nonconservative_powell, TreeMesh",
                "structured_ec": "structured_3d_dgsem/elixir_euler_ec.jl at polydeg 3: 3D Euler EC flux differencing "
                                 "(flux_ranocha) on the warped StructuredMesh",
                "p4est_ec": "p4est_3d_dgsem/elixir_euler_ec.jl at polydeg 3 on a programmatic warped forest (4^3 "
                            "trees): 3D Euler EC flux differencing (flux_ranocha), P4estMesh",
                "p4est_tgv_p5": "benchmark/CUDA/elixir_euler_taylor_green_vortex.jl (the reference's GPU benchmark): "
                                "P4estMesh 4^3 trees, flux_ranocha volume + flux_lax_friedrichs surface"}[workload]
        pd = 5 if workload == "p4est_tgv_p5" else 3
        return (f"{desc}, polydeg={pd}, {n}^3 elements ({(pd + 1)**3 * n**3 / 1e6:.1f} M DOF) per rank, periodic")
    base = ("tree_3d_dgsem/elixir_euler_ec.jl: 3D Euler EC flux differencing (flux_ranocha), polydeg=3, ")
    if world == 1 and not CELLS:
        return base + (f"TreeMesh level {level} ({n}^3 elements, {64 * n**3 / 1e6:.1f} M DOF), periodic, "
                       "weak blast wave IC")
    nx, ny, nz = box_cells(level, world)
    return base + (f"Cartesian box {nx}x{ny}x{nz} elements in TreeMesh (Morton) order = {n}^3 elements "
                   f"({64 * n**3 / 1e6:.1f} M DOF) per rank, {64 * nx * ny * nz / 1e6:.1f} M DOF total, periodic, "
                   "weak blast wave IC")


def sample_name(level, workload):
    """Name of the bounded CPU sample (always a uniform TreeMesh level, whatever --cells says)."""
    global CELLS
    saved, CELLS = CELLS, None
    try:
        return workload_name(level, 1, workload)
    finally:
        CELLS = saved


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t_begin", 0.0), getattr(self, "t_end", float("inf"))
        window = [l for (ts, l) in self.lines if t0 <= ts <= t1 + 0.06]
        if not window:  # very short timed region: fall back to the nearest samples
            window = [l for (_, l) in self.lines[-3:]]
        for line in window:
            parts = [p.strip() for p in line.split(",")][1:]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def host_cores():
    """Cores this process may run on.  Asked for explicitly: torch.distributed.run exports OMP_NUM_THREADS=1, which
    would silently put the OpenMP reference on one core."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the host cores next to its GPU, so that the pinned host buffers it allocates afterwards are
    first-touched on that NUMA node and its H2D/D2H copies do not cross the socket interconnect (the end-to-end path
    is PCIe- and host-memory-bound; with 8 ranks on one socket's memory it collapses).  Returns a short description."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bus = getattr(props, "pci_bus_id", None)
        dom = getattr(props, "pci_domain_id", 0)
        dev = getattr(props, "pci_device_id", 0)
        if bus is None:
            return "pci bus id unknown: affinity unchanged"
        base = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        with open(os.path.join(base, "local_cpulist")) as f:
            cpulist = f.read().strip()
        node = "?"
        try:
            with open(os.path.join(base, "numa_node")) as f:
                node = f.read().strip()
        except OSError:
            pass
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa node {node}, {len(allowed)} cores"
        return f"numa node {node}: no allowed cores, affinity unchanged"
    except Exception as exc:  # noqa: BLE001 (best effort: the bench must run without sysfs)
        return f"affinity unchanged ({type(exc).__name__})"


def oracle_backend(semi, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    return oracle.OracleBackend(semi, num_threads=threads or host_cores())


def time_cpu_reference(level, steps, warmup, workload="euler_ec", turbo=False):
    """The reference's CPU path restated (oracle/trixi_oracle.c, OpenMP over all host cores):
    CarpenterKennedy2N54 steps on a bounded sample (smaller TreeMesh level of the same workload)."""
    import trixi_b200 as T
    global CELLS
    saved, CELLS = CELLS, None  # the bounded CPU sample is always a uniform TreeMesh level
    try:
        semi = make_semi(level, workload=workload)
    finally:
        CELLS = saved
    if turbo:
        # the reference's fastest CPU specialization of this volume integral: flux_ranocha_turbo
        # (dg_3d_compressible_euler.jl:265-617); benchmark_ec.jl itself runs plain flux_ranocha
        semi.solver.volume_integral = T.VolumeIntegralFluxDifferencing(T.flux_ranocha_turbo)
        semi._desc = None
    ob = oracle_backend(semi)
    u0 = T.compute_coefficients(0.0, semi)
    ob.upload(0, u0)
    alg = T.CarpenterKennedy2N54()
    dt = 1.3 * ob.max_dt()
    for _ in range(warmup):
        ob.step_2n(0.0, dt, alg.a, alg.b, alg.c)
    t0 = time.perf_counter()
    for _ in range(steps):
        ob.step_2n(0.0, dt, alg.a, alg.b, alg.c)
        ob.max_dt()
    dt_wall = time.perf_counter() - t0
    ndofs = semi.ndofs()
    value = ndofs * 5 * steps / dt_wall
    return value, dt_wall, ob.num_threads(), ndofs


def run_reference(args, rank, world):
    if rank != 0:
        return
    level = args.cpu_level
    value, wall, threads, ndofs = time_cpu_reference(level, args.steps, args.warmup, args.workload)
    sample = (f"{sample_name(level, args.workload)}; {args.steps} CK54 steps (5 rhs! + stage updates + max_dt each) "
              f"after {args.warmup} warm-up, OpenMP C restatement of the reference's CPU rhs! "
              f"(Julia/Trixi.jl not installable on this box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "pid_ns_per_dof_rhs": 1e9 / value * threads,
        "config": {"workload": workload_name(args.level, 1, args.workload), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "cpu_model": cpu_model(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload in ("euler_ec", "tgv"):
        # for the record: the same sample through the reference's SIMD specialization flux_ranocha_turbo (hoisted
        # logarithms, dg_3d_compressible_euler.jl:265-617); benchmark_ec.jl and this line's value use flux_ranocha
        line["cpu_baseline"]["flux_ranocha_turbo_value"] = time_cpu_reference(level, args.steps, args.warmup,
                                                                              args.workload, turbo=True)[0]
    emit(line)


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import trixi_b200 as T

    torch.cuda.set_device(local_rank)
    all_cores = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: affinity unchanged"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from trixi_b200.parallel import allreduce_min
    semi = make_semi(args.level, device=local_rank, rank=rank, world=world, comm=dist if world > 1 else None,
                     workload=args.workload)
    wl = WORKLOADS[args.workload]
    gpu = semi.backend()
    # u is owned by the handle during the run: the last RK stage also reduces the CFL maxima (max_dt fused)
    gpu.set_option(gpu.OPT_FUSED_CFL, 0 if args.no_fused_cfl else 1)
    if args.prefetch is not None:
        gpu.set_option(gpu.OPT_PREFETCH_DISTANCE, args.prefetch)
    if args.generic_kernels:
        gpu.set_option(gpu.OPT_KERNEL_PATH, 1)
    elif args.kernel_path:
        gpu.set_option(gpu.OPT_KERNEL_PATH, args.kernel_path)
    if args.no_reduce_update:
        gpu.set_option(gpu.OPT_RK_REDUCE_UPDATE, 0)
    if args.two_copy_face_flux:
        gpu.set_option(gpu.OPT_SINGLE_FACE_FLUX, 0)
    if args.l2_hints is not None:
        gpu.set_option(gpu.OPT_L2_HINTS, args.l2_hints)
    ndofs = semi.ndofs()
    u0 = T.compute_coefficients(0.0, semi)
    gpu.upload(0, u0)
    alg = T.CarpenterKennedy2N54()
    cfl = 0.1 if args.workload == "p4est_tgv_p5" else 1.3  # (the benchmark elixir's own StepsizeCallback(cfl = 0.1))

    def new_dt():
        # StepsizeCallback (stepsize.jl:93-126): device max_dt reduction, min over ranks
        local = cfl * gpu.max_dt()
        return allreduce_min(local, dist) if world > 1 else local

    dt = new_dt()

    def one_step(t):
        gpu.step_2n(t, dt_holder[0], alg.a, alg.b, alg.c)
        dt_holder[0] = new_dt()

    dt_holder = [dt]
    t = 0.0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # started before the warm-up so samples exist even for a short timed region
    for _ in range(args.warmup):
        one_step(t)
        t += dt_holder[0]

    # ---- device-resident timed region --------------------------------------------------------------
    gpu.profile_enable(True)
    launches0 = gpu.launch_count()
    barrier()
    sampler.mark_begin()
    gpu.timer_start()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(t)
        t += dt_holder[0]
    ms = gpu.timer_stop()
    barrier()
    sampler.mark_end()
    wall = time.perf_counter() - t_wall0
    launches = gpu.launch_count() - launches0
    elem_ms, elem_n = gpu.profile_read(1)
    surf_ms, surf_n = gpu.profile_read(0)
    cfl_ms, cfl_n = gpu.profile_read(2)
    halo_ms, halo_n = gpu.profile_read(3)
    wait_ms, wait_n = gpu.profile_read(4)
    gpu.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None

    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    u_final = gpu.download(0)
    finite = bool(np.isfinite(u_final).all())
    total_dofs = semi.ndofsglobal()  # weak scaling: every rank owns a (2^level)^3-element block
    value = total_dofs * 5 * args.steps / (ms_max * 1e-3)

    # ---- end-to-end through the public API with host buffers ---------------------------------------
    n = semi.u_length()
    u_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    du_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    u_host, du_host = u_pin.numpy(), du_pin.numpy()
    u_host[:] = u0.ravel(order="F")
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def timed(fn, reps):
        fn()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # (1) the bench's own step through the host-buffer API: one CK54 step on a host-resident u (H2D u, 5 RHS with
    # fused stage updates, D2H u, synchronised) + the CFL step size read back for the next step
    e2e_dt = [dt]

    def host_step():
        T.step_2n_host(u_host, semi, 0.0, e2e_dt[0], alg)
        e2e_dt[0] = new_dt()

    e2e_wall = timed(host_step, e2e_steps)
    e2e_value = total_dofs * 5 * e2e_steps / e2e_wall
    e2e_finite = bool(np.isfinite(u_host).all())
    # (2) a single rhs! call with host buffers (1 RHS per round trip)
    u_host[:] = u0.ravel(order="F")
    rhs_wall = timed(lambda: T.rhs_hyperbolic(du_host, u_host, semi, 0.0), e2e_steps)
    e2e_rhs_value = total_dofs * e2e_steps / rhs_wall
    e2e_finite = e2e_finite and bool(np.isfinite(du_host).all())

    # ---- BASELINE config 5 beside the headline: Taylor-Green vortex, 100^3 elements (64 M DOF) per rank; on 8 GPUs this
    # is the 200^3-element / 512 M DOF problem of the config as written (same metric, own timed region) ----------
    config5 = None
    if args.workload == "euler_ec" and not args.no_config5:
        global CELLS
        saved_cells, CELLS = CELLS, 100
        try:
            semi5 = make_semi(args.level, device=local_rank, rank=rank, world=world, comm=dist if world > 1 else None,
                              workload="tgv")
            gpu5 = semi5.backend()
            gpu5.set_option(gpu5.OPT_FUSED_CFL, 1)
            gpu5.upload(0, T.compute_coefficients(0.0, semi5))
            dt5 = [cfl * gpu5.max_dt()]
            if world > 1:
                dt5[0] = allreduce_min(dt5[0], dist)

            def step5():
                gpu5.step_2n(0.0, dt5[0], alg.a, alg.b, alg.c)
                local = cfl * gpu5.max_dt()
                dt5[0] = allreduce_min(local, dist) if world > 1 else local

            for _ in range(args.warmup):
                step5()
            steps5 = min(args.steps, 10)
            barrier()
            gpu5.timer_start()
            for _ in range(steps5):
                step5()
            ms5 = gpu5.timer_stop()
            barrier()
            t5 = torch.tensor([ms5], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            ms5 = float(t5.item())
            config5 = {"workload": workload_name(args.level, world, "tgv"), "total_dofs": semi5.ndofsglobal(),
                       "value": semi5.ndofsglobal() * 5 * steps5 / (ms5 * 1e-3), "unit": UNIT, "steps": steps5,
                       "ms_per_step": ms5 / steps5,
                       "note": "weak scaling at 64 M DOF per GPU: the 8-GPU line is BASELINE config 5 as written "
                               "(512 M DOF); its 1-GPU line is the per-GPU share to compare with"}
            del gpu5, semi5
        finally:
            CELLS = saved_cells

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        fp64_peak = gpu.measure_fp64_peak()
        copy_gbs = gpu.measure_copy_bandwidth()
        # dominant kernel: the fused element kernel (volume + surface + Jacobian + 2N stage).
        # algorithmic bytes/DOF in RK mode: read u 40 + surface_flux_values 60 + u_tmp 40 (stages 2-5),
        # write u_tmp 40 + u 40 -> 220 B (stage 1: 180 B); average over the 5 stages = 212 B
        elem_avg_ms = elem_ms / max(elem_n, 1)
        algo_bytes = ndofs * wl["bytes"]
        achieved_gbs = algo_bytes / (elem_avg_ms * 1e-3) * 1e-9
        algo_flops = ndofs * (wl["flop"] or 0.0)
        # DRAM traffic of the dominant kernel per launch from the committed `ncu --set full` capture
        # (dram__bytes_read.sum + dram__bytes_write.sum, scaled by DOF count)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f).get(args.workload)
            if tj:
                traffic, traffic_src = tj["dram_bytes_per_dof"] * ndofs, tj["source"]
        cpu = None
        os.sched_setaffinity(0, all_cores)  # the CPU reference below uses every host core again
        if not args.no_cpu_baseline:
            cv, cwall, cthreads, cdofs = time_cpu_reference(args.cpu_level, 5, 2, args.workload)
            turbo = None
            if args.workload in ("euler_ec", "tgv"):
                turbo = time_cpu_reference(args.cpu_level, 5, 2, args.workload, turbo=True)[0]
            cpu = {"value": cv, "unit": UNIT, "cores": cthreads, "cpu_model": cpu_model(), "kind": "port",
                   "flux_ranocha_turbo_value": turbo,
                   "note": "value = the generic flux_differencing_kernel! with flux_ranocha, the configuration "
                           "benchmark/benchmark_ec.jl times; flux_ranocha_turbo_value = the same sample through the "
                           "reference's SIMD specialization with hoisted logarithms "
                           "(dg_3d_compressible_euler.jl:265-617), its fastest CPU path for this volume integral",
                   "sample": f"{sample_name(args.cpu_level, args.workload)}; 5 CK54 steps (25 rhs!) after 2 warm-up; "
                             "OpenMP C restatement of the reference's CPU rhs! (oracle/trixi_oracle.c)"}
        single_copy = args.workload in ("euler_ec", "tgv") and not args.two_copy_face_flux and not args.generic_kernels \
            and args.kernel_path == 0
        curved_wl = "curved" in args.workload or args.workload in ("structured_ec", "p4est_ec", "p4est_tgv_p5")
        nn1 = 6 if args.workload == "p4est_tgv_p5" else 4  # 3 / n face nodes per DOF
        if_bytes = 3.0 / nn1 * ((3 if single_copy else 4) * wl["nvars"] * 8 + (32 if curved_wl else 0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "pid_ns_per_dof_rhs": 1e9 / value * world,
            "config": {"workload": workload_name(args.level, world, args.workload), "ndofs_per_gpu": ndofs,
                       "rhs_per_step": 5, "time_integrator": "CarpenterKennedy2N54 (fused stage update)",
                       "max_dt": "separate kernel after every step" if args.no_fused_cfl else
                       "every step (StepsizeCallback interval 1); reduced by the last RK stage kernel",
                       "l2_hygiene": ("inputs larger than L2 (u alone is %.1f GB)" % (n * 8 / 1e9)) if n * 8 > 2.5e8
                       else "WARNING: working set comparable to the 126 MB L2; use a larger --level",
                       "parallelism": "1 rank per GPU" if world == 1 else
                       f"{world} ranks: Morton-order element partition, face halo exchange by direct peer "
                       "stores over NVLink (CUDA IPC) + sequence flags inside libtrixi_b200, dt min-allreduce "
                       "over NCCL"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
                    "call": "step_2n_host(u_host, semi, t, dt, CarpenterKennedy2N54) -> trixi_b200_step_2n_host + "
                            "trixi_b200_max_dt, pinned host u updated in place; chunked copies overlap the first and "
                            "last stage",
                    "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3, "finite": e2e_finite,
                    "host_affinity_rank0": numa,
                    "rhs_call": {"value": e2e_rhs_value, "unit": UNIT, "ms_per_call": rhs_wall / e2e_steps * 1e3,
                                 "call": "rhs_hyperbolic(du_host, u_host, semi, t) -> trixi_b200_rhs_host: H2D u, "
                                         "1 RHS, D2H du per call, chunked so both PCIe directions overlap",
                                 "h2d_bytes": n * 8, "d2h_bytes": n * 8}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": wl["kernel"],
                         "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved_gbs / peaks["hbm_gbs"], "peak_kind": peak_kind,
                         "peak_note": "the peak is a copy (half reads, half writes); a read-dominated stream such as "
                                      "the weak-form kernels (73% reads) can exceed it",
                         "traffic": traffic,
                         "traffic_source": traffic_src,
                         "avg_launch_ms": elem_avg_ms, "launches": elem_n,
                         "algorithmic_bytes_per_dof": wl["bytes"],
                         "fp64": {"achieved_tflops": algo_flops / (elem_avg_ms * 1e-3) * 1e-12,
                                  "peak_tflops_measured_dfma": fp64_peak,
                                  "algorithmic_flop_per_dof": wl["flop"]},
                         "copy_gbs_measured_here": copy_gbs},
            "roofline_interface_kernel": {
                # prolong2interfaces! + calc_interface_flux! fused: 3/n face nodes per DOF, each reads both states and
                # writes the flux to both elements (+ normal and Jacobian sign on curved meshes)
                "bound": "hbm", "avg_launch_ms": surf_ms / max(surf_n, 1), "launches": surf_n,
                # (euler_ec / tgv: one copy of the flux per interface unless --two-copy-face-flux: 3 instead of 4 records)
                "algorithmic_bytes_per_dof": if_bytes,
                "achieved": (if_bytes * ndofs / (surf_ms / max(surf_n, 1) * 1e-3) * 1e-9) if surf_n else None,
                "peak": peaks["hbm_gbs"], "unit": "GB/s"},
            "kernel_time_share": {"surface_flux_ms": surf_ms, "element_ms": elem_ms, "max_dt_ms": cfl_ms,
                                  "halo_pack_wait_mpiflux_ms": halo_ms + wait_ms, "halo_wait_ms": wait_ms,
                                  "timed_region_ms": ms_max},
            "wall_s": wall, "finite": finite,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if config5:
            line["config5_tgv"] = config5
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _protect_stdout():
    """Libraries (NCCL's version banner, torchrun notices) may print to fd 1; the driver wants exactly one
    JSON line there.  Everything else goes to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--level", type=int, default=7, help="TreeMesh refinement level per GPU (7 = 134 M DOF)")
    ap.add_argument("--cells", type=int, default=None,
                    help="elements per direction and rank on a Morton-ordered Cartesian box instead of a uniform "
                         "TreeMesh level (euler_ec / tgv): 100 -> 64 M DOF per rank")
    ap.add_argument("--cpu-level", type=int, default=6,
                    help="refinement level of the bounded CPU sample (6 = 16.8 M DOF, SURVEY.md §8d)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="euler_ec", choices=sorted(WORKLOADS),
                    help="euler_ec is the headline (BASELINE.json); the others are SURVEY.md §8d's secondary configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true",
                    help="skip the secondary timed region on BASELINE config 5 (TGV, 64 M DOF per GPU)")
    ap.add_argument("--prefetch", type=int, default=None, help="L2 prefetch distance of the tuned element kernel")
    ap.add_argument("--generic-kernels", action="store_true",
                    help="TRIXI_B200_OPT_KERNEL_PATH = 1: the generic one-thread-per-node kernels (before/after numbers)")
    ap.add_argument("--kernel-path", type=int, default=0,
                    help="TRIXI_B200_OPT_KERNEL_PATH: 2 = the previous generation of the tuned headline kernel (A/B runs)")
    ap.add_argument("--no-reduce-update", action="store_true",
                    help="TRIXI_B200_OPT_RK_REDUCE_UPDATE = 0: keep a resident u tile instead of the L2 reduce-add")
    ap.add_argument("--l2-hints", type=int, default=None, help="TRIXI_B200_OPT_L2_HINTS (0/1)")
    ap.add_argument("--two-copy-face-flux", action="store_true",
                    help="TRIXI_B200_OPT_SINGLE_FACE_FLUX = 0: the interface kernel writes both neighbours' copies")
    ap.add_argument("--no-fused-cfl", action="store_true", help="run max_dt as its own kernel after every step")
    args = ap.parse_args()
    global CELLS
    CELLS = args.cells
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
