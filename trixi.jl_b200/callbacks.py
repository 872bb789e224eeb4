"""Step callbacks.

* ``StepsizeCallback`` (``src/callbacks_step/stepsize.jl:93-154``): the ``max_dt`` reduction is the
  hot-path part and runs on the device (``trixi_b200_max_dt``); the callback only multiplies by cfl.
* ``AnalysisCallback`` (``src/callbacks_step/analysis.jl:227-376``, error norms
  ``analysis_dg3d.jl:123-161`` / ``analysis_dg2d.jl``): stays on the host like in the reference GPU
  path, which copies ``u`` back (``analysis_dg3d.jl:172-177``).  Out of the timed region.
"""
from __future__ import annotations

import numpy as np

from .basis import SolutionAnalyzer
from .parallel import allreduce_min


class _Callback:
    def initialize(self, integrator):
        pass

    def condition(self, integrator):
        return False

    def affect(self, integrator):
        pass

    def finalize(self, integrator):
        pass


class SummaryCallback(_Callback):
    pass


class AliveCallback(_Callback):
    def __init__(self, analysis_interval=0, alive_interval=None):
        pass


class StepsizeCallback(_Callback):
    """``StepsizeCallback(; cfl, interval=1)`` (stepsize.jl:67-68,93-126)."""

    def __init__(self, cfl=1.0, interval=1):
        self.cfl = cfl
        self.interval = interval

    def _cfl(self, t):
        return self.cfl(t) if callable(self.cfl) else self.cfl

    def initialize(self, integrator):
        self.affect(integrator)

    def condition(self, integrator):
        return self.interval > 0 and integrator.stats.naccept % self.interval == 0

    def affect(self, integrator):
        # calculate_dt (stepsize.jl:146-154): cfl(t) * max_dt(u, t, mesh, ...)
        dt = self._cfl(integrator.t) * integrator.backend.max_dt(integrator.t)
        if integrator.semi.world_size > 1:
            # MPI.Allreduce!(dt, min) (stepsize_dg3d.jl:264-279)
            dt = allreduce_min(dt, integrator.semi.comm)
        integrator.dt = dt
        integrator.dtcache = dt

    def __call__(self, ode):
        """``stepsize_callback(ode)`` (stepsize.jl:128-143)."""
        backend = ode.p.backend()
        backend.upload(backend.U, ode.u0)
        dt = self._cfl(ode.tspan[0]) * backend.max_dt(ode.tspan[0])
        return allreduce_min(dt, ode.p.comm) if ode.p.world_size > 1 else dt


class GlmSpeedCallback(_Callback):
    """``GlmSpeedCallback(; glm_scale, cfl)`` (glm_speed.jl:55-121, glm_speed_dg.jl:8-27): after every step
    c_h = glm_scale * dt(c_h = 1) / dt; a host-side scalar update pushed to the device with
    ``trixi_b200_set_eq_param`` (the equations' c_h is mutable per step)."""

    def __init__(self, glm_scale=0.5, cfl=1.0):
        assert 0 <= glm_scale <= 1, "glm_scale must be between 0 and 1"
        self.glm_scale, self.cfl = glm_scale, cfl

    def initialize(self, integrator):
        self.affect(integrator)

    def condition(self, integrator):
        return True

    def affect(self, integrator):
        semi = integrator.semi
        cfl = self.cfl(integrator.t) if callable(self.cfl) else self.cfl
        max_scaled_speed_for_c_h = float(np.max(semi.cache.elements.inverse_jacobian)) * semi.mesh.ndims
        if semi.world_size > 1:
            max_scaled_speed_for_c_h = -allreduce_min(-max_scaled_speed_for_c_h, semi.comm)
        c_h_deltat = cfl * 2 / (semi.solver.nnodes * max_scaled_speed_for_c_h)
        semi.equations.c_h = self.glm_scale * c_h_deltat / integrator.dt
        integrator.backend.set_eq_param(2, semi.equations.c_h)


def multiply_dimensionwise(matrix, data):
    """``multiply_dimensionwise`` (basis_lobatto_legendre.jl / interpolation.jl): apply ``matrix`` along
    every spatial axis of ``data[var, i, j, (k)]`` (leading axis untouched, trailing axes batch)."""
    nd = data.ndim - 2  # [var, i, j, (k), element]
    out = data
    for d in range(nd):
        out = np.moveaxis(np.tensordot(matrix, out, axes=([1], [1 + d])), 0, 1 + d)
    return out


def calc_error_norms(u, t, semi, analyzer=None):
    """L2 / Linf errors against ``initial_condition(x, t)`` on the analysis grid
    (analysis_dg3d.jl:123-161, analysis_dg2d.jl:133-168), conservative variables."""
    mesh, eq, dg, cache = semi.mesh, semi.equations, semi.solver, semi.cache
    if analyzer is None:
        analyzer = SolutionAnalyzer(dg.basis)
    nd = mesh.ndims
    V, w = analyzer.vandermonde, analyzer.weights
    u = np.asarray(u).reshape(semi.u_shape(), order="F")
    u_local = multiply_dimensionwise(V, u)
    x_local = multiply_dimensionwise(V, cache.elements.node_coordinates)
    u_exact = semi.initial_condition(x_local, t, eq)
    diff = u_exact - u_local
    wprod = w
    for _ in range(nd - 1):
        wprod = np.multiply.outer(wprod, w)
    if getattr(semi, "is_curved", False):
        # analysis_dg3d.jl:163-216: the Jacobian is interpolated to the analysis nodes, |J| weights the
        # quadrature and the total volume is accumulated the same way
        jac = 1.0 / cache.elements.inverse_jacobian  # [n.., nelem]
        jac_local = multiply_dimensionwise(V, jac[None])[0]
        weight = wprod[..., None] * np.abs(jac_local)
        total_volume = weight.sum()
    else:
        volume_jacobian = (1.0 / cache.elements.inverse_jacobian) ** nd  # dgsem_tree/dg.jl:8-10
        weight = wprod[..., None] * volume_jacobian  # [na.., nelem]
        total_volume = None
    l2sq = (diff**2 * weight[None]).reshape(eq.nvars, -1).sum(axis=1)
    linf = np.abs(diff).reshape(eq.nvars, -1).max(axis=1)
    if semi.world_size > 1 and semi.comm is not None:
        # global reductions like the reference's MPI analysis (analysis_dg2d.jl:170-215)
        import torch
        # the total volume of a curved mesh is accumulated with the same quadrature and reduced with the
        # squared errors (analysis_dg3d.jl:218-275)
        sums = np.concatenate([l2sq, [0.0 if total_volume is None else total_volume]])
        t2, tinf = torch.from_numpy(sums), torch.from_numpy(linf.copy())
        if semi.comm.get_backend() == "nccl":
            t2, tinf = t2.cuda(), tinf.cuda()
        semi.comm.all_reduce(t2, op=semi.comm.ReduceOp.SUM)
        semi.comm.all_reduce(tinf, op=semi.comm.ReduceOp.MAX)
        sums, linf = t2.cpu().numpy(), tinf.cpu().numpy()
        l2sq = sums[:-1]
        if total_volume is not None:
            total_volume = float(sums[-1])
    l2 = np.sqrt(l2sq / (mesh.total_volume() if total_volume is None else total_volume))
    return l2, linf


_DEVICE_ICS = {
    "LinearScalarAdvectionEquation2D": (1, 2),
    "LinearScalarAdvectionEquation3D": (1, 2),
    "CompressibleEulerEquations2D": (1, 2, 3),
    "CompressibleEulerEquations3D": (1, 2, 3),
    "IdealGlmMhdEquations3D": (1,),
}


def calc_error_norms_device(backend, t, semi, analyzer=None):
    """``calc_error_norms`` with the interpolation to the analysis grid, the exact solution and the reductions
    done by ``trixi_b200_calc_error_norms`` on the resident ``u``.  Returns (None, None) when the backend or
    the initial condition cannot do it (the caller then takes the host path)."""
    if not hasattr(backend, "calc_error_norms") or not hasattr(backend, "lib"):
        return None, None
    ic_id = getattr(semi.initial_condition, "ic_id", 0)
    if ic_id not in _DEVICE_ICS.get(type(semi.equations).__name__, ()):
        return None, None
    if analyzer is None:
        analyzer = SolutionAnalyzer(semi.solver.basis)
    if analyzer.weights.shape[0] > 16:
        return None, None
    l2sq, linf, volume = backend.calc_error_norms(t, ic_id, analyzer.vandermonde, analyzer.weights, semi.equations.nvars)
    curved = getattr(semi, "is_curved", False)
    if semi.world_size > 1 and semi.comm is not None:
        import torch
        t2 = torch.from_numpy(np.concatenate([l2sq, [volume]]))
        tinf = torch.from_numpy(linf.copy())
        if semi.comm.get_backend() == "nccl":
            t2, tinf = t2.cuda(), tinf.cuda()
        semi.comm.all_reduce(t2, op=semi.comm.ReduceOp.SUM)
        semi.comm.all_reduce(tinf, op=semi.comm.ReduceOp.MAX)
        sums, linf = t2.cpu().numpy(), tinf.cpu().numpy()
        l2sq, volume = sums[:-1], float(sums[-1])
    total_volume = volume if curved else semi.mesh.total_volume()
    return np.sqrt(l2sq / total_volume), linf


# analysis_integrals of the AnalysisCallback (analysis.jl:680-760): name -> integrand id of the C ABI
ANALYSIS_INTEGRALS = {"conservation": 0, "entropy": 1, "energy_total": 2, "energy_kinetic": 3, "energy_internal": 4,
                      "entropy_timederivative": 5}


def integrate_device(backend, semi, name, normalize=True):
    """``integrate(func, u, mesh, equations, dg, cache; normalize)`` / ``analyze(entropy_timederivative, du, u, ...)``
    (analysis_dg3d.jl:364-517) on the resident u (and du): quadrature on the device, sums over ranks here, division by
    the total volume last -- exactly where the reference's MPI analysis reduces (analysis_dg2d.jl:170-215)."""
    sums, volume = backend.integrate(ANALYSIS_INTEGRALS[name], semi.equations.nvars)
    if semi.world_size > 1 and semi.comm is not None:
        import torch
        t = torch.from_numpy(np.concatenate([sums, [volume]]))
        if semi.comm.get_backend() == "nccl":
            t = t.cuda()
        semi.comm.all_reduce(t, op=semi.comm.ReduceOp.SUM)
        r = t.cpu().numpy()
        sums, volume = r[:-1], float(r[-1])
    if not normalize:
        return sums
    total = volume if getattr(semi, "is_curved", False) else semi.mesh.total_volume()
    return sums / total


class AnalysisCallback(_Callback):
    """``AnalysisCallback(semi; interval)`` (analysis.jl:95-158): records L2/Linf errors; the final
    values are what the reference's tests compare (``analysis_callback(sol)``, analysis.jl:640-668)."""

    def __init__(self, semi, interval=0, on_device=True, analysis_integrals=()):
        self.semi = semi
        self.interval = interval
        # analysis_integrals (analysis.jl:95-158; the reference's default is (entropy_timederivative,)): evaluated on
        # the device by trixi_b200_integrate; ``integrals`` collects (iter, t, {name: value})
        self.analysis_integrals = tuple(analysis_integrals)
        self.integrals = []
        self.analyzer = SolutionAnalyzer(semi.solver.basis)
        self.history = []
        # reduce the norms on the device when the backend offers it (no download of u)
        self.on_device = on_device

    def initialize(self, integrator):
        # analysis.jl:160-231: the callback also fires once at initialization (iter 0)
        if self.interval > 0:
            self.affect(integrator)

    def condition(self, integrator):
        # analysis.jl:123-127: interval > 0 && (iter % interval == 0 || isfinished(integrator))
        return self.interval > 0 and (integrator.iter % self.interval == 0 or integrator.finalstep)

    def affect(self, integrator):
        l2, linf = None, None
        if self.on_device:
            l2, linf = calc_error_norms_device(integrator.backend, integrator.t, self.semi, self.analyzer)
        if l2 is None:
            l2, linf = calc_error_norms(integrator.download_u(), integrator.t, self.semi, self.analyzer)
        self.history.append((integrator.iter, integrator.t, l2, linf))
        if self.analysis_integrals and hasattr(integrator.backend, "integrate"):
            vals = {}
            for name in self.analysis_integrals:
                if name == "entropy_timederivative":
                    integrator.backend.rhs(integrator.t)  # du of the current state (analysis.jl:349-350)
                vals[name] = integrate_device(integrator.backend, self.semi, name)
            self.integrals.append((integrator.iter, integrator.t, vals))

    def __call__(self, sol):
        l2, linf = calc_error_norms(sol.u[-1], sol.t[-1], self.semi, self.analyzer)
        return l2, linf
