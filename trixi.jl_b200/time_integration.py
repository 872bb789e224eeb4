"""Trixi-native 2N low-storage Runge-Kutta integrators with the OrdinaryDiffEq-like facade.

Mirrors ``src/time_integration/methods_2N.jl:24-48`` (CarpenterKennedy2N54), ``:50-76``
(CarpenterKennedy2N43), ``:95-129`` (SimpleIntegrator2N, init), ``:131-168`` (step!) and
``time_integration.jl:46-55,72-124`` (limit_dt!, callbacks, solve!).  The stage loop itself
(RHS + fused ``u_tmp``/``u`` update, ``methods_2N.jl:144-159``) runs inside
``trixi_b200_step_2n`` on device-resident vectors; the host only sequences steps and callbacks.
"""
from __future__ import annotations

import math

import numpy as np


class CarpenterKennedy2N54:
    """Coefficients from methods_2N.jl:29-46."""

    def __init__(self):
        self.a = np.array([0.0,
                           567301805773.0 / 1357537059087.0,
                           2404267990393.0 / 2016746695238.0,
                           3550918686646.0 / 2091501179385.0,
                           1275806237668.0 / 842570457699.0])
        self.b = np.array([1432997174477.0 / 9575080441755.0,
                           5161836677717.0 / 13612068292357.0,
                           1720146321549.0 / 2090206949498.0,
                           3134564353537.0 / 4481467310338.0,
                           2277821191437.0 / 14882151754819.0])
        self.c = np.array([0.0,
                           1432997174477.0 / 9575080441755.0,
                           2526269341429.0 / 6820363962896.0,
                           2006345519317.0 / 3224310063776.0,
                           2802321613138.0 / 2924317926251.0])


class CarpenterKennedy2N43:
    """Coefficients from methods_2N.jl:60-74."""

    def __init__(self):
        self.a = np.array([0.0, 756391.0 / 934407.0, 36441873.0 / 15625000.0, 1953125.0 / 1085297.0])
        self.b = np.array([8.0 / 141.0, 6627.0 / 2000.0, 609375.0 / 1085297.0, 198961.0 / 526383.0])
        self.c = np.array([0.0, 8.0 / 141.0, 86.0 / 125.0, 1.0])


class ParsaniKetchesonDeconinck3Sstar94:
    """Nine-stage, fourth-order 3S* scheme; coefficients from methods_3Sstar.jl:63-96 (Parsani, Ketcheson,
    Deconinck 2013, DOI 10.1137/120885899)."""

    def __init__(self):
        self.gamma1 = np.array([0.0000000000000000E+00, -4.6556413837561301E+00, -7.7202649689034453E-01,
                                -4.0244202720632174E+00, -2.1296873883702272E-02, -2.4350219407769953E+00,
                                1.9856336960249132E-02, -2.8107894116913812E-01, 1.6894354373677900E-01])
        self.gamma2 = np.array([1.0000000000000000E+00, 2.4992627683300688E+00, 5.8668202764174726E-01,
                                1.2051419816240785E+00, 3.4747937498564541E-01, 1.3213458736302766E+00,
                                3.1196363453264964E-01, 4.3514189245414447E-01, 2.3596980658341213E-01])
        self.gamma3 = np.array([0.0000000000000000E+00, 0.0000000000000000E+00, 0.0000000000000000E+00,
                                7.6209857891449362E-01, -1.9811817832965520E-01, -6.2289587091629484E-01,
                                -3.7522475499063573E-01, -3.3554373281046146E-01, -4.5609629702116454E-02])
        self.beta = np.array([2.8363432481011769E-01, 9.7364980747486463E-01, 3.3823592364196498E-01,
                              -3.5849518935750763E-01, -4.1139587569859462E-03, 1.4279689871485013E+00,
                              1.8084680519536503E-02, 1.6057708856060501E-01, 2.9522267863254809E-01])
        self.delta = np.array([1.0000000000000000E+00, 1.2629238731608268E+00, 7.5749675232391733E-01,
                               5.1635907196195419E-01, -2.7463346616574083E-02, -4.3826743572318672E-01,
                               1.2735870231839268E+00, -6.2947382217730230E-01, 0.0000000000000000E+00])
        self.c = np.array([0.0000000000000000E+00, 2.8363432481011769E-01, 5.4840742446661772E-01,
                           3.6872298094969475E-01, -6.8061183026103156E-01, 3.5185265855105619E-01,
                           1.6659419385562171E+00, 9.7152778807463247E-01, 9.0515694340066954E-01])


class ParsaniKetchesonDeconinck3Sstar32:
    """Three-stage, second-order 3S* scheme; coefficients from methods_3Sstar.jl:114-133."""

    def __init__(self):
        self.gamma1 = np.array([0.0000000000000000E+00, -1.2664395576322218E-01, 1.1426980685848858E+00])
        self.gamma2 = np.array([1.0000000000000000E+00, 6.5427782599406470E-01, -8.2869287683723744E-02])
        self.gamma3 = np.array([0.0, 0.0, 0.0])
        self.beta = np.array([7.2366074728360086E-01, 3.4217876502651023E-01, 3.6640216242653251E-01])
        self.delta = np.array([1.0000000000000000E+00, 7.2196567116037724E-01, 0.0000000000000000E+00])
        self.c = np.array([0.0000000000000000E+00, 7.2366074728360086E-01, 5.9236433182015646E-01])


class SimpleSSPRK33:
    """Shu-Osher SSPRK(3,3) with numerators and denominators kept apart (methods_SSP.jl:23-52); stage callbacks
    (subcell limiters) are outside the hot path and not supported."""

    def __init__(self):
        self.numerator_a = np.array([0.0, 3.0, 1.0])
        self.numerator_b = np.array([1.0, 1.0, 2.0])
        self.denominator = np.array([1.0, 4.0, 3.0])
        self.c = np.array([0.0, 1.0, 0.5])


def _stage_loop(backend, alg, t, dt):
    """Dispatch on the algorithm type like the reference's step! methods (methods_2N.jl:131, methods_3Sstar.jl:173,
    methods_SSP.jl:170)."""
    if isinstance(alg, (ParsaniKetchesonDeconinck3Sstar94, ParsaniKetchesonDeconinck3Sstar32)):
        backend.step_3sstar(t, dt, alg.gamma1, alg.gamma2, alg.gamma3, alg.beta, alg.delta, alg.c)
    elif isinstance(alg, SimpleSSPRK33):
        backend.step_ssp(t, dt, alg.numerator_a, alg.numerator_b, alg.denominator, alg.c)
    else:
        backend.step_2n(t, dt, alg.a, alg.b, alg.c)


class CallbackSet:
    def __init__(self, *callbacks):
        self.discrete_callbacks = [cb for cb in callbacks if cb is not None]


class _Stats:
    def __init__(self):
        self.naccept = 0
        self.nf = 0


class SimpleIntegrator2N:
    """methods_2N.jl:95-111 (and SimpleIntegrator3Sstar methods_3Sstar.jl:136-155, SimpleIntegratorSSP
    methods_SSP.jl:77-102: same facade, the algorithm type selects the stage loop).  ``u`` lives on the device (handle-owned); ``integrator.u`` downloads."""

    def __init__(self, ode, alg, dt, callback, maxiters=None):
        self.semi = self.p = ode.p
        self.backend = ode.p.backend()
        if hasattr(self.backend, "OPT_FUSED_CFL"):
            # the integrator owns u on the device: the last RK stage may also produce the next max_dt
            self.backend.set_option(self.backend.OPT_FUSED_CFL, 1)
        self.alg = alg
        self.t = float(ode.tspan[0])
        self.tspan = ode.tspan
        self.dt = float(dt)
        self.dtcache = float(dt)
        self.iter = 0
        self.stats = _Stats()
        self.callback = callback
        self.maxiters = maxiters if maxiters is not None else np.iinfo(np.int64).max
        self.finalstep = False
        self.backend.upload(self.backend.U, ode.u0)
        self._u0 = ode.u0

    @property
    def u(self):
        return self.download_u()

    def download_u(self):
        flat = np.empty(self.semi.u_length())
        self.backend.download(self.backend.U, flat)
        return flat.reshape(self.semi.u_shape(), order="F")

    def terminate(self):
        self.finalstep = True


def limit_dt(integrator, t_end):
    """time_integration.jl:46-55."""
    if integrator.t + integrator.dt > t_end or math.isclose(integrator.t + integrator.dt, t_end,
                                                             rel_tol=math.sqrt(np.finfo(float).eps), abs_tol=0.0):
        integrator.dt = t_end - integrator.t
        integrator.terminate()


def step(integrator):
    """``step!(integrator::SimpleIntegrator2N)`` (methods_2N.jl:131-168)."""
    alg = integrator.alg
    t_end = integrator.tspan[1]
    assert not integrator.finalstep
    if math.isnan(integrator.dt):
        raise RuntimeError("time step size `dt` is NaN")
    limit_dt(integrator, t_end)
    _stage_loop(integrator.backend, alg, integrator.t, integrator.dt)
    integrator.stats.nf += len(alg.c)
    integrator.iter += 1
    integrator.stats.naccept += 1
    integrator.t += integrator.dt
    if integrator.callback is not None:
        for cb in integrator.callback.discrete_callbacks:
            if cb.condition(integrator):
                cb.affect(integrator)
    if integrator.iter >= integrator.maxiters and not integrator.finalstep:
        integrator.terminate()


def step_2n_host(u_ode, semi, t, dt, alg):
    """The stage loop of ``step!(integrator::SimpleIntegrator2N)`` (methods_2N.jl:144-159) on a host-resident
    ``u_ode`` (updated in place), for callers that keep the integrator's vectors in host memory like the
    reference does: u travels to the device and back inside the call (trixi_b200_step_2n_host)."""
    semi.backend().step_2n_host(u_ode, t, dt, alg.a, alg.b, alg.c)
    return None


class TimeIntegratorSolution:
    def __init__(self, t, u, prob, integrator):
        self.t, self.u, self.prob, self.integrator = t, u, prob, integrator


def init(ode, alg, dt, callback=None, maxiters=None):
    """methods_2N.jl:113-129."""
    integrator = SimpleIntegrator2N(ode, alg, dt, callback, maxiters)
    if callback is not None:
        for cb in callback.discrete_callbacks:
            cb.initialize(integrator)
    return integrator


def solve(ode, alg, dt, callback=None, maxiters=None):
    """``Trixi.solve(ode, alg; dt, callback)`` (time_integration.jl:103-124)."""
    integrator = init(ode, alg, dt, callback, maxiters)
    while not integrator.finalstep:
        step(integrator)
    if callback is not None:
        for cb in callback.discrete_callbacks:
            cb.finalize(integrator)
    return TimeIntegratorSolution((ode.tspan[0], integrator.t), (ode.u0, integrator.download_u()), ode,
                                  integrator)
