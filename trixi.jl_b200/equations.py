"""Equation, numerical-flux, source-term, initial-condition and boundary-condition *types*.

These are the reference's API names (``src/Trixi.jl:204-396`` exports).  On this side of the C-ABI
they are only descriptors: the pointwise arithmetic runs in the CUDA kernels, selected by the enum
ids below (``include/trixi_b200.h``); SURVEY.md §7 "closed-world physics across a C ABI".
Initial conditions are additionally evaluated on the host (NumPy) for ``compute_coefficients`` and
for the AnalysisCallback -- both outside the timed hot path, exactly where the reference runs them
on the CPU (``semidiscretization.jl:224-242``, ``analysis_dg3d.jl:123-161``).
"""
from __future__ import annotations

import math
import numpy as np

# ---- enum ids (must match include/trixi_b200.h) ---------------------------------------------------
EQ_ADVECTION_2D, EQ_EULER_2D, EQ_EULER_3D, EQ_MHD_3D, EQ_ADVECTION_3D = 1, 2, 3, 4, 5

FLUX_CENTRAL = 0
FLUX_RANOCHA = 1
FLUX_LLF = 2            # FluxLaxFriedrichs(max_abs_speed)
FLUX_LLF_NAIVE = 3      # FluxLaxFriedrichs(max_abs_speed_naive)
FLUX_HLL_DAVIS = 4      # FluxHLL(min_max_speed_davis) == flux_hll
FLUX_HLL_NAIVE = 5      # FluxHLL(min_max_speed_naive)
FLUX_SHIMA_ETAL = 6
FLUX_KENNEDY_GRUBER = 7
FLUX_CHANDRASHEKAR = 8
FLUX_HINDENLANG_GASSNER = 9   # MHD, with flux_nonconservative_powell
FLUX_GODUNOV = 10
FLUX_RANOCHA_TURBO = 11
FLUX_LLF_MHD_POWELL = 12      # (flux_lax_friedrichs, flux_nonconservative_powell)
FLUX_HINDENLANG_GASSNER_POWELL = 13
FLUX_LLF_NAIVE_MHD_POWELL = 14  # (FluxLaxFriedrichs(max_abs_speed_naive), flux_nonconservative_powell)
FLUX_HLLE_MHD_POWELL = 15  # (flux_hlle, flux_nonconservative_powell)
FLUX_CENTRAL_MHD_POWELL = 16  # (flux_central, flux_nonconservative_powell)
FLUX_HLLC = 18  # flux_hllc (compressible Euler)
FLUX_HLLE = 17  # flux_hlle = FluxHLL(min_max_speed_einfeldt): compressible Euler; with the Powell term -> 15

SRC_NONE, SRC_CONVERGENCE_TEST, SRC_EOC_TEST_EULER, SRC_EOC_TEST_COUPLED_EULER_GRAVITY = 0, 1, 2, 3

IC_NONE, IC_CONSTANT, IC_CONVERGENCE_TEST, IC_WEAK_BLAST_WAVE, IC_EOC_TEST_COUPLED_EULER_GRAVITY = 0, 1, 2, 3, 4
IC_DENSITY_WAVE = 5

BC_PERIODIC, BC_DIRICHLET, BC_SLIP_WALL = 0, 1, 2


# ---- numerical fluxes ------------------------------------------------------------------------------
class _Flux:
    def __init__(self, name, flux_id):
        self.name, self.flux_id = name, flux_id

    def __repr__(self):
        return self.name


def max_abs_speed():  # marker objects, like the reference's function singletons
    pass


def max_abs_speed_naive():
    pass


def min_max_speed_davis():
    pass


def min_max_speed_naive():
    pass


def min_max_speed_einfeldt():
    pass


flux_central = _Flux("flux_central", FLUX_CENTRAL)
flux_ranocha = _Flux("flux_ranocha", FLUX_RANOCHA)
flux_ranocha_turbo = _Flux("flux_ranocha_turbo", FLUX_RANOCHA_TURBO)
flux_shima_etal = _Flux("flux_shima_etal", FLUX_SHIMA_ETAL)
flux_kennedy_gruber = _Flux("flux_kennedy_gruber", FLUX_KENNEDY_GRUBER)
flux_chandrashekar = _Flux("flux_chandrashekar", FLUX_CHANDRASHEKAR)
flux_hindenlang_gassner = _Flux("flux_hindenlang_gassner", FLUX_HINDENLANG_GASSNER)
flux_godunov = _Flux("flux_godunov", FLUX_GODUNOV)
flux_nonconservative_powell = _Flux("flux_nonconservative_powell", -1)


def FluxLaxFriedrichs(speed=max_abs_speed):
    """``FluxLaxFriedrichs`` (numerical_fluxes.jl:229-253)."""
    if speed is max_abs_speed:
        return _Flux("FluxLaxFriedrichs(max_abs_speed)", FLUX_LLF)
    if speed is max_abs_speed_naive:
        return _Flux("FluxLaxFriedrichs(max_abs_speed_naive)", FLUX_LLF_NAIVE)
    raise ValueError("unsupported wave speed estimate for FluxLaxFriedrichs")


def FluxHLL(speed=min_max_speed_davis):
    """``FluxHLL`` (numerical_fluxes.jl:356-360,422-449)."""
    if speed is min_max_speed_davis:
        return _Flux("FluxHLL(min_max_speed_davis)", FLUX_HLL_DAVIS)
    if speed is min_max_speed_naive:
        return _Flux("FluxHLL(min_max_speed_naive)", FLUX_HLL_NAIVE)
    if speed is min_max_speed_einfeldt:
        return _Flux("FluxHLL(min_max_speed_einfeldt)", FLUX_HLLE)
    raise ValueError("unsupported wave speed estimate for FluxHLL")


flux_lax_friedrichs = FluxLaxFriedrichs()
flux_hll = FluxHLL()
flux_hlle = FluxHLL(min_max_speed_einfeldt)  # numerical_fluxes.jl:457
flux_hllc = _Flux("flux_hllc", FLUX_HLLC)  # compressible_euler_3d.jl:1423-1665, compressible_euler_2d.jl:1720-1925


class _IndicatorVariable:
    """Indicator variables of IndicatorHennemannGassner (``density_pressure`` compressible_euler_3d.jl:1951-1956,
    ``density``/``pressure`` :1937-1949); evaluated inside the indicator kernel, here only a tag."""

    def __init__(self, name, var_id):
        self.name, self.var_id = name, var_id

    def __repr__(self):
        return self.name


density_pressure = _IndicatorVariable("density_pressure", 0)
density = _IndicatorVariable("density", 1)
pressure = _IndicatorVariable("pressure", 2)


def resolve_flux(flux):
    """Map a flux object or a (conservative, nonconservative) tuple to its enum id."""
    if isinstance(flux, tuple):
        cons, noncons = flux
        if noncons is not flux_nonconservative_powell:
            raise ValueError("only flux_nonconservative_powell is supported as nonconservative flux")
        if cons.flux_id == FLUX_HINDENLANG_GASSNER:
            return FLUX_HINDENLANG_GASSNER_POWELL
        if cons.flux_id == FLUX_LLF:
            return FLUX_LLF_MHD_POWELL
        if cons.flux_id == FLUX_LLF_NAIVE:
            return FLUX_LLF_NAIVE_MHD_POWELL
        if cons.flux_id == FLUX_HLLE:
            return FLUX_HLLE_MHD_POWELL
        if cons.flux_id == FLUX_CENTRAL:
            return FLUX_CENTRAL_MHD_POWELL
        raise ValueError(f"unsupported conservative flux {cons} with Powell term")
    if not isinstance(flux, _Flux):
        raise TypeError(f"numerical flux {flux!r} is not in the libtrixi_b200 registry")
    return flux.flux_id


# ---- source terms / boundary conditions --------------------------------------------------------------
class _Tagged:
    def __init__(self, name, tag):
        self.name, self.tag = name, tag

    def __repr__(self):
        return self.name


source_terms_convergence_test = _Tagged("source_terms_convergence_test", SRC_CONVERGENCE_TEST)
source_terms_eoc_test_euler = _Tagged("source_terms_eoc_test_euler", SRC_EOC_TEST_EULER)
source_terms_eoc_test_coupled_euler_gravity = _Tagged("source_terms_eoc_test_coupled_euler_gravity",
                                                      SRC_EOC_TEST_COUPLED_EULER_GRAVITY)

boundary_condition_periodic = _Tagged("boundary_condition_periodic", BC_PERIODIC)
boundary_condition_slip_wall = _Tagged("boundary_condition_slip_wall", BC_SLIP_WALL)


class BoundaryConditionDirichlet:
    """``BoundaryConditionDirichlet(boundary_value_function)`` (equations.jl:159-183).  The boundary
    value function must be one of the registered initial conditions (device enum)."""

    tag = BC_DIRICHLET

    def __init__(self, boundary_value_function):
        ic_id = getattr(boundary_value_function, "ic_id", IC_NONE)
        if ic_id == IC_NONE:
            raise ValueError("BoundaryConditionDirichlet needs a registered initial condition "
                             "(closed-world physics across the C ABI)")
        self.boundary_value_function = boundary_value_function
        self.ic_id = ic_id

    def __repr__(self):
        return f"BoundaryConditionDirichlet({self.boundary_value_function.__name__})"


def _ic(ic_id):
    def deco(f):
        f.ic_id = ic_id
        return f
    return deco


# ---- equations -----------------------------------------------------------------------------------------
class AbstractEquations:
    ndims = 0
    nvars = 0
    eq_id = 0
    have_nonconservative_terms = False
    varnames_cons = ()

    def params(self):
        return [0.0] * 8


class LinearScalarAdvectionEquation2D(AbstractEquations):
    """``LinearScalarAdvectionEquation2D`` (linear_scalar_advection_2d.jl:17-27)."""
    ndims, nvars, eq_id = 2, 1, EQ_ADVECTION_2D
    varnames_cons = ("scalar",)

    def __init__(self, a1, a2=None):
        if a2 is None:
            a1, a2 = a1
        self.advection_velocity = (float(a1), float(a2))

    def params(self):
        return [self.advection_velocity[0], self.advection_velocity[1]] + [0.0] * 6

    def cons2cons(self, u):
        return u


class LinearScalarAdvectionEquation3D(AbstractEquations):
    """``LinearScalarAdvectionEquation3D`` (linear_scalar_advection_3d.jl:17-27)."""
    ndims, nvars, eq_id = 3, 1, 5
    varnames_cons = ("scalar",)

    def __init__(self, a1, a2=None, a3=None):
        if a2 is None:
            a1, a2, a3 = a1
        self.advection_velocity = (float(a1), float(a2), float(a3))

    def params(self):
        return list(self.advection_velocity) + [0.0] * 5

    def cons2cons(self, u):
        return u


class CompressibleEulerEquations2D(AbstractEquations):
    """``CompressibleEulerEquations2D`` (compressible_euler_2d.jl:44-53)."""
    ndims, nvars, eq_id = 2, 4, EQ_EULER_2D
    varnames_cons = ("rho", "rho_v1", "rho_v2", "rho_e_total")

    def __init__(self, gamma):
        self.gamma = float(gamma)
        self.inv_gamma_minus_one = 1.0 / (self.gamma - 1.0)

    def params(self):
        return [self.gamma, self.inv_gamma_minus_one] + [0.0] * 6

    def prim2cons(self, prim):
        rho, v1, v2, p = prim
        return np.stack([rho, rho * v1, rho * v2,
                         p * self.inv_gamma_minus_one + 0.5 * (rho * v1 * v1 + rho * v2 * v2)])

    def cons2prim(self, u):
        rho, rv1, rv2, e = u
        v1, v2 = rv1 / rho, rv2 / rho
        p = (self.gamma - 1) * (e - 0.5 * (rv1 * v1 + rv2 * v2))
        return np.stack([rho, v1, v2, p])


class CompressibleEulerEquations3D(AbstractEquations):
    """``CompressibleEulerEquations3D`` (compressible_euler_3d.jl:45-54)."""
    ndims, nvars, eq_id = 3, 5, EQ_EULER_3D
    varnames_cons = ("rho", "rho_v1", "rho_v2", "rho_v3", "rho_e_total")

    def __init__(self, gamma):
        self.gamma = float(gamma)
        self.inv_gamma_minus_one = 1.0 / (self.gamma - 1.0)

    def params(self):
        return [self.gamma, self.inv_gamma_minus_one] + [0.0] * 6

    def prim2cons(self, prim):
        # compressible_euler_3d.jl:1832-1842
        rho, v1, v2, v3, p = prim
        rv1, rv2, rv3 = rho * v1, rho * v2, rho * v3
        e = p * self.inv_gamma_minus_one + 0.5 * (rv1 * v1 + rv2 * v2 + rv3 * v3)
        return np.stack([rho, rv1, rv2, rv3, e])

    def cons2prim(self, u):
        rho, rv1, rv2, rv3, e = u
        v1, v2, v3 = rv1 / rho, rv2 / rho, rv3 / rho
        p = (self.gamma - 1) * (e - 0.5 * (rv1 * v1 + rv2 * v2 + rv3 * v3))
        return np.stack([rho, v1, v2, v3, p])


class IdealGlmMhdEquations3D(AbstractEquations):
    """``IdealGlmMhdEquations3D`` (ideal_glm_mhd_3d.jl:49-59); ``c_h`` is mutable per step."""
    ndims, nvars, eq_id = 3, 9, EQ_MHD_3D
    have_nonconservative_terms = True
    varnames_cons = ("rho", "rho_v1", "rho_v2", "rho_v3", "rho_e_total", "B1", "B2", "B3", "psi")

    def __init__(self, gamma, initial_c_h=float("nan")):
        self.gamma = float(gamma)
        self.inv_gamma_minus_one = 1.0 / (self.gamma - 1.0)
        self.c_h = float(initial_c_h)

    def params(self):
        return [self.gamma, self.inv_gamma_minus_one, self.c_h] + [0.0] * 5

    def prim2cons(self, prim):
        # ideal_glm_mhd_3d.jl:1273-1284
        rho, v1, v2, v3, p, B1, B2, B3, psi = np.broadcast_arrays(*prim)
        rv1, rv2, rv3 = rho * v1, rho * v2, rho * v3
        e = (p * self.inv_gamma_minus_one + 0.5 * (rv1 * v1 + rv2 * v2 + rv3 * v3)
             + 0.5 * (B1 * B1 + B2 * B2 + B3 * B3) + 0.5 * psi**2)
        return np.stack([rho, rv1, rv2, rv3, e, B1 + 0 * rho, B2 + 0 * rho, B3 + 0 * rho, psi + 0 * rho])


# ---- initial conditions (host, NumPy-vectorised: x has shape [ndims, ...]) ---------------------------
@_ic(IC_CONSTANT)
def initial_condition_constant(x, t, equations):
    shape = x.shape[1:]
    if isinstance(equations, CompressibleEulerEquations3D):
        # compressible_euler_3d.jl:78-86
        vals = (1.0, 0.1, -0.2, 0.7, 10.0)
    elif isinstance(equations, CompressibleEulerEquations2D):
        # compressible_euler_2d.jl:78-85
        vals = (1.0, 0.1, -0.2, 10.0)
    elif isinstance(equations, (LinearScalarAdvectionEquation2D, LinearScalarAdvectionEquation3D)):
        vals = (2.0,)
    elif isinstance(equations, IdealGlmMhdEquations3D):
        # ideal_glm_mhd_3d.jl:101-113 (conservative values)
        vals = (1.0, 0.1, -0.2, -0.5, 50.0, 3.0, -1.2, 0.5, 0.0)
    else:
        raise NotImplementedError
    return np.stack([np.full(shape, v) for v in vals])


@_ic(IC_CONVERGENCE_TEST)
def initial_condition_convergence_test(x, t, equations):
    if isinstance(equations, (LinearScalarAdvectionEquation2D, LinearScalarAdvectionEquation3D)):
        # linear_scalar_advection_2d.jl:67-80, linear_scalar_advection_3d.jl:58-71
        a = equations.advection_velocity
        xs = (x[0] - a[0] * t) + (x[1] - a[1] * t)
        if equations.ndims == 3:
            xs = xs + (x[2] - a[2] * t)
        omega = 2 * math.pi * 0.5
        return (1 + 0.5 * np.sin(omega * xs))[None]
    if isinstance(equations, CompressibleEulerEquations3D):
        # compressible_euler_3d.jl:94-111
        omega = 2 * math.pi * 0.5
        ini = 2 + 0.1 * np.sin(omega * (x[0] + x[1] + x[2] - t))
        return np.stack([ini, ini, ini, ini, ini**2])
    if isinstance(equations, CompressibleEulerEquations2D):
        # compressible_euler_2d.jl:93-109
        omega = 2 * math.pi * 0.5
        ini = 2 + 0.1 * np.sin(omega * (x[0] + x[1] - t))
        return np.stack([ini, ini, ini, ini**2])
    if isinstance(equations, IdealGlmMhdEquations3D):
        # Alfven wave, ideal_glm_mhd_3d.jl:124-150
        p, omega, r, e = 1.0, 2 * math.pi, 2.0, 0.2
        nx, ny = 1 / math.sqrt(r**2 + 1), r / math.sqrt(r**2 + 1)
        sqr = 1.0
        Va = omega / (ny * sqr)
        phi_alv = omega / ny * (nx * (x[0] - 0.5 * r) + ny * (x[1] - 0.5 * r)) - Va * t
        rho = np.ones_like(phi_alv)
        v1 = -e * ny * np.cos(phi_alv) / rho
        v2 = e * nx * np.cos(phi_alv) / rho
        v3 = e * np.sin(phi_alv) / rho
        B1 = nx - rho * v1 * sqr
        B2 = ny - rho * v2 * sqr
        B3 = -rho * v3 * sqr
        return equations.prim2cons((rho, v1, v2, v3, p * rho, B1, B2, B3, 0.0 * rho))
    raise NotImplementedError


def initial_condition_gauss(x, t, equations):
    """linear_scalar_advection_2d.jl:88-94 with the periodic translation x_trans_periodic_2d (:43-49, domain
    length 10, centre 0; Julia's % is the remainder with the sign of the dividend).  Host only."""
    if not isinstance(equations, LinearScalarAdvectionEquation2D):
        raise NotImplementedError
    a = equations.advection_velocity
    xs = []
    for d in range(2):
        shifted = np.fmod(x[d] - a[d] * t, 10.0)
        offset = ((shifted < -5.0).astype(float) - (shifted > 5.0).astype(float)) * 10.0
        xs.append(shifted + offset)
    return np.exp(-(xs[0]**2 + xs[1]**2))[None]


def initial_condition_isentropic_vortex(x, t, equations):
    """The isentropic vortex of examples/tree_2d_dgsem/elixir_euler_vortex_shockcapturing.jl:8-44 (Shu 1997), evaluated
    like the elixir does (t_loc = 0: the centre is not advected).  Host only."""
    if not isinstance(equations, CompressibleEulerEquations2D):
        raise NotImplementedError
    gamma = equations.gamma
    iniamplitude, rho, v1, v2, p = 5.0, 1.0, 1.0, 1.0, 25.0
    rt = p / rho
    cx, cy = -(x[1] - 0.0), x[0] - 0.0  # cross product of the distance to the centre with the z axis
    r2 = cx**2 + cy**2
    du = iniamplitude / (2 * math.pi) * np.exp(0.5 * (1 - r2))
    dtemp = -(gamma - 1) / (2 * gamma * rt) * du**2
    rho_ = rho * (1 + dtemp) ** (1 / (gamma - 1))
    p_ = p * (1 + dtemp) ** (gamma / (gamma - 1))
    return equations.prim2cons((rho_, v1 + du * cx, v2 + du * cy, p_))


def initial_condition_blast_wave(x, t, equations):
    """The "medium blast wave" of examples/tree_2d_dgsem/elixir_euler_blast_wave.jl:7-30 (Hennemann, Gassner 2020,
    Sec. 6.3).  Host only."""
    if not isinstance(equations, CompressibleEulerEquations2D):
        raise NotImplementedError
    r = np.sqrt(x[0]**2 + x[1]**2)
    phi = np.arctan2(x[1], x[0])
    inside = ~(r > 0.5)
    rho = np.where(inside, 1.1691, 1.0)
    v1 = np.where(inside, 0.1882 * np.cos(phi), 0.0)
    v2 = np.where(inside, 0.1882 * np.sin(phi), 0.0)
    p = np.where(inside, 1.245, 1.0e-3)
    return equations.prim2cons((rho, v1, v2, p))


@_ic(IC_WEAK_BLAST_WAVE)
def initial_condition_weak_blast_wave(x, t, equations):
    if isinstance(equations, CompressibleEulerEquations3D):
        # compressible_euler_3d.jl:163-184
        r = np.sqrt(x[0]**2 + x[1]**2 + x[2]**2)
        phi = np.arctan2(x[1], x[0])
        with np.errstate(invalid="ignore", divide="ignore"):
            theta = np.where(r == 0, 0.0, np.arccos(np.where(r == 0, 0.0, x[2] / np.where(r == 0, 1.0, r))))
        outside = r > 0.5
        rho = np.where(outside, 1.0, 1.1691)
        v1 = np.where(outside, 0.0, 0.1882 * np.cos(phi) * np.sin(theta))
        v2 = np.where(outside, 0.0, 0.1882 * np.sin(phi) * np.sin(theta))
        v3 = np.where(outside, 0.0, 0.1882 * np.cos(theta))
        p = np.where(outside, 1.0, 1.245)
        return equations.prim2cons((rho, v1, v2, v3, p))
    if isinstance(equations, IdealGlmMhdEquations3D):
        # ideal_glm_mhd_3d.jl:160-180
        r = np.sqrt(x[0]**2 + x[1]**2 + x[2]**2)
        phi = np.arctan2(x[1], x[0])
        with np.errstate(invalid="ignore", divide="ignore"):
            theta = np.where(r == 0, 0.0, np.arccos(np.where(r == 0, 0.0, x[2] / np.where(r == 0, 1.0, r))))
        outside = r > 0.5
        rho = np.where(outside, 1.0, 1.1691)
        v1 = np.where(outside, 0.0, 0.1882 * np.cos(phi) * np.sin(theta))
        v2 = np.where(outside, 0.0, 0.1882 * np.sin(phi) * np.sin(theta))
        v3 = np.where(outside, 0.0, 0.1882 * np.cos(theta))
        p = np.where(outside, 1.0, 1.245)
        one = np.ones_like(rho)
        return equations.prim2cons((rho, v1, v2, v3, p, one, one, one, 0.0 * one))
    if isinstance(equations, CompressibleEulerEquations2D):
        # compressible_euler_2d.jl:181-199
        r = np.sqrt(x[0]**2 + x[1]**2)
        phi = np.arctan2(x[1], x[0])
        sin_phi, cos_phi = np.sin(phi), np.cos(phi)
        outside = r > 0.5
        rho = np.where(outside, 1.0, 1.1691)
        v1 = np.where(outside, 0.0, 0.1882 * cos_phi)
        v2 = np.where(outside, 0.0, 0.1882 * sin_phi)
        p = np.where(outside, 1.0, 1.245)
        return equations.prim2cons((rho, v1, v2, p))
    raise NotImplementedError


@_ic(IC_EOC_TEST_COUPLED_EULER_GRAVITY)
def initial_condition_eoc_test_coupled_euler_gravity(x, t, equations):
    # compressible_euler_3d.jl:196-215 (gamma must be 2)
    if equations.gamma != 2:
        raise ValueError("adiabatic constant must be 2 for the coupling convergence test")
    if isinstance(equations, CompressibleEulerEquations3D):
        s = x[0] + x[1] + x[2] - t
        ini = 2 + 0.1 * _sinpi(s)
        one = np.ones_like(ini)
        p = ini**2 * 1 * 2 / (3 * math.pi)
        return equations.prim2cons((ini, one, one, one, p))
    s = x[0] + x[1] - t
    ini = 2 + 0.1 * _sinpi(s)
    one = np.ones_like(ini)
    p = ini**2 * 1 / math.pi
    return equations.prim2cons((ini, one, one, p))


def _sinpi(s):
    # Julia's sinpi is exact at integers/half-integers; reduce the argument first
    r = np.mod(s, 2.0)
    return np.sin(math.pi * np.where(r > 1.0, r - 2.0, r))


def initial_condition_taylor_green_vortex(x, t, equations):
    # examples/tree_3d_dgsem/elixir_euler_taylor_green_vortex.jl:14-31
    A, Ms, rho = 1.0, 0.1, 1.0
    v1 = A * np.sin(x[0]) * np.cos(x[1]) * np.cos(x[2])
    v2 = -A * np.cos(x[0]) * np.sin(x[1]) * np.cos(x[2])
    v3 = np.zeros_like(v1)
    p = (A / Ms)**2 * rho / equations.gamma
    p = p + 1.0 / 16.0 * A**2 * rho * (np.cos(2 * x[0]) * np.cos(2 * x[2]) + 2 * np.cos(2 * x[1])
                                       + 2 * np.cos(2 * x[0]) + np.cos(2 * x[1]) * np.cos(2 * x[2]))
    return equations.prim2cons((np.full_like(v1, rho), v1, v2, v3, p))


def initial_condition_density_pulse(x, t, equations):
    # examples/tree_3d_dgsem/elixir_euler_density_pulse.jl:14-25
    rho = 1 + np.exp(-(x[0]**2 + x[1]**2 + x[2]**2)) / 2
    p = 1.0
    e = p / (equations.gamma - 1) + 1 / 2 * rho * 3.0
    return np.stack([rho, rho, rho, rho, e])
