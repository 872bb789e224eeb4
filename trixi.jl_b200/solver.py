"""DG solver types: ``DGSEM`` and the volume/surface integral types (reference
``src/solvers/dgsem/dgsem.jl:65-73``, ``src/solvers/dg.jl:105,135-141,913-918``)."""
from __future__ import annotations

from .basis import LobattoLegendreBasis
from .equations import flux_central, resolve_flux

VOLINT_WEAK_FORM, VOLINT_FLUX_DIFFERENCING, VOLINT_SHOCK_CAPTURING_HG, VOLINT_PURE_LGL_FV = 0, 1, 2, 3


class VolumeIntegralWeakForm:
    """``VolumeIntegralWeakForm`` (solvers/dg.jl:105)."""
    kind = VOLINT_WEAK_FORM
    volume_flux = flux_central  # unused by the weak form

    def __repr__(self):
        return "VolumeIntegralWeakForm()"


class VolumeIntegralFluxDifferencing:
    """``VolumeIntegralFluxDifferencing(volume_flux)`` (solvers/dg.jl:135-141)."""
    kind = VOLINT_FLUX_DIFFERENCING

    def __init__(self, volume_flux):
        resolve_flux(volume_flux)
        self.volume_flux = volume_flux

    def __repr__(self):
        return f"VolumeIntegralFluxDifferencing({self.volume_flux})"


class IndicatorHennemannGassner:
    """``IndicatorHennemannGassner(equations, basis; alpha_max, alpha_min, alpha_smooth, variable)``
    (dgsem/indicators.jl:48-70).  The blending factors are computed on the device by the indicator kernels."""

    def __init__(self, equations, basis, alpha_max=0.5, alpha_min=0.001, alpha_smooth=True, variable=None):
        if variable is None or not hasattr(variable, "var_id"):
            raise TypeError("variable must be density_pressure, density or pressure")
        if equations.nvars < 4:
            raise TypeError("IndicatorHennemannGassner is supported for the compressible Euler equations")
        self.equations, self.basis = equations, basis
        self.alpha_max, self.alpha_min = float(alpha_max), float(alpha_min)
        self.alpha_smooth, self.variable = bool(alpha_smooth), variable


class VolumeIntegralShockCapturingHG:
    """``VolumeIntegralShockCapturingHG(indicator; volume_flux_dg, volume_flux_fv)`` (solvers/dg.jl): blend of the
    flux-differencing volume integral and first-order subcell finite volumes per element."""
    kind = VOLINT_SHOCK_CAPTURING_HG

    def __init__(self, indicator, volume_flux_dg=None, volume_flux_fv=None):
        if not isinstance(indicator, IndicatorHennemannGassner):
            raise TypeError("indicator must be an IndicatorHennemannGassner")
        from .equations import FluxLaxFriedrichs, flux_lax_friedrichs  # noqa: F401
        self.indicator = indicator
        self.volume_flux_dg = volume_flux_dg if volume_flux_dg is not None else flux_central
        self.volume_flux_fv = volume_flux_fv if volume_flux_fv is not None else flux_lax_friedrichs
        resolve_flux(self.volume_flux_dg)
        resolve_flux(self.volume_flux_fv)
        self.volume_flux = self.volume_flux_dg

    def __repr__(self):
        return f"VolumeIntegralShockCapturingHG({self.volume_flux_dg}, {self.volume_flux_fv})"


class VolumeIntegralPureLGLFiniteVolume:
    """``VolumeIntegralPureLGLFiniteVolume(volume_flux_fv)`` (solvers/dg.jl:559-583): first-order finite volumes on the
    LGL subcells of every element."""
    kind = VOLINT_PURE_LGL_FV

    def __init__(self, volume_flux_fv=None):
        from .equations import flux_lax_friedrichs
        self.volume_flux_fv = volume_flux_fv if volume_flux_fv is not None else flux_lax_friedrichs
        resolve_flux(self.volume_flux_fv)
        self.volume_flux = self.volume_flux_fv

    def __repr__(self):
        return f"VolumeIntegralPureLGLFiniteVolume({self.volume_flux_fv})"


class SurfaceIntegralWeakForm:
    """``SurfaceIntegralWeakForm(surface_flux)`` (solvers/dg.jl:829-838)."""

    def __init__(self, surface_flux=flux_central):
        resolve_flux(surface_flux)
        self.surface_flux = surface_flux


class DGSEM:
    """``DGSEM(; polydeg, surface_flux, surface_integral, volume_integral)`` (dgsem.jl:65-73)."""

    def __init__(self, polydeg=None, surface_flux=flux_central, surface_integral=None, volume_integral=None,
                 basis=None):
        # DGSEM(basis, surface_flux, volume_integral) (dgsem.jl:41-63) or DGSEM(; polydeg, ...)
        self.basis = basis if basis is not None else LobattoLegendreBasis(polydeg)
        self.surface_integral = surface_integral or SurfaceIntegralWeakForm(surface_flux)
        self.volume_integral = volume_integral or VolumeIntegralWeakForm()
        if not isinstance(self.volume_integral, (VolumeIntegralWeakForm, VolumeIntegralFluxDifferencing,
                                                 VolumeIntegralShockCapturingHG, VolumeIntegralPureLGLFiniteVolume)):
            # other volume integral types are rejected at the boundary
            raise TypeError("libtrixi_b200 supports VolumeIntegralWeakForm, VolumeIntegralFluxDifferencing and "
                            "VolumeIntegralShockCapturingHG")

    @property
    def polydeg(self):
        return self.basis.polydeg

    @property
    def nnodes(self):
        return self.basis.nnodes
