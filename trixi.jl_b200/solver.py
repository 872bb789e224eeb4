"""DG solver types: ``DGSEM`` and the volume/surface integral types (reference
``src/solvers/dgsem/dgsem.jl:65-73``, ``src/solvers/dg.jl:105,135-141,913-918``)."""
from __future__ import annotations

from .basis import LobattoLegendreBasis
from .equations import flux_central, resolve_flux

VOLINT_WEAK_FORM, VOLINT_FLUX_DIFFERENCING = 0, 1


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


class SurfaceIntegralWeakForm:
    """``SurfaceIntegralWeakForm(surface_flux)`` (solvers/dg.jl:829-838)."""

    def __init__(self, surface_flux=flux_central):
        resolve_flux(surface_flux)
        self.surface_flux = surface_flux


class DGSEM:
    """``DGSEM(; polydeg, surface_flux, surface_integral, volume_integral)`` (dgsem.jl:65-73)."""

    def __init__(self, polydeg, surface_flux=flux_central, surface_integral=None, volume_integral=None):
        self.basis = LobattoLegendreBasis(polydeg)
        self.surface_integral = surface_integral or SurfaceIntegralWeakForm(surface_flux)
        self.volume_integral = volume_integral or VolumeIntegralWeakForm()
        if not isinstance(self.volume_integral, (VolumeIntegralWeakForm, VolumeIntegralFluxDifferencing)):
            # other volume integral types are rejected at the boundary
            raise TypeError("libtrixi_b200 supports VolumeIntegralWeakForm and VolumeIntegralFluxDifferencing")

    @property
    def polydeg(self):
        return self.basis.polydeg

    @property
    def nnodes(self):
        return self.basis.nnodes
