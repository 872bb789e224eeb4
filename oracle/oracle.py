"""TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

ctypes wrapper of ``oracle/libtrixi_oracle.so`` (the C restatement of the reference's CPU hot path,
``oracle/trixi_oracle.c``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs may import this module.  ``OracleBackend`` exposes the same
methods as ``trixi_b200.lib.B200Backend`` so the host-side integrator/callbacks can be driven by
either; the product never constructs it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libtrixi_oracle.so")
_LIB_NOFMA = os.path.join(_HERE, "libtrixi_oracle_nofma.so")
_lib = None
_lib_nofma = None


def build(force=False):
    src = os.path.join(_HERE, "trixi_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "trixi_b200.h")
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if (not force and all(os.path.exists(p) and os.path.getmtime(p) >= newest for p in (_LIB, _LIB_NOFMA))):
        return _LIB
    subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    return _LIB


def _open(path):
    lib = C.CDLL(path)
    lib.oracle_max_dt.restype = C.c_double
    lib.oracle_num_threads.restype = C.c_int
    return lib


def load(nofma=False):
    global _lib, _lib_nofma
    build()
    if nofma:
        if _lib_nofma is None:
            _lib_nofma = _open(_LIB_NOFMA)
        return _lib_nofma
    if _lib is None:
        _lib = _open(_LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleBackend:
    U, DU, U_TMP = 0, 1, 2

    def __init__(self, semi, num_threads=None, nofma=False):
        self.lib = load(nofma)
        if num_threads is not None:
            self.lib.oracle_set_num_threads(int(num_threads))
        self.semi = semi
        self.holder = semi.descriptor()
        self.desc = self.holder.desc
        n = semi.u_length()
        self.u_length = n
        self.vec = [np.zeros(n), np.zeros(n), np.zeros(n)]
        d = self.desc
        nf = d.nnodes ** (d.ndims - 1)
        self.interfaces_u = np.zeros(max(2 * d.nvars * nf * max(d.ninterfaces, 1),
                                         d.nvars * nf * 2 * d.ndims * d.nelements))
        self.boundaries_u = np.zeros(2 * d.nvars * nf * max(d.nboundaries, 1))
        self.sfv = np.zeros(d.nvars * nf * 2 * d.ndims * d.nelements)
        self.mpi_u = np.zeros(2 * d.nvars * nf * max(d.nmpiinterfaces, 1))
        self.halo = None  # set_halo_exchange(callable) for world_size > 1
        self.nrhs = 0

    def num_threads(self):
        return self.lib.oracle_num_threads()

    # same surface as B200Backend ------------------------------------------------------------------
    def upload(self, which, host):
        self.vec[which][:] = np.asarray(host).ravel(order="F")

    def download(self, which, host=None):
        if host is None:
            return self.vec[which].copy()
        host[:] = self.vec[which]
        return host

    def synchronize(self):
        pass

    def set_eq_param(self, index, value):
        self.desc.eq_params[index] = float(value)

    def integrate(self, quantity, nvars):
        """same surface as B200Backend.integrate: (sums, volume) of vec[0] (and vec[1] for the entropy time derivative)"""
        out, vol = np.zeros(nvars), C.c_double(0.0)
        self.lib.oracle_integrate(self.holder.byref(), C.c_int(int(quantity)), _p(self.vec[0]), _p(self.vec[1]), _p(out),
                                  C.byref(vol))
        return (out if quantity == 0 else out[:1]), float(vol.value)

    def set_halo_exchange(self, exchange):
        """``exchange(mpi_u_flat)`` fills the remote side of mpi_u (tests: gloo isend/irecv)."""
        self.halo = exchange

    def rhs_arrays(self, du, u, t):
        h = self.holder.byref()
        if self.desc.world_size > 1:
            # rhs_hyperbolic! for distributed meshes (dg_2d_parallel.jl:453-563): local work, halo
            # exchange, then the MPI interface fluxes and the element-local tail
            if self.halo is None:
                raise RuntimeError("world_size > 1 needs set_halo_exchange()")
            self.lib.oracle_rhs_parallel_part1(h, _p(du), _p(u), C.c_double(t), _p(self.interfaces_u),
                                               _p(self.boundaries_u), _p(self.sfv), _p(self.mpi_u))
            self.halo(self.mpi_u)
            self.lib.oracle_rhs_parallel_part2(h, _p(du), _p(u), C.c_double(t), _p(self.sfv), _p(self.mpi_u))
        else:
            self.lib.oracle_rhs(h, _p(du), _p(u), C.c_double(t), _p(self.interfaces_u),
                                _p(self.boundaries_u), _p(self.sfv))
        self.nrhs += 1

    def rhs_host(self, du_host, u_host, t):
        u = np.ascontiguousarray(np.asarray(u_host).ravel(order="K"))
        du = np.empty_like(u)
        self.rhs_arrays(du, u, float(t))
        flat = du_host.ravel(order="K")
        if not np.shares_memory(flat, du_host):
            raise ValueError("du_host must be contiguous")
        flat[:] = du

    def rhs(self, t):
        self.rhs_arrays(self.vec[1], self.vec[0], float(t))

    def max_dt(self, t=0.0):
        return self.lib.oracle_max_dt(self.holder.byref(), _p(self.vec[0]))

    def step_2n(self, t, dt, a, b, c):
        a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (a, b, c))
        if self.desc.world_size > 1:
            # stage loop of methods_2N.jl:144-159 around the distributed RHS
            u, du, u_tmp = self.vec
            u_tmp[:] = 0.0
            for s in range(len(c)):
                self.rhs_arrays(du, u, t + dt * c[s])
                u_tmp[:] = du - u_tmp * a[s]
                u += u_tmp * (b[s] * dt)
            return
        self.lib.oracle_step_2n(self.holder.byref(), _p(self.vec[0]), _p(self.vec[1]), _p(self.vec[2]),
                                C.c_double(t), C.c_double(dt), _p(a), _p(b), _p(c), C.c_int(len(c)),
                                _p(self.interfaces_u), _p(self.boundaries_u), _p(self.sfv))
        self.nrhs += len(c)

    def step_3sstar(self, t, dt, gamma1, gamma2, gamma3, beta, delta, c):
        """step!(integrator::SimpleIntegrator3Sstar) stage loop (methods_3Sstar.jl:186-207)."""
        u, du, u_tmp1 = self.vec
        u_tmp1[:] = 0.0
        u_tmp2 = u.copy()
        for s in range(len(c)):
            self.rhs_arrays(du, u, t + dt * c[s])
            u_tmp1 += delta[s] * u
            u[:] = gamma1[s] * u + gamma2[s] * u_tmp1 + gamma3[s] * u_tmp2 + (beta[s] * dt) * du

    def step_ssp(self, t, dt, numerator_a, numerator_b, denominator, c):
        """step!(integrator::SimpleIntegratorSSP) stage loop without stage callbacks (methods_SSP.jl:185-202)."""
        u, du, u_tmp = self.vec
        u_tmp[:] = u
        for s in range(len(c)):
            self.rhs_arrays(du, u, t + dt * c[s])
            u += dt * du
            u[:] = (numerator_a[s] * u_tmp + numerator_b[s] * u) / denominator[s]

    def calc_indicator(self):
        """IndicatorHennemannGassner blending factors of the resident u (dgsem/indicators.jl:114-148)."""
        alpha = np.zeros(self.desc.nelements)
        self.lib.oracle_calc_indicator_hg(self.holder.byref(), _p(alpha), _p(self.vec[0]))
        return alpha

    # stage-level ------------------------------------------------------------------------------------
    def calc_volume_integral(self):
        self.lib.oracle_set_zero(self.holder.byref(), _p(self.vec[1]))
        if self.desc.mesh_kind != 0:
            self.lib.oracle_calc_volume_integral_curved(self.holder.byref(), _p(self.vec[1]), _p(self.vec[0]))
            return
        self.lib.oracle_calc_volume_integral(self.holder.byref(), _p(self.vec[1]), _p(self.vec[0]))

    def calc_surface_fluxes(self, t):
        h = self.holder.byref()
        self.sfv[:] = np.nan
        if self.desc.mesh_kind == 2:
            self.lib.oracle_prolong2interfaces_p4est(h, _p(self.interfaces_u), _p(self.vec[0]))
            self.lib.oracle_calc_interface_flux_p4est(h, _p(self.sfv), _p(self.interfaces_u))
            if self.desc.nboundaries > 0:
                self.lib.oracle_calc_boundary_flux_p4est(h, _p(self.sfv), _p(self.vec[0]), C.c_double(t))
            return
        if self.desc.mesh_kind == 1:
            self.lib.oracle_prolong2interfaces_structured(h, _p(self.interfaces_u), _p(self.vec[0]))
            self.lib.oracle_calc_interface_flux_structured(h, _p(self.sfv), _p(self.interfaces_u))
            if self.desc.nboundaries > 0:
                self.lib.oracle_calc_boundary_flux_structured(h, _p(self.sfv), _p(self.interfaces_u), C.c_double(t))
            return
        self.lib.oracle_prolong2interfaces(h, _p(self.interfaces_u), _p(self.vec[0]))
        self.lib.oracle_calc_interface_flux(h, _p(self.sfv), _p(self.interfaces_u))
        if self.desc.nboundaries > 0:
            self.lib.oracle_prolong2boundaries(h, _p(self.boundaries_u), _p(self.vec[0]))
            self.lib.oracle_calc_boundary_flux(h, _p(self.sfv), _p(self.boundaries_u), C.c_double(t))
        self.lib.oracle_calc_mortar_flux(h, _p(self.sfv), _p(self.vec[0]))

    def download_surface_flux_values(self, host):
        host[:] = self.sfv
        return host

    def numflux(self, flux_id, u_ll, u_rr, orientation):
        u_ll = np.ascontiguousarray(u_ll, dtype=np.float64)
        u_rr = np.ascontiguousarray(u_rr, dtype=np.float64)
        f = np.empty_like(u_ll)
        self.lib.oracle_numflux(self.holder.byref(), C.c_int(flux_id), _p(u_ll), _p(u_rr), C.c_int(orientation), _p(f))
        return f

    def flux(self, u, orientation):
        u = np.ascontiguousarray(u, dtype=np.float64)
        f = np.empty_like(u)
        self.lib.oracle_flux(self.holder.byref(), _p(u), C.c_int(orientation), _p(f))
        return f
