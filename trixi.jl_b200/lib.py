"""ctypes binding of ``libtrixi_b200.so`` (the C ABI of ``include/trixi_b200.h``).

This is the Python twin of ``julia/TrixiB200.jl``: every call is a thin forward, no arithmetic
happens on this side.  The library must be present and a CUDA device must be usable -- there is no
CPU fallback (BASELINE.json north_star); failures raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrixi_b200.so")

EXPORTS = [
    "trixi_b200_create", "trixi_b200_destroy", "trixi_b200_last_error", "trixi_b200_abi_version",
    "trixi_b200_upload", "trixi_b200_download", "trixi_b200_device_ptr", "trixi_b200_synchronize",
    "trixi_b200_stream", "trixi_b200_rhs_host", "trixi_b200_rhs", "trixi_b200_max_dt",
    "trixi_b200_step_2n", "trixi_b200_step_2n_host", "trixi_b200_step_3sstar", "trixi_b200_step_ssp",
    "trixi_b200_solve_2n", "trixi_b200_set_eq_param", "trixi_b200_calc_error_norms", "trixi_b200_integrate",
    "trixi_b200_calc_volume_integral", "trixi_b200_calc_surface_fluxes",
    "trixi_b200_download_surface_flux_values", "trixi_b200_calc_indicator", "trixi_b200_comm_info_size", "trixi_b200_comm_info",
    "trixi_b200_comm_connect",
    "trixi_b200_launch_count", "trixi_b200_last_elapsed_ms", "trixi_b200_profile_enable",
    "trixi_b200_profile_read", "trixi_b200_timer_start", "trixi_b200_timer_stop",
    "trixi_b200_measure_fp64_peak", "trixi_b200_measure_copy_bandwidth", "trixi_b200_set_option",
]

_lib = None


class TrixiB200Error(RuntimeError):
    pass


def load_library(path=None):
    """dlopen the shared library and declare the prototypes.  Raises if it is missing: the CUDA
    extension is the product, not an optional accelerator."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise TrixiB200Error(f"{path} not found -- build it with `python -m __graft_entry__` / "
                             "trixi.jl_b200/build.py; there is no CPU fallback")
    lib = C.CDLL(path)
    vp, dp, i64p = C.c_void_p, _abi.c_double_p, _abi.c_int64_p
    lib.trixi_b200_create.argtypes = [C.POINTER(_abi.Desc), C.POINTER(vp)]
    lib.trixi_b200_create.restype = C.c_int
    lib.trixi_b200_destroy.argtypes = [vp]
    lib.trixi_b200_destroy.restype = None
    lib.trixi_b200_last_error.argtypes = [vp]
    lib.trixi_b200_last_error.restype = C.c_char_p
    lib.trixi_b200_abi_version.argtypes = []
    lib.trixi_b200_abi_version.restype = C.c_int
    lib.trixi_b200_upload.argtypes = [vp, C.c_int, dp]
    lib.trixi_b200_download.argtypes = [vp, C.c_int, dp]
    lib.trixi_b200_device_ptr.argtypes = [vp, C.c_int]
    lib.trixi_b200_device_ptr.restype = vp
    lib.trixi_b200_synchronize.argtypes = [vp]
    lib.trixi_b200_stream.argtypes = [vp]
    lib.trixi_b200_stream.restype = vp
    lib.trixi_b200_rhs_host.argtypes = [vp, dp, dp, C.c_double]
    lib.trixi_b200_rhs.argtypes = [vp, C.c_double]
    lib.trixi_b200_max_dt.argtypes = [vp, C.c_double, dp]
    lib.trixi_b200_step_2n.argtypes = [vp, C.c_double, C.c_double, dp, dp, dp, C.c_int]
    lib.trixi_b200_step_2n_host.argtypes = [vp, dp, C.c_double, C.c_double, dp, dp, dp, C.c_int]
    lib.trixi_b200_step_3sstar.argtypes = [vp, C.c_double, C.c_double, dp, dp, dp, dp, dp, dp, C.c_int]
    lib.trixi_b200_step_ssp.argtypes = [vp, C.c_double, C.c_double, dp, dp, dp, dp, C.c_int]
    lib.trixi_b200_solve_2n.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int64, dp, dp, dp,
                                        C.c_int, i64p, dp, dp]
    lib.trixi_b200_set_eq_param.argtypes = [vp, C.c_int, C.c_double]
    lib.trixi_b200_calc_error_norms.argtypes = [vp, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp, dp]
    lib.trixi_b200_integrate.argtypes = [vp, C.c_int, dp, dp]
    lib.trixi_b200_calc_volume_integral.argtypes = [vp]
    lib.trixi_b200_calc_surface_fluxes.argtypes = [vp, C.c_double]
    lib.trixi_b200_download_surface_flux_values.argtypes = [vp, dp]
    lib.trixi_b200_calc_indicator.argtypes = [vp, dp]
    lib.trixi_b200_comm_info_size.argtypes = []
    lib.trixi_b200_comm_info_size.restype = C.c_int64
    lib.trixi_b200_comm_info.argtypes = [vp, vp]
    lib.trixi_b200_comm_connect.argtypes = [vp, vp, C.c_int]
    lib.trixi_b200_launch_count.argtypes = [vp]
    lib.trixi_b200_launch_count.restype = C.c_int64
    lib.trixi_b200_last_elapsed_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.trixi_b200_profile_enable.argtypes = [vp, C.c_int]
    lib.trixi_b200_profile_read.argtypes = [vp, C.c_int, dp, i64p]
    lib.trixi_b200_timer_start.argtypes = [vp]
    lib.trixi_b200_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.trixi_b200_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    lib.trixi_b200_measure_fp64_peak.argtypes = [vp, dp]
    lib.trixi_b200_measure_copy_bandwidth.argtypes = [vp, dp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("trixi_b200_abi_version",):
            pass
    if lib.trixi_b200_abi_version() != _abi.ABI_VERSION:
        raise TrixiB200Error("libtrixi_b200.so ABI version mismatch")
    if path == LIB_PATH:
        _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(_abi.c_double_p)


def _check_host(a, length, writable=False):
    if not isinstance(a, np.ndarray) or a.dtype != np.float64:
        raise TypeError("expected a float64 NumPy array")
    if a.size != length:
        raise ValueError(f"array has {a.size} entries, expected {length}")
    if not (a.flags.f_contiguous or a.flags.c_contiguous):
        raise ValueError("array must be contiguous")
    if writable and not a.flags.writeable:
        raise ValueError("array must be writeable")
    return a


class B200Backend:
    """One handle = one GPU's share of the semidiscretization (device-resident u, du, u_tmp and all
    containers).  Method names follow the C exports."""

    U, DU, U_TMP = 0, 1, 2

    def __init__(self, desc_holder, u_length):
        self.lib = load_library()
        self._holder = desc_holder
        self.u_length = int(u_length)
        h = C.c_void_p()
        rc = self.lib.trixi_b200_create(desc_holder.byref(), C.byref(h))
        if rc != 0:
            msg = self.lib.trixi_b200_last_error(None)
            raise TrixiB200Error(f"trixi_b200_create failed ({rc}): {msg.decode() if msg else ''}")
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.trixi_b200_last_error(self.h)
            raise TrixiB200Error(f"libtrixi_b200 error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.trixi_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # data movement
    def upload(self, which, host):
        self._ck(self.lib.trixi_b200_upload(self.h, which, _dptr(_check_host(host, self.u_length))))

    def download(self, which, host=None):
        if host is None:
            host = np.empty(self.u_length)
        self._ck(self.lib.trixi_b200_download(self.h, which, _dptr(_check_host(host, self.u_length, True))))
        return host

    def synchronize(self):
        self._ck(self.lib.trixi_b200_synchronize(self.h))

    def device_ptr(self, which):
        return self.lib.trixi_b200_device_ptr(self.h, which)

    # hot path
    def rhs_host(self, du_host, u_host, t):
        self._ck(self.lib.trixi_b200_rhs_host(self.h, _dptr(_check_host(du_host, self.u_length, True)),
                                              _dptr(_check_host(u_host, self.u_length)), float(t)))

    def rhs(self, t):
        self._ck(self.lib.trixi_b200_rhs(self.h, float(t)))

    def max_dt(self, t=0.0):
        out = C.c_double()
        self._ck(self.lib.trixi_b200_max_dt(self.h, float(t), C.byref(out)))
        return out.value

    def step_2n(self, t, dt, a, b, c):
        a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (a, b, c))
        self._ck(self.lib.trixi_b200_step_2n(self.h, float(t), float(dt), _dptr(a), _dptr(b), _dptr(c), len(c)))

    def step_2n_host(self, u_host, t, dt, a, b, c):
        """One 2N Runge-Kutta step on a host-resident ``u`` (updated in place)."""
        a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (a, b, c))
        self._ck(self.lib.trixi_b200_step_2n_host(self.h, _dptr(_check_host(u_host, self.u_length, True)), float(t),
                                                  float(dt), _dptr(a), _dptr(b), _dptr(c), len(c)))

    def step_3sstar(self, t, dt, gamma1, gamma2, gamma3, beta, delta, c):
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (gamma1, gamma2, gamma3, beta, delta, c)]
        self._ck(self.lib.trixi_b200_step_3sstar(self.h, float(t), float(dt), *[_dptr(x) for x in arrs], len(c)))

    def step_ssp(self, t, dt, numerator_a, numerator_b, denominator, c):
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (numerator_a, numerator_b, denominator, c)]
        self._ck(self.lib.trixi_b200_step_ssp(self.h, float(t), float(dt), *[_dptr(x) for x in arrs], len(c)))

    def solve_2n(self, t0, t_end, cfl, max_steps, a, b, c):
        a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (a, b, c))
        steps, t_out, dt_out = C.c_int64(), C.c_double(), C.c_double()
        self._ck(self.lib.trixi_b200_solve_2n(self.h, float(t0), float(t_end), float(cfl), int(max_steps),
                                              _dptr(a), _dptr(b), _dptr(c), len(c), C.byref(steps),
                                              C.byref(t_out), C.byref(dt_out)))
        return steps.value, t_out.value, dt_out.value

    def calc_error_norms(self, t, ic_id, vandermonde, weights, nvars):
        """(sum of squared errors per variable, Linf per variable, quadrature volume) of the resident u."""
        V = np.ascontiguousarray(np.asarray(vandermonde, dtype=np.float64).ravel(order="F"))
        w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
        l2, linf, vol = np.empty(nvars), np.empty(nvars), np.empty(1)
        self._ck(self.lib.trixi_b200_calc_error_norms(self.h, float(t), int(ic_id), int(w.shape[0]), _dptr(V), _dptr(w),
                                                      _dptr(l2), _dptr(linf), _dptr(vol)))
        return l2, linf, float(vol[0])

    def integrate(self, quantity, nvars):
        """integrate_via_indices of one of the registered integrands over the resident u (and du): (quadrature sums,
        quadrature volume) of this rank, not normalised."""
        out, vol = np.zeros(nvars), np.empty(1)
        self._ck(self.lib.trixi_b200_integrate(self.h, int(quantity), _dptr(out), _dptr(vol)))
        return (out if quantity == 0 else out[:1]), float(vol[0])

    def set_eq_param(self, index, value):
        self._ck(self.lib.trixi_b200_set_eq_param(self.h, int(index), float(value)))

    # stage-level
    def calc_volume_integral(self):
        self._ck(self.lib.trixi_b200_calc_volume_integral(self.h))

    def calc_surface_fluxes(self, t):
        self._ck(self.lib.trixi_b200_calc_surface_fluxes(self.h, float(t)))

    def calc_indicator(self):
        """IndicatorHennemannGassner blending factors [nelements] of the resident u."""
        alpha = np.empty(int(self._holder.desc.nelements))
        self._ck(self.lib.trixi_b200_calc_indicator(self.h, _dptr(alpha)))
        return alpha

    def download_surface_flux_values(self, host):
        self._ck(self.lib.trixi_b200_download_surface_flux_values(self.h, _dptr(host)))
        return host

    # measurement
    def launch_count(self):
        return int(self.lib.trixi_b200_launch_count(self.h))

    def last_elapsed_ms(self):
        out = C.c_float()
        self._ck(self.lib.trixi_b200_last_elapsed_ms(self.h, C.byref(out)))
        return out.value

    def profile_enable(self, on=True):
        self._ck(self.lib.trixi_b200_profile_enable(self.h, int(on)))

    def profile_read(self, kernel_class):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.lib.trixi_b200_profile_read(self.h, kernel_class, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    OPT_KERNEL_PATH = 0
    OPT_FUSED_CFL = 1
    OPT_PREFETCH_DISTANCE = 2
    OPT_HOST_PIPELINE_CHUNK = 3
    OPT_RK_REDUCE_UPDATE = 4
    OPT_SINGLE_FACE_FLUX = 5
    OPT_L2_HINTS = 6
    OPT_FUSED_STAGE = 7

    def set_option(self, option, value):
        self._ck(self.lib.trixi_b200_set_option(self.h, int(option), int(value)))

    def timer_start(self):
        self._ck(self.lib.trixi_b200_timer_start(self.h))

    def timer_stop(self):
        out = C.c_float()
        self._ck(self.lib.trixi_b200_timer_stop(self.h, C.byref(out)))
        return out.value

    def measure_fp64_peak(self):
        out = C.c_double()
        self._ck(self.lib.trixi_b200_measure_fp64_peak(self.h, C.byref(out)))
        return out.value

    def measure_copy_bandwidth(self):
        out = C.c_double()
        self._ck(self.lib.trixi_b200_measure_copy_bandwidth(self.h, C.byref(out)))
        return out.value

    # distributed
    def comm_info(self):
        """Opaque connection blob of this rank (to be all-gathered by the host process group)."""
        buf = C.create_string_buffer(int(self.lib.trixi_b200_comm_info_size()))
        self._ck(self.lib.trixi_b200_comm_info(self.h, buf))
        return buf.raw

    def comm_connect(self, blobs):
        """``blobs``: list of every rank's ``comm_info()`` in rank order."""
        joined = b"".join(blobs)
        buf = C.create_string_buffer(joined, len(joined))
        self._ck(self.lib.trixi_b200_comm_connect(self.h, buf, len(blobs)))

    def connect(self, dist):
        """Wire the halo exchange through a ``torch.distributed`` process group (plumbing only)."""
        blobs = [None] * dist.get_world_size()
        dist.all_gather_object(blobs, self.comm_info())
        self.comm_connect(blobs)
        dist.barrier()
