"""ctypes mirror of ``include/trixi_b200.h`` (struct trixi_b200_desc) and descriptor assembly.

The descriptor is pure data; both the product library (``lib.py``) and the test oracle consume it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 5

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("ndims", C.c_int32), ("nvars", C.c_int32), ("nnodes", C.c_int32), ("mesh_kind", C.c_int32),
        ("nelements", C.c_int64),
        ("equation", C.c_int32),
        ("volume_integral", C.c_int32), ("volume_flux", C.c_int32), ("surface_flux", C.c_int32),
        ("source_terms", C.c_int32),
        ("boundary_conditions", C.c_int32 * 6),
        ("boundary_ic", C.c_int32 * 6),
        ("reserved0", C.c_int32),
        ("eq_params", C.c_double * 8),
        ("derivative_split", c_double_p), ("derivative_hat", c_double_p), ("inverse_weights", c_double_p),
        ("inverse_jacobian", c_double_p), ("node_coordinates", c_double_p),
        ("contravariant_vectors", c_double_p),
        ("ninterfaces", C.c_int64),
        ("interface_neighbor_ids", c_int64_p), ("interface_orientations", c_int64_p),
        ("interface_node_indices", c_int64_p),
        ("nboundaries", C.c_int64),
        ("boundary_neighbor_ids", c_int64_p), ("boundary_orientations", c_int64_p),
        ("boundary_neighbor_sides", c_int64_p), ("boundary_node_coordinates", c_double_p),
        ("n_boundaries_per_direction", C.c_int64 * 6),
        ("nmortars", C.c_int64),
        ("mortar_neighbor_ids", c_int64_p), ("mortar_large_sides", c_int64_p),
        ("mortar_orientations", c_int64_p),
        ("mortar_forward_upper", c_double_p), ("mortar_forward_lower", c_double_p),
        ("mortar_reverse_upper", c_double_p), ("mortar_reverse_lower", c_double_p),
        ("left_neighbors", c_int64_p),
        ("rank", C.c_int32), ("world_size", C.c_int32),
        ("nmpiinterfaces", C.c_int64),
        ("mpi_local_neighbor_ids", c_int64_p), ("mpi_local_sides", c_int64_p),
        ("mpi_orientations", c_int64_p), ("mpi_neighbor_ranks", c_int64_p),
        ("boundary_node_indices", c_int64_p),
        ("mpi_node_indices", c_int64_p),
        ("volume_flux_fv", C.c_int32), ("indicator_variable", C.c_int32),
        ("indicator_alpha_smooth", C.c_int32), ("reserved1", C.c_int32),
        ("indicator_alpha_max", C.c_double), ("indicator_alpha_min", C.c_double),
        ("inverse_vandermonde_legendre", c_double_p),
        ("mortar_node_indices", c_int64_p),
        ("nmpimortars", C.c_int64),
        ("mpi_mortar_neighbor_ids", c_int64_p), ("mpi_mortar_large_sides", c_int64_p),
        ("mpi_mortar_orientations", c_int64_p), ("mpi_mortar_node_indices", c_int64_p),
        ("mpi_mortar_normal_directions", c_double_p), ("mpi_is_mortar_piece", c_int64_p),
        ("subcell_normal_vectors", c_double_p * 3),
    ]


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).ravel(order="F"))


class DescHolder:
    """Owns the NumPy buffers a ``Desc`` points to (GC.@preserve analogue)."""

    def __init__(self):
        self.desc = Desc()
        self._keep = []

    def set_f64(self, name, arr):
        if arr is None:
            setattr(self.desc, name, None)
            return
        a = _f64(arr)
        self._keep.append(a)
        setattr(self.desc, name, a.ctypes.data_as(c_double_p))

    def set_i64(self, name, arr):
        if arr is None:
            setattr(self.desc, name, None)
            return
        a = _i64(arr)
        self._keep.append(a)
        setattr(self.desc, name, a.ctypes.data_as(c_int64_p))

    def set_f64_item(self, name, index, arr):
        a = _f64(arr)
        self._keep.append(a)
        getattr(self.desc, name)[index] = a.ctypes.data_as(c_double_p)

    def byref(self):
        return C.byref(self.desc)
