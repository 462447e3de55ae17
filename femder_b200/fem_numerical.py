"""Function-level drop-ins for the assembly call sites inside the reference's ``FEM3D.compute()``
(``FEM_3D.py:1267-1290``), same names, argument meaning and return values, running on the B200.

==============================================  =====================================================
reference                                        here
==============================================  =====================================================
``fem_numerical.pre_compute_volume_assemble_vars``  uploads the mesh, builds CSR pattern + maps (K4), K1
``fem_numerical.assemble_volume_matrices``          K1 Tet4 -> ``H, Q`` csc float64
``FEM_3D.assemble_Q_H_4_FAST_2order``               K2 Tet10 -> ``H, Q`` csc float64
``FEM_3D.compute_areas``                            Heron areas (K3), complex128 like the reference
``FEM_3D.assemble_surface_matrices``                K3 Tri3 -> list of csc
``FEM_3D.assemble_A10_3_FAST``                      K3 Tri6 -> list of csc
==============================================  =====================================================

All of them work on ONE device system per GPU (``device_system``), kept for the life of the process: a maintainer calling
them in a loop (the room optimizer re-assembles per geometry) re-uses its allocations, and a call that gets the arrays
the previous call got skips the upload and the pattern build.
"""
from __future__ import annotations

import numpy as np

from .system import System

_SYSTEMS = {}       # device -> System
_LOADED = {}        # device -> (key of the mesh the system holds, the arrays it was uploaded from)


def device_system(device=0):
    """The process-wide device system the function-level drop-ins share on GPU ``device``."""
    s = _SYSTEMS.get(device)
    if s is None:
        s = _SYSTEMS[device] = System(device)
    return s


def _stamp(a):
    """Identity plus a checksum: the optimizer loop edits coordinate arrays in place."""
    return None if a is None else (id(a), np.shape(a), float(np.sum(a)))


def _volume(device, vertices, elem_vol):
    s = device_system(device)
    key = ("volume", _stamp(vertices), _stamp(elem_vol))
    if _LOADED.get(device, (None,))[0] != key:
        s.set_volume_mesh(vertices, elem_vol)
        _LOADED[device] = (key, vertices, elem_vol)     # the arrays are kept alive so their ids cannot be recycled
    return s


def _nodes(device, vertices):
    """The surface routines need the node coordinates only; a volume mesh uploaded with the same array serves as well."""
    s = device_system(device)
    held = _LOADED.get(device, (None,))[0]
    if not (held is not None and held[1] == _stamp(vertices)):
        s.set_nodes(vertices)
        _LOADED[device] = (("nodes", _stamp(vertices), None), vertices, None)
    return s


class DeviceBacked(np.ndarray):
    """A host array that remembers the device system and mesh it was computed from (``pre_compute_volume_assemble_vars``
    hands its ``det_ja`` to ``assemble_volume_matrices``, which then finds the pattern already on the GPU)."""

    def __new__(cls, array, device, vertices, elem_vol):
        obj = np.asarray(array).view(cls)
        obj.gpu, obj.vertices, obj.elem_vol = device, vertices, elem_vol
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self.gpu = getattr(obj, "gpu", None)
            self.vertices = getattr(obj, "vertices", None)
            self.elem_vol = getattr(obj, "elem_vol", None)


def _uniform(d, what):
    vals = np.unique(np.concatenate([np.atleast_1d(np.asarray(v)).ravel() for v in d.values()]))
    if len(vals) != 1 or np.imag(vals[0]) != 0:
        raise NotImplementedError(f"per-domain / per-frequency {what}: the function-level drop-in assembles one fluid; "
                                  "equivalent-fluid domains go through femder_b200.FEM3D.compute()")
    return float(np.real(vals[0]))


def pre_compute_volume_assemble_vars(elem_vol, vertices, order, device=0):
    """``fem_numerical.py:129-136``: ``(det_ja (E,), arg_stiff (4, 4, E))`` with ``arg_stiff = (grad N J^-1)(grad N J^-1)^T det J``.
    Computed by the Tet4 assembly kernel with unit density and sound speed (``H_e = arg_stiff / 6``, ``sum Q_e = det J / 6``,
    ``fem_numerical.py:326-330``); the mesh, pattern and element -> nnz map stay on the device for ``assemble_volume_matrices``."""
    if order != 1:
        raise ValueError("order must be 1")   # the reference's compute() calls this for Tet4 only (FEM_3D.py:1270)
    if np.asarray(elem_vol).shape[1] != 4:
        raise ValueError("connectivity does not match the mesh order")
    s = _volume(device, vertices, elem_vol)
    s.assemble_volume(1.0, 1.0)
    He, Qe = s.element_matrices()
    det_ja = 6.0 * Qe.sum(axis=(0, 1))
    return DeviceBacked(det_ja, device, vertices, elem_vol), 6.0 * He


def assemble_volume_matrices(elem_vol, vertices, fluid_c, fluid_rho, order, domain_indices_vol, frequency_index,
                             det_ja=None, arg_stiff=None, device=None):
    """``fem_numerical.py:267-341``: Tet4 ``H, Q`` for one fluid.  ``det_ja`` from ``pre_compute_volume_assemble_vars`` above names
    the device that already holds the mesh; the reference's own arrays are accepted too (the mesh is uploaded then)."""
    if order == 2:
        raise ValueError("order must be 1")  # fem_numerical.py:305
    if device is None:
        device = det_ja.gpu if isinstance(det_ja, DeviceBacked) and det_ja.gpu is not None else 0
    s = _volume(device, vertices, elem_vol)
    s.assemble_volume(_uniform(fluid_rho, "density"), _uniform(fluid_c, "sound speed"))
    return s.get_HQ()


def assemble_Q_H_4_FAST_2order(NumElemC, NumNosC, elem_vol, nos, c0, rho0, device=0):
    """``FEM_3D.py:398-417``."""
    nos, elem_vol = np.asarray(nos), np.asarray(elem_vol)
    if len(nos) != NumNosC:
        nos = nos[:NumNosC]
    if len(elem_vol) != NumElemC:
        elem_vol = elem_vol[:NumElemC]
    s = _volume(device, nos, elem_vol)
    s.assemble_volume(rho0, c0)
    return s.get_HQ()


def compute_areas(vertices, faces, device=0):
    """``FEM_3D.py:722-744``."""
    s = _nodes(device, vertices)
    s.set_surface_mesh(faces, np.zeros(len(faces), dtype=np.int64), [])
    return s.heron_areas().astype(np.complex128)


def assemble_surface_matrices(elem_surf, vertices, areas, domain_indices_surface, unique_domain_indices, order, device=0):
    """``FEM_3D.py:475-523``."""
    if order == 2:
        return None  # FEM_3D.py:498
    s = _nodes(device, vertices)
    s.set_surface_mesh(elem_surf, domain_indices_surface, unique_domain_indices)
    s.assemble_surface()
    return s.get_A()


def assemble_A10_3_FAST(domain_index_surf, number_ID_faces, NumElemC, NumNosC, elem_surf, nos, c0, rho0, device=0):
    """``FEM_3D.py:541-554``."""
    s = _nodes(device, nos)
    s.set_surface_mesh(elem_surf, domain_index_surf, number_ID_faces)
    s.assemble_surface()
    return s.get_A(drop_zeros=True)
