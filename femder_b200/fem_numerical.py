"""Function-level drop-ins for the assembly call sites inside the reference's ``FEM3D.compute()``
(``FEM_3D.py:1267-1290``), same names and argument meaning, running on the B200.

==============================================  =====================================================
reference                                        here
==============================================  =====================================================
``fem_numerical.pre_compute_volume_assemble_vars``  uploads the mesh, builds CSR pattern + maps (K4)
``fem_numerical.assemble_volume_matrices``          K1 Tet4 -> ``H, Q`` csc float64
``FEM_3D.assemble_Q_H_4_FAST_2order``               K2 Tet10 -> ``H, Q`` csc float64
``FEM_3D.compute_areas``                            Heron areas (K3), complex128 like the reference
``FEM_3D.assemble_surface_matrices``                K3 Tri3 -> list of csc
``FEM_3D.assemble_A10_3_FAST``                      K3 Tri6 -> list of csc
==============================================  =====================================================
"""
from __future__ import annotations

import numpy as np

from .system import System


class VolumeSetup:
    """What ``pre_compute_volume_assemble_vars`` hands to ``assemble_volume_matrices``: the reference passes
    ``det_ja`` / ``arg_stiff`` arrays (``fem_numerical.py:129-136``); here the expensive precomputation is the
    pattern and element->nnz map on the device, so the handle carries the device system instead."""

    def __init__(self, system):
        self.system = system


def _uniform(d, what):
    vals = np.unique(np.concatenate([np.atleast_1d(np.real(np.asarray(v))).ravel() for v in d.values()]))
    if len(vals) != 1:
        raise NotImplementedError(f"per-domain / per-frequency {what} (equivalent fluids) is outside the B200 hot path")
    return float(vals[0])


def pre_compute_volume_assemble_vars(elem_vol, vertices, order, device=0):
    elem_vol = np.asarray(elem_vol)
    if elem_vol.shape[1] != (4 if order == 1 else 10):
        raise ValueError("connectivity does not match the mesh order")
    s = System(device)
    s.set_volume_mesh(vertices, elem_vol)
    return VolumeSetup(s), None


def assemble_volume_matrices(elem_vol, vertices, fluid_c, fluid_rho, order, domain_indices_vol, frequency_index,
                             det_ja=None, arg_stiff=None):
    if order == 2:
        raise ValueError("order must be 1")  # fem_numerical.py:305
    setup = det_ja if isinstance(det_ja, VolumeSetup) else pre_compute_volume_assemble_vars(elem_vol, vertices, 1)[0]
    setup.system.assemble_volume(_uniform(fluid_rho, "density"), _uniform(fluid_c, "sound speed"))
    return setup.system.get_HQ()


def assemble_Q_H_4_FAST_2order(NumElemC, NumNosC, elem_vol, nos, c0, rho0, device=0):
    s = System(device)
    s.set_volume_mesh(np.asarray(nos)[:NumNosC], np.asarray(elem_vol)[:NumElemC])
    s.assemble_volume(rho0, c0)
    return s.get_HQ()


def _surface_system(vertices, elem_surf, device):
    """Surface routines only need the node coordinates; a one-tet dummy volume mesh carries them."""
    s = System(device)
    nos = np.ascontiguousarray(vertices, dtype=np.float64)
    nper = 4 if np.asarray(elem_surf).shape[1] == 3 else 10
    s.set_volume_mesh(nos, np.arange(nper, dtype=np.int32).reshape(1, nper) % len(nos))
    return s


def compute_areas(vertices, faces, device=0):
    s = _surface_system(vertices, faces, device)
    s.set_surface_mesh(faces, np.zeros(len(faces), dtype=np.int64), [])
    return s.heron_areas().astype(np.complex128)


def assemble_surface_matrices(elem_surf, vertices, areas, domain_indices_surface, unique_domain_indices, order, device=0):
    if order == 2:
        return None  # FEM_3D.py:498
    s = _surface_system(vertices, elem_surf, device)
    s.set_surface_mesh(elem_surf, domain_indices_surface, unique_domain_indices)
    s.assemble_surface()
    return s.get_A()


def assemble_A10_3_FAST(domain_index_surf, number_ID_faces, NumElemC, NumNosC, elem_surf, nos, c0, rho0, device=0):
    s = _surface_system(nos, elem_surf, device)
    s.set_surface_mesh(elem_surf, domain_index_surf, number_ID_faces)
    s.assemble_surface()
    return s.get_A(drop_zeros=True)
