"""femder_b200 - B200-native (sm_100a) implementation of femder's ``FEM3D.compute()`` hot path.

Public names mirror the reference (``femder/__init__.py:7-21``) for the path this package covers: ``FEM3D(...).compute()`` /
``.evaluate(R)``, ``p2SPL``.  The problem-setup objects (``AirProperties, AlgControls, BC, Source, Receiver``) are the reference's own;
``room_setup.py`` at the repository root holds minimal stand-ins for machines without the reference.
All numerical work runs in ``libfemder_b200.so`` (hand-written CUDA behind a C ABI, ``include/femder_b200.h``);
there is no CPU fallback.
"""
from . import fem_numerical, mesh, sharding  # noqa: F401
from .fem3d import FEM3D, closest_node, p2SPL, receiver_stencil  # noqa: F401
from .mesh import Grid, read_msh41, shoebox_tet4, to_tet10  # noqa: F401
from .system import System  # noqa: F401

__version__ = "0.1.0"
