"""Golden fixtures for BASELINE config 3 (control-room geometry, mixed admittance boundary conditions, several receivers) on the
reference repository's own unstructured gmsh meshes - SURVEY.md 8d names them as the stand-ins for ``room_treat.geo``, which
cannot be meshed here (no gmsh).  Same procedure as make_golden.py: the UNMODIFIED reference through ``oracle/reference_shim.py``.

    NUMBA_DISABLE_JIT=1 python tests/golden/make_golden_config3.py

* ``Mshs/FEM_3D/caixa.msh``: 3 324 nodes, 15 407 Tet4, 3 244 Tri3 in groups 2, 3 and untagged faces (group 0); the file is in
  millimetres, scaled to metres here as ``GridImport3D(scale=...)`` does (``grid_importing.py:118-120``).
* ``current_mesh.msh``: 3 395 nodes, 15 066 Tet4, 3 698 Tri3, one group.
"""
import os
import sys

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

from femder_b200.mesh import read_msh41  # noqa: E402
from oracle.reference_shim import REFERENCE_ROOT  # noqa: E402
sys.path.insert(0, HERE)
from make_golden import C0, RHO0, run_case  # noqa: E402


def main():
    freq = np.array([20.0, 47.5, 95.0, 160.0, 230.0, 300.0])
    ones = np.ones(len(freq))
    rc = RHO0 * C0

    g = read_msh41(os.path.join(REFERENCE_ROOT, "Mshs/FEM_3D/caixa.msh"))
    g.nos = g.nos * 1e-3
    # frequency-dependent absorber on group 2, complex admittance on group 3, nearly rigid untagged faces
    mu = {0: ones * 0.005 / rc, 2: (0.05 + 0.45 * freq / 300.0) / rc, 3: ones * (0.2 - 0.1j) / rc}
    c = g.nos.mean(axis=0)
    on = [g.nos[g.closest_node(c + [0.9, -0.6, 0.3])], g.nos[g.closest_node(c + [-1.1, 0.7, -0.4])], g.nos[g.closest_node([0.4, 0.4, 0.4])]]
    off = [c + [0.317, 0.211, -0.153], [1.234, 2.345, 1.456], [3.21, 0.87, 0.65]]
    run_case("tet4_caixa", g, freq, mu, [g.nos[g.closest_node([0.7, 0.8, 1.1])]], on, rcv_off=off)

    g = read_msh41(os.path.join(REFERENCE_ROOT, "current_mesh.msh"))
    mu = {2: (0.02 + 0.2 * (freq / 300.0) ** 2) * (1.0 + 0.3j) / rc}
    c = g.nos.mean(axis=0)
    on = [g.nos[g.closest_node(c + [0.8, 0.9, 0.2])], g.nos[g.closest_node(c + [-0.9, -1.2, -0.5])]]
    off = [c + [0.123, -0.456, 0.321], c + [-1.01, 1.37, 0.6]]
    run_case("tet4_current_mesh", g, freq, mu, [g.nos[g.closest_node(c + [0.0, -1.5, 0.1])], g.nos[g.closest_node(c + [0.5, 1.5, -0.3])]], on,
             rcv_off=off)


if __name__ == "__main__":
    main()
