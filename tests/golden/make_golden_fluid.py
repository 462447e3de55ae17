"""Golden fixture for the equivalent-fluid branch of ``FEM3D.compute()`` (``FEM_3D.py:1367-1421``): a jittered Tet4 shoebox whose
lowest 0.7 m is a porous layer (volume domain 2, complex c and rho), an absorbing wall, rigid elsewhere.  The UNMODIFIED reference
runs through ``oracle/reference_shim.py``; the one extra accommodation this branch needs is the NumPy-1 dtype alias ``'cfloat'``
(``int_tetra_4gauss``, ``FEM_3D.py:775-776``), registered in ``np.sctypeDict`` by the shim.

    NUMBA_DISABLE_JIT=1 python tests/golden/make_golden_fluid.py

Kept in its own directory: its H and Q depend on the frequency, so it does not fit the schema of the fixtures next door."""
import os
import sys

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

from femder_b200.mesh import Grid, shoebox_tet4  # noqa: E402
from oracle.reference_shim import run_reference_compute  # noqa: E402

C0, RHO0 = 343.0, 1.21


def main():
    g0 = shoebox_tet4(7, 9, 6, jitter=0.2)
    cent = g0.nos[np.asarray(g0.elem_vol)].mean(axis=1)
    dom = np.where(cent[:, 2] < 0.7, 2, 1).astype(np.int64)
    g = Grid(g0.nos, g0.elem_vol, g0.elem_surf, g0.domain_index_surf, domain_index_vol=dom)
    freq = np.array([30.0, 63.0, 125.0, 200.0])
    mu = {t: np.zeros(len(freq)) for t in range(2, 8)}
    mu[7] = np.ones(len(freq)) * 0.05 / (C0 * RHO0)
    cc, rhoc = 180.0 - 35.0j, 2.1 - 0.9j
    src, rcv = [[1.0, 1.0, 1.2]], [g.nos[g.closest_node([2.0, 3.0, 1.25])]]
    obj = run_reference_compute(g, freq, mu, src, [1e-4], rcv, C0, RHO0, fluid={2: (cc, rhoc)})
    out = dict(nos=g.nos, elem_vol=np.asarray(g.elem_vol, dtype=np.int32), elem_surf=np.asarray(g.elem_surf, dtype=np.int32),
               domain_index_surf=np.asarray(g.domain_index_surf, dtype=np.int64), domain_index_vol=dom, freq=freq, c0=C0, rho0=RHO0,
               group_ids=np.arange(2, 8), mu=np.array([mu[t] for t in range(2, 8)], dtype=np.complex128),
               fluid_domain=np.int64(2), cc=np.complex128(cc), rhoc=np.complex128(rhoc), src=np.atleast_2d(src), src_q=np.array([1e-4]),
               rcv=np.atleast_2d(rcv), pN=obj.pN, pR=obj.pR, q=np.asarray(obj.q.todense()).ravel())
    np.savez_compressed(os.path.join(HERE, "fluid", "tet4_fluid.npz"), **out)
    print(f"tet4_fluid: N={g.NumNosC} E={g.NumElemC} fluid elements {int((dom == 2).sum())} pN={obj.pN.shape} H.dtype={obj.H.dtype}")


if __name__ == "__main__":
    main()
