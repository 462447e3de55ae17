"""Generate the golden fixtures in this directory by executing the UNMODIFIED reference
(``/root/reference/femder``) through ``oracle/reference_shim.py``.

Run in the build container only (the reference is not on the GPU box)::

    NUMBA_DISABLE_JIT=1 python tests/golden/make_golden.py

``NUMBA_DISABLE_JIT=1`` is required for the Tet10 routines and for off-node ``evaluate`` (they relied on
numba's removed object-mode fallback).  Each ``.npz`` stores the mesh arrays the case
was run on, the reference's own ``H, Q, A[i]`` (data / indices / indptr, dtype as the reference produced
them), ``q``, ``areas``, ``pN`` and ``pR``.  The solver behind ``pN`` is SuperLU standing in for PARDISO.
"""
import os
import sys

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

from femder_b200.mesh import read_msh41, shoebox_tet4, to_tet10  # noqa: E402
from oracle.reference_shim import REFERENCE_ROOT, run_reference_compute  # noqa: E402

C0, RHO0 = 343.0, 1.21


def csc_fields(prefix, M, out):
    M = M.tocsc()
    out[prefix + "_data"] = M.data
    out[prefix + "_indices"] = M.indices
    out[prefix + "_indptr"] = M.indptr


def run_case(name, grid, freq, mu, src, rcv, rcv_off=None, v=None):
    freq = np.asarray(freq, dtype=np.float64)
    obj = run_reference_compute(grid, freq, mu, src, [1e-4] * len(np.atleast_2d(src)), rcv, C0, RHO0, v_by_group=v)
    extra = {}
    if rcv_off is not None:
        # the reference's evaluate() cannot mix on-node and off-node receivers in one call
        # (FEM_3D.py:2083-2086 stacks (n_f,) and (n_f,1) arrays), so off-node ones get their own call
        import contextlib, io
        from oracle.reference_shim import PlainReceiver
        pR_on = obj.pR.copy()
        with contextlib.redirect_stdout(io.StringIO()):
            extra["pR_off"] = obj.evaluate(PlainReceiver(rcv_off)).copy()
        extra["rcv_off"] = np.atleast_2d(rcv_off)
        obj.pR = pR_on
    out = dict(nos=grid.nos, elem_vol=np.asarray(grid.elem_vol, dtype=np.int32),
               elem_surf=np.asarray(grid.elem_surf, dtype=np.int32),
               domain_index_surf=np.asarray(grid.domain_index_surf, dtype=np.int64),
               order=np.int64(grid.order), freq=freq, c0=C0, rho0=RHO0,
               src=np.atleast_2d(src), rcv=np.atleast_2d(rcv), src_q=np.full(len(np.atleast_2d(src)), 1e-4),
               group_ids=np.sort([*mu]).astype(np.int64),
               mu=np.array([np.asarray(mu[g], dtype=np.complex128) for g in np.sort([*mu])]),
               areas=np.asarray(obj.areas), q=np.asarray(obj.q.todense()).ravel(),
               pN=obj.pN, pR=obj.pR, **extra)
    if v is not None:
        out["v_ids"] = np.sort([*v]).astype(np.int64)
        out["v"] = np.array([np.asarray(v[g], dtype=np.complex128) for g in np.sort([*v])])
    csc_fields("H", obj.H, out)
    csc_fields("Q", obj.Q, out)
    for i, A in enumerate(obj.A):
        csc_fields(f"A{i}", A, out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: N={grid.NumNosC} E={grid.NumElemC} T={len(grid.elem_surf)} H.dtype={obj.H.dtype} "
          f"A.dtype={obj.A[0].dtype} pN={obj.pN.shape} pR={obj.pR.shape}")


def main():
    nf = 4
    freq = [25.0, 60.0, 110.0, 180.0]
    ones = np.ones(nf)

    # 1. jittered Tet4 shoebox: rigid walls + real admittance wall + complex admittance wall,
    #    receivers: one mesh node, one off-node (Tet4 barycentric interpolation)
    g = shoebox_tet4(6, 8, 5, jitter=0.2)
    mu = {t: np.zeros(nf) for t in range(2, 8)}
    mu[7] = ones * 0.02 / (C0 * RHO0)
    mu[3] = ones * (0.3 + 0.1j) / (C0 * RHO0)
    rnode = g.nos[g.closest_node([2.0, 3.0, 1.25])]
    run_case("tet4_shoebox_jitter", g, freq, mu, [[1.0, 1.0, 1.2]], [rnode],
             rcv_off=[[1.93, 2.71, 1.33], [0.41, 3.6, 0.77]])

    # 2. regular (un-jittered) Tet4 shoebox, all rigid: exact zeros appear in the H pattern
    g = shoebox_tet4(5, 6, 4)
    mu = {t: np.zeros(nf) for t in range(2, 8)}
    run_case("tet4_shoebox_rigid", g, freq, mu, [[1.2, 1.3, 1.25]], [g.nos[g.closest_node([2.0, 3.0, 1.25])]])

    # 3. real unstructured gmsh Tet4 mesh from the reference repository
    g = read_msh41(os.path.join(REFERENCE_ROOT, "Mshs/FEM_3D/cplx_room.msh"))
    mu = {2: ones * 0.05 / (C0 * RHO0), 3: ones * (0.5 - 0.2j) / (C0 * RHO0)}
    c = g.nos.mean(axis=0)
    run_case("tet4_cplx_room", g, freq, mu, [g.nos[g.closest_node(c + 0.3)]],
             [g.nos[g.closest_node(c - 0.4)], g.nos[g.closest_node(c + [0.5, -0.3, 0.1])]])

    # 4. Tet10 shoebox with a normalized-admittance wall (config 2 in miniature)
    g = to_tet10(shoebox_tet4(3, 4, 2, jitter=0.15))
    mu = {t: np.zeros(nf) for t in range(2, 8)}
    mu[7] = ones * 0.02 / (C0 * RHO0)
    run_case("tet10_shoebox", g, freq, mu, [g.nos[g.closest_node([1.0, 1.0, 1.2])]],
             [g.nos[g.closest_node([2.0, 3.0, 1.25])]])

    # 4b. velocity boundary conditions (FEM_3D.py:1320-1347): two driven walls sharing an edge, one absorbing wall
    g = shoebox_tet4(5, 6, 4, jitter=0.2)
    mu = {t: np.zeros(nf) for t in range(2, 8)}
    mu[7] = ones * 0.05 / (C0 * RHO0)
    v = {2: np.array([1e-3, 2e-3 + 1e-3j, 5e-4, -1e-3j]), 4: np.array([5e-4, -1e-3, 2e-3j, 1e-3])}
    run_case("tet4_velocity", g, freq, mu, [[1.0, 1.0, 1.2]], [g.nos[g.closest_node([2.0, 3.0, 1.25])]], v=v)

    # 5. real gmsh Tet10 mesh from the reference repository (Mshs/FEM_3D/bixa1_mesh.msh is not usable:
    #    its element 25 is a blank line)
    for fname, name in (("testmesh.msh", "tet10_testmesh"),):
        g = read_msh41(os.path.join(REFERENCE_ROOT, fname))
        ids = np.unique(g.domain_index_surf)
        mu = {int(t): ones * (0.1 * (i + 1)) / (C0 * RHO0) for i, t in enumerate(ids)}
        c = g.nos.mean(axis=0)
        run_case(name, g, freq, mu, [g.nos[g.closest_node(c + 0.05)]], [g.nos[g.closest_node(c - 0.05)]])


if __name__ == "__main__":
    main()
