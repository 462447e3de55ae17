"""-m gpu: the FEM3D drop-in (same calls as the reference's examples/3d_fem_run.py) against the oracle and golden."""
import numpy as np
import pytest

import room_setup as rs
from oracle import femder_oracle as O

pytestmark = pytest.mark.gpu


def _run(g, **kw):
    import femder_b200 as fd
    AP = rs.AirProperties(c0=g.c0, rho0=g.rho0)
    AC = rs.AlgControls(AP, freq_vec=g.freq)
    BC = rs.BC(AC, AP)
    for k, v in g.mu.items():
        BC.mu[k] = v
    for k, v in (g.v or {}).items():
        BC.v[k] = v
    S = rs.Source(coord=g.src, q=g.src_q)
    R = rs.Receiver(coord=g.rcv)
    obj = fd.FEM3D(g.grid, S, R, AP, AC, BC, tol=1e-10, batch=4, **kw)
    obj.compute(timeit=False)
    return obj, R


def test_compute_leaves_reference_state(gpu_lib, golden):
    g = golden
    obj, R = _run(g)
    N = g.N
    assert obj.H.shape == (N, N) and obj.H.format == "csc" and obj.H.indices.dtype == np.int32
    assert np.array_equal(obj.H.indices, g.H.indices) and np.array_equal(obj.H.indptr, g.H.indptr)
    assert isinstance(obj.A, list) and len(obj.A) == len(g.A)
    assert obj.q.shape == (N, 1) and obj.q.format == "csc" and np.array_equal(np.asarray(obj.q.todense()).ravel(), g.q)
    assert obj.areas.dtype == np.complex128 and np.abs(obj.areas - g.areas).max() < 1e-15
    assert obj.pN.shape == g.pN.shape and obj.pN.dtype == np.complex128 and obj.t > 0
    # end to end against the oracle with FP64 assembly (north star: 1e-6 relative L2, 0.01 dB)
    r = O.compute(g.grid, g.freq, g.mu, g.src, g.src_q, g.c0, g.rho0, v_by_group=g.v)
    rel = np.linalg.norm(obj.pN - r["pN"], axis=1) / np.linalg.norm(r["pN"], axis=1)
    assert rel.max() < 1e-6
    if g.v:
        assert len(obj.V) == len(g.v) and all(M.format == "csc" for M in obj.V)
    pR = obj.evaluate(R)
    assert pR.shape == (len(g.freq), len(g.rcv))
    assert np.abs(obj.spl - O.p2SPL(O.evaluate(g.grid.nos, g.grid.elem_vol, r["pN"], g.rcv))).max() < 0.01
    # against the reference's verbatim output: Tet10 to solver tolerance, Tet4 limited by its complex64 matrices
    relg = np.linalg.norm(obj.pN - g.pN, axis=1) / np.linalg.norm(g.pN, axis=1)
    assert relg.max() < (1e-6 if g.order == 2 else 5e-3)


def test_compat_complex64(gpu_lib):
    """compat_complex64 rounds H/Q to float32 like the reference does before solving."""
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    obj, _ = _run(g, compat_complex64=True)
    assert obj.H.dtype == np.complex64
    rel = np.linalg.norm(obj.pN - g.pN, axis=1) / np.linalg.norm(g.pN, axis=1)
    obj64, _ = _run(g)
    rel64 = np.linalg.norm(obj64.pN - g.pN, axis=1) / np.linalg.norm(g.pN, axis=1)
    assert rel.max() < 5e-4 and rel.max() <= rel64.max() * 1.5


def test_receivers_only_and_offnode(gpu_lib):
    import femder_b200 as fd
    from conftest import Golden
    g = Golden("tet4_shoebox_jitter")
    obj, R = _run(g)
    Roff = rs.Receiver(coord=g.rcv_off)
    pRoff = obj.evaluate(Roff)
    assert np.abs(fd.p2SPL(pRoff) - fd.p2SPL(g.pR_off)).max() < 0.01
    obj.R = Roff
    obj.compute(timeit=False, keep_pN=False)
    assert obj.pN is None and np.abs(obj.pR - pRoff).max() / np.abs(pRoff).max() < 1e-8


def test_missing_bc_raises_keyerror(gpu_lib):
    import femder_b200 as fd
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    AP = rs.AirProperties()
    AC = rs.AlgControls(AP, freq_vec=g.freq)
    BC = rs.BC(AC, AP)
    BC.rigid(2)  # group 3 has no boundary condition
    obj = fd.FEM3D(g.grid, rs.Source(coord=g.src, q=g.src_q), None, AP, AC, BC)
    with pytest.raises(KeyError):
        obj.compute(timeit=False)
    BC.rhoc[1] = np.ones(len(g.freq))   # a fluid domain needs both its density and its sound speed (BC.fluid sets both)
    with pytest.raises(ValueError):
        fd.FEM3D(g.grid, None, None, AP, AC, BC)


def test_h_is_cached_between_computes(gpu_lib):
    import femder_b200 as fd
    from conftest import Golden
    g = Golden("tet4_cplx_room")
    obj, _ = _run(g)
    H0 = obj.H
    obj.S = rs.Source(coord=g.grid.nos[10], q=[2e-4])
    obj.compute(timeit=False)
    assert obj.H is H0  # "H is None" is the cache test (FEM_3D.py:1269)
    r = O.compute(g.grid, g.freq, g.mu, g.grid.nos[10:11], [2e-4], g.c0, g.rho0)
    assert (np.linalg.norm(obj.pN - r["pN"], axis=1) / np.linalg.norm(r["pN"], axis=1)).max() < 1e-6


def test_no_boundary_conditions_branch(gpu_lib):
    """BC is None: G = H - w^2 Q (FEM_3D.py:1348-1365)."""
    import femder_b200 as fd
    from conftest import Golden
    g = Golden("tet4_shoebox_rigid")
    AP = rs.AirProperties(c0=g.c0, rho0=g.rho0)
    AC = rs.AlgControls(AP, freq_vec=g.freq)
    S = rs.Source(coord=g.src, q=g.src_q)
    obj = fd.FEM3D(g.grid, S, None, AP, AC, None, tol=1e-10)
    obj.compute(timeit=False)
    assert obj.A is None and obj.stats.real_arithmetic
    r = O.compute(g.grid, g.freq, {}, g.src, g.src_q, g.c0, g.rho0)
    assert (np.linalg.norm(obj.pN - r["pN"], axis=1) / np.linalg.norm(r["pN"], axis=1)).max() < 1e-6


def test_empty_surface_mesh_and_bad_batch(gpu_lib):
    import femder_b200 as fd
    from femder_b200._lib import FemderB200Error
    grid = fd.shoebox_tet4(3, 3, 3, jitter=0.1)
    s = fd.System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.assemble_volume(1.21, 343.0)
    s.set_surface_mesh(np.zeros((0, 3), dtype=np.int32), np.zeros(0, dtype=np.int64), [2, 3])   # groups without triangles
    s.assemble_surface()
    A = s.get_A()
    assert len(A) == 2 and all(a.nnz == 0 for a in A) and len(s.heron_areas()) == 0
    pN, _, st = s.sweep(2 * np.pi * np.array([50.0]), np.zeros((2, 1)), [7], [1e-4], tol=1e-9, batch=4)
    assert st.n_unconverged == 0 and pN.shape == (1, grid.NumNosC)
    with pytest.raises(ValueError):
        s.sweep(2 * np.pi * np.array([50.0]), np.zeros((2, 1)), [7], [1e-4], batch=3)
    with pytest.raises(FemderB200Error, match="source node out of range"):
        s.sweep(2 * np.pi * np.array([50.0]), np.zeros((2, 1)), [grid.NumNosC], [1e-4], batch=4)


def test_gpu_point_location_matches_host_rule(gpu_lib):
    """fdb_locate_points against the host restatement of closest_node / which_tetra / coord_interpolation
    (FEM_3D.py:173-224) on random points, mesh nodes and points outside the mesh (sentinel rule)."""
    import femder_b200 as fd
    from femder_b200 import receiver_stencil
    grid = fd.shoebox_tet4(7, 9, 6, jitter=0.2)
    s = fd.System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    rng = np.random.default_rng(5)
    pts = np.vstack([rng.uniform([0.05, 0.05, 0.05], [2.95, 3.95, 2.45], size=(300, 3)), grid.nos[[0, 17, 200]],
                     grid.nos[5] + 4e-4, [[3.2, 4.1, 2.6], [-0.1, 2.0, 1.0]]])
    nodes, w = s.locate_points(pts, 1e-3)
    assert nodes.shape == (len(pts), 4) and w.shape == (len(pts), 4)
    pN = rng.standard_normal((2, grid.NumNosC)) + 1j * rng.standard_normal((2, grid.NumNosC))
    ref = O.evaluate(grid.nos, grid.elem_vol, pN, pts)
    got = np.einsum("fpj,pj->fp", pN[:, nodes.astype(np.int64)], w)
    assert np.abs(got - ref).max() < 1e-10
    n_same = 0
    for i, c in enumerate(pts):
        hn, hw = receiver_stencil(grid.nos, grid.elem_vol, c)
        n_same += int(np.array_equal(hn, nodes[i]))
        assert abs(w[i].sum() - 1) < 1e-12
    assert n_same >= len(pts) - 3          # a point exactly on a face may legitimately pick the neighbouring element
    # Tet10: corner nodes of the quadratic element
    g10 = fd.to_tet10(grid)
    s10 = fd.System(0)
    s10.set_volume_mesh(g10.nos, g10.elem_vol)
    n10, w10 = s10.locate_points(pts[:50], 1e-3)
    p10 = rng.standard_normal((1, g10.NumNosC))
    for i in range(50):   # same rule on the Tet10 connectivity: the nearest node may be a mid-edge node
        hn, hw = receiver_stencil(g10.nos, g10.elem_vol, pts[i])
        assert abs(p10[0, n10[i].astype(np.int64)] @ w10[i] - p10[0, hn.astype(np.int64)] @ hw) < 1e-10
