"""-m gpu: K4 pattern / maps and K1-K3 assembly through the C ABI against the oracle and the golden vectors."""
import numpy as np
import pytest

from oracle import femder_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _system(grid, group_ids=None):
    from femder_b200 import System
    s = System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    if group_ids is not None:
        s.set_surface_mesh(grid.elem_surf, grid.domain_index_surf, group_ids)
    return s


def _oracle_HQ(grid, c0=343.0, rho0=1.21):
    if grid.order == 1:
        return O.assemble_tet4(grid.elem_vol, grid.nos, c0, rho0)
    return O.assemble_tet10(grid.elem_vol, grid.nos, c0, rho0)


def _elem2nnz_numpy(grid, indptr, indices):
    conn = np.asarray(grid.elem_vol, dtype=np.int64)
    E, n = conn.shape
    N = len(indptr) - 1
    key = np.repeat(np.arange(N, dtype=np.int64), np.diff(indptr)) * N + indices
    want = conn[:, :, None] * N + conn[:, None, :]
    pos = np.searchsorted(key, want.ravel())
    assert np.array_equal(key[pos], want.ravel())
    return pos.reshape(E, n, n).astype(np.int32)


def test_pattern_bit_exact_golden(gpu_lib, golden):
    s = _system(golden.grid)
    indptr, indices = s.pattern()
    assert indptr.dtype == np.int32 and indices.dtype == np.int32
    assert np.array_equal(indptr, golden.H.indptr) and np.array_equal(indices, golden.H.indices)
    assert np.array_equal(s.elem2nnz(), _elem2nnz_numpy(golden.grid, indptr, indices))


@pytest.mark.parametrize("shape,jitter,order", [((7, 5, 6), 0.2, 1), ((6, 6, 6), 0.0, 1), ((1, 1, 1), 0.0, 1),
                                                ((4, 3, 5), 0.15, 2), ((26, 34, 22), 0.2, 1)])
def test_pattern_bit_exact_shoebox(gpu_lib, shape, jitter, order):
    from femder_b200.mesh import shoebox_tet4, to_tet10
    grid = shoebox_tet4(*shape, jitter=jitter)
    if order == 2:
        grid = to_tet10(grid)
    s = _system(grid)
    indptr, indices = s.pattern()
    H, _ = _oracle_HQ(grid)
    assert np.array_equal(indptr, H.indptr) and np.array_equal(indices, H.indices)
    assert np.array_equal(s.elem2nnz(), _elem2nnz_numpy(grid, indptr, indices))


def test_volume_values_golden(gpu_lib, golden):
    g = golden
    s = _system(g.grid)
    s.assemble_volume(g.rho0, g.c0)
    H, Q = s.get_HQ()
    Ho, Qo = _oracle_HQ(g.grid, g.c0, g.rho0)
    assert H.dtype == np.float64 and H.indices.dtype == np.int32 and H.has_canonical_format
    # north star: assembled values within 1e-12 relative of the FP64 restatement
    assert _rel(H.data, Ho.data) < 1e-12 and _rel(Q.data, Qo.data) < 1e-12
    # against the reference's own arrays: float32 eps for Tet4 (complex64 storage), 1e-12 for Tet10
    tol = 4e-7 if g.order == 1 else 1e-12   # float32 storage and accumulation in the reference (up to ~30 contributions per entry)
    assert _rel(H.data, g.H.data) < tol and _rel(Q.data, g.Q.data) < tol
    # element matrices before the scatter
    He, Qe = s.element_matrices()
    if g.order == 1:
        Heo, Qeo = O.tet4_element_matrices(g.grid.elem_vol, g.grid.nos, g.c0, g.rho0)
    else:
        Heo, Qeo = O.tet10_element_matrices(g.grid.elem_vol, g.grid.nos, g.c0, g.rho0)
    assert _rel(He, Heo) < 1e-12 and _rel(Qe, Qeo) < 1e-12


def test_volume_deterministic(gpu_lib):
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(9, 7, 8, jitter=0.2)
    vals = []
    for _ in range(3):
        s = _system(grid)
        s.assemble_volume(1.21, 343.0)
        vals.append(s.get_HQ_values())
    for H, Q in vals[1:]:
        assert np.array_equal(H, vals[0][0]) and np.array_equal(Q, vals[0][1])  # bit-reproducible segmented sum


def test_volume_invariants_config1(gpu_lib):
    """Config 1 size (26x34x22 cells, 21 735 DOF): analytic invariants of SURVEY section 4."""
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(26, 34, 22, jitter=0.2)
    c0, rho0 = 343.0, 1.21
    s = _system(grid, np.arange(2, 8))
    s.assemble_volume(rho0, c0)
    H, Q = s.get_HQ()
    assert H.shape == (21735, 21735)
    assert abs(Q.sum() * rho0 * c0 ** 2 - 30.0) < 1e-9
    assert np.abs(H @ np.ones(H.shape[0])).max() < 1e-10
    Ho, Qo = _oracle_HQ(grid, c0, rho0)
    assert _rel(H.data, Ho.data) < 1e-12 and _rel(Q.data, Qo.data) < 1e-12
    s.assemble_surface()
    A = s.get_A()
    wall = [10.0, 10.0, 7.5, 7.5, 12.0, 12.0]
    for a, w in zip(A, wall):
        assert abs(a.sum() - w) < 1e-10


def test_negative_orientation_subtracts(gpu_lib):
    """det(J) is signed: flipping a tet's orientation flips the sign of its Q_e."""
    from femder_b200.mesh import Grid, shoebox_tet4
    grid = shoebox_tet4(1, 1, 1)
    ev = np.asarray(grid.elem_vol).copy()
    ev[:, [1, 2]] = ev[:, [2, 1]]
    flipped = Grid(grid.nos, ev, grid.elem_surf, grid.domain_index_surf)
    s = _system(flipped)
    s.assemble_volume(1.21, 343.0)
    _, Q = s.get_HQ()
    assert abs(Q.sum() * 1.21 * 343.0 ** 2 + 30.0) < 1e-10
    Ho, Qo = _oracle_HQ(flipped)
    assert _rel(Q.data, Qo.data) < 1e-12


def test_heron_areas_bit_exact(gpu_lib, golden):
    s = _system(golden.grid, golden.group_ids)
    a = s.heron_areas()
    ao = O.heron_areas(golden.grid.nos, golden.grid.elem_surf)
    assert np.array_equal(a, ao)                         # same IEEE operations in the same order
    assert np.abs(a - golden.areas.real).max() < 1e-15   # the reference's own values


def test_surface_golden(gpu_lib, golden):
    g = golden
    s = _system(g.grid, g.group_ids)
    s.assemble_surface()
    A = s.get_A(drop_zeros=(g.order == 2))
    assert len(A) == len(g.A)
    for M, R in zip(A, g.A):
        assert M.dtype == np.float64
        assert np.array_equal(M.indptr, R.indptr) and np.array_equal(M.indices, R.indices)
        if R.nnz:
            assert _rel(M.data, R.data) < 1e-12


def test_surface_group_without_bc_is_skipped(gpu_lib):
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(3, 3, 3, jitter=0.1)
    s = _system(grid, np.array([3, 7]))
    s.assemble_surface()
    A = s.get_A()
    Ao = O.assemble_surface(grid.elem_surf, grid.nos, grid.domain_index_surf, [3, 7], 1)
    for M, R in zip(A, Ao):
        assert np.array_equal(M.indices, R.indices) and _rel(M.data, R.data) < 1e-13


def test_function_level_drop_ins(gpu_lib, golden):
    """Same names / arguments as the reference call sites FEM_3D.py:1267-1290."""
    from femder_b200 import fem_numerical as fn
    g = golden
    grid = g.grid
    areas = fn.compute_areas(grid.nos, grid.elem_surf)
    assert areas.dtype == np.complex128 and np.abs(areas - g.areas).max() < 1e-15
    if g.order == 1:
        fluid_c = {1: np.ones(len(grid.elem_vol)) * g.c0}
        fluid_rho = {1: np.ones(len(grid.elem_vol)) * g.rho0}
        det_ja, arg_stiff = fn.pre_compute_volume_assemble_vars(grid.elem_vol, grid.nos, 1)
        # the reference's return values (fem_numerical.py:129-146): det J and (J^-1 grad N)^T (J^-1 grad N) det J per element
        xc = np.asarray(grid.nos)[np.asarray(grid.elem_vol)]
        J = np.stack([xc[:, 1] - xc[:, 0], xc[:, 2] - xc[:, 0], xc[:, 3] - xc[:, 0]], axis=1)
        gradN = np.array([[-1.0, 1, 0, 0], [-1, 0, 1, 0], [-1, 0, 0, 1]])
        B = np.linalg.inv(J) @ gradN
        det = np.abs(np.linalg.det(J))
        ref_stiff = np.einsum("eil,eij->lje", B, B) * det
        assert det_ja.shape == (len(grid.elem_vol),) and _rel(np.asarray(det_ja), det) < 1e-12
        assert arg_stiff.shape == (4, 4, len(grid.elem_vol)) and _rel(arg_stiff, ref_stiff) < 1e-12
        sys_before = fn.device_system(0)
        H, Q = fn.assemble_volume_matrices(grid.elem_vol, grid.nos, fluid_c, fluid_rho, 1, grid.domain_index_vol, 0,
                                           det_ja, arg_stiff)
        A = fn.assemble_surface_matrices(grid.elem_surf, grid.nos, areas, grid.domain_index_surf, g.group_ids, 1)
        assert fn.device_system(0) is sys_before       # one device system serves every call
        with pytest.raises(ValueError):
            fn.assemble_volume_matrices(grid.elem_vol, grid.nos, fluid_c, fluid_rho, 2, grid.domain_index_vol, 0)
        tol = 4e-7
    else:
        H, Q = fn.assemble_Q_H_4_FAST_2order(grid.NumElemC, grid.NumNosC, grid.elem_vol, grid.nos, g.c0, g.rho0)
        A = fn.assemble_A10_3_FAST(grid.domain_index_surf, g.group_ids, grid.NumElemC, grid.NumNosC, grid.elem_surf,
                                   grid.nos, g.c0, g.rho0)
        assert fn.assemble_surface_matrices(grid.elem_surf, grid.nos, areas, grid.domain_index_surf, g.group_ids, 2) is None
        tol = 1e-12
    assert np.array_equal(H.indices, g.H.indices) and _rel(H.data, g.H.data) < tol and _rel(Q.data, g.Q.data) < tol
    for M, R in zip(A, g.A):
        assert np.array_equal(M.indices, R.indices) and _rel(M.data, R.data) < 1e-12


def test_bad_inputs_raise(gpu_lib):
    from femder_b200 import System
    from femder_b200._lib import FemderB200Error
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(2, 2, 2)
    s = System(0)
    ev = np.asarray(grid.elem_vol).astype(np.int32).copy()
    ev[0, 0] = grid.NumNosC + 5
    with pytest.raises(FemderB200Error, match="outside"):
        s.set_volume_mesh(grid.nos, ev)
    with pytest.raises(ValueError):
        s.set_volume_mesh(grid.nos, ev[:, :3])
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    with pytest.raises(FemderB200Error):
        s.assemble_volume(0.0, 343.0)


def test_config4_full_size_properties(gpu_lib):
    """BASELINE config 4 at full size (99x133x74 cells, 1 005 000 DOF): size-independent properties of the pattern, the
    assembled matrices and two solves, checked on the host with numpy / scipy."""
    from scipy.sparse import csr_matrix
    from femder_b200 import System
    from femder_b200.mesh import shoebox_tet4
    grid = shoebox_tet4(99, 133, 74)
    N = grid.NumNosC
    assert N == 1005000 and grid.NumElemC == 5846148
    s = System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    indptr, indices = s.pattern()
    # canonical CSR: int32, sorted and duplicate-free inside every row
    rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(indptr))
    key = rows * N + indices
    assert indptr[0] == 0 and indptr[-1] == len(indices) and np.all(np.diff(key) > 0)
    # exactly the couplings of the connectivity: nnz = N + 2 * (number of mesh edges), pattern symmetric
    ev = np.asarray(grid.elem_vol, dtype=np.int64)
    pairs = np.concatenate([np.minimum(ev[:, a], ev[:, b]) * N + np.maximum(ev[:, a], ev[:, b])
                            for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))])
    edges = np.unique(pairs)
    assert len(indices) == N + 2 * len(edges)
    upper = key[indices > rows]
    assert np.array_equal(upper, edges)                                   # row-major upper triangle == sorted edge keys
    lower = (indices[indices < rows].astype(np.int64) * N + rows[indices < rows])
    assert np.array_equal(np.sort(lower), edges)                          # symmetric
    # assembled values
    c0, rho0 = 343.0, 1.21
    s.assemble_volume(rho0, c0)
    Hv, Qv = s.get_HQ_values()
    H = csr_matrix((Hv, indices, indptr), shape=(N, N))
    Q = csr_matrix((Qv, indices, indptr), shape=(N, N))
    assert abs(Qv.sum() * rho0 * c0 ** 2 - 30.0) < 1e-8                   # sum(Q) rho c^2 = room volume
    assert np.abs(H @ np.ones(N)).max() < 1e-9                            # H 1 = 0
    assert np.abs((H - H.T).data).max() < 1e-12 if (H - H.T).nnz else True
    # two solves verified with scipy: ||b - G x|| / ||b|| <= tol, purely imaginary pressure for a rigid room
    freq = np.array([20.0, 45.0])
    w = 2 * np.pi * freq
    src = grid.closest_node([1.0, 1.0, 1.2])
    pN, _, st = s.sweep(w, np.zeros((0, 2)), [src], [1e-4], tol=1e-8, batch=16)
    assert st.n_unconverged == 0 and st.real_arithmetic and not pN.real.any()
    for n in range(2):
        b = np.zeros(N, dtype=complex); b[src] = -1j * w[n] * 1e-4
        res = b - (H @ pN[n] - w[n] ** 2 * (Q @ pN[n]))
        assert np.linalg.norm(res) / np.linalg.norm(b) < 1.05e-8
