"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy / scipy, FP64) of femder's ``FEM3D.compute()`` path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` leg may
import this module.  The product (``femder_b200``) never does; it fails loudly without its CUDA library.

Every function cites the reference lines it restates (paths relative to the reference root).  Parity is
PINNED: ``tests/golden/*.npz`` were produced by executing the unmodified reference here through
``oracle/reference_shim.py`` (script: ``tests/golden/make_golden.py``) and ``tests/test_oracle_golden.py``
checks this restatement against them.  Two deliberate differences, both documented in SURVEY.md (F1, F2):

* Tet4 H/Q are kept in float64.  The reference stores them in ``complex64`` arrays
  (``fem_numerical.py:307-312``); :func:`assemble_tet4` offers ``compat_complex64=True`` to reproduce that.
* The per-frequency solve uses scipy SuperLU (``splu``) where the reference calls MKL PARDISO through the
  un-vendored, unpinned ``pyMKL`` (``FEM_3D.py:1315-1318``).  Both are direct solvers of the same ``G``.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csc_matrix
from scipy.sparse.linalg import splu

# fem_numerical.py:153-171 / FEM_3D.py:897-901
GAUSS_A = 0.5854101966249685
GAUSS_B = 0.1381966011250105


# --------------------------------------------------------------------------------------------------
# scatter: fem_numerical.py:332-343 (and 347-361), FEM_3D.py:410-416
# --------------------------------------------------------------------------------------------------
def coo_rows_cols(conn, n):
    """Row/col index arrays in the reference's ravel order of an ``(n, n, E)`` element array:
    ``assemble_y[a, b, e] = conn[e, a]`` is passed as the COLUMN and its transpose ``conn[e, b]`` as the ROW."""
    conn = np.asarray(conn, dtype=np.int64)
    E = conn.shape[0]
    y = np.broadcast_to(conn.T[:, None, :], (n, n, E))  # [a, b, e] -> conn[e, a]
    x = np.transpose(y, (1, 0, 2))                      # [a, b, e] -> conn[e, b]
    return x.ravel(), y.ravel()


def scatter_csc(elem_mats, conn, n_nodes):
    """``csc_matrix((data, (rows, cols)))`` exactly as the reference builds it; ``elem_mats`` is (n, n, E)."""
    n = elem_mats.shape[0]
    rows, cols = coo_rows_cols(conn, n)
    return csc_matrix((elem_mats.ravel(), (rows, cols)), shape=[n_nodes, n_nodes])


# --------------------------------------------------------------------------------------------------
# Tet4: fem_numerical.py:129-150 (precompute) and 267-345 (element matrices + scatter)
# --------------------------------------------------------------------------------------------------
def tet4_precompute(elem_vol, nos):
    """``det_ja`` (signed, E) and ``arg_stiff = B^T B det`` (4, 4, E), float64."""
    conn = np.asarray(elem_vol, dtype=np.int64)
    X = np.asarray(nos, dtype=np.float64)[conn]                      # (E, 4, 3)
    GNi = np.array([[-1, 1, 0, 0], [-1, 0, 1, 0], [-1, 0, 0, 1]], dtype=np.float64)
    ja = np.einsum("ij,ejl->eil", GNi, X)                            # rows = edge vectors from node 0
    det_ja = np.linalg.det(ja)
    B = np.einsum("eij,jl->eil", np.linalg.inv(ja), GNi)              # (E, 3, 4)
    arg_stiff = np.einsum("eia,eib->abe", B, B) * det_ja
    return det_ja, arg_stiff


def tet4_element_matrices(elem_vol, nos, c0, rho0):
    """H_e = arg_stiff / (6 rho),  Q_e = S det / (24 rho c^2), S = sum_g N N^T  (fem_numerical.py:314-330)."""
    det_ja, arg_stiff = tet4_precompute(elem_vol, nos)
    ptx = np.array([GAUSS_A, GAUSS_B, GAUSS_B, GAUSS_B])
    pty = np.array([GAUSS_B, GAUSS_A, GAUSS_B, GAUSS_B])
    ptz = np.array([GAUSS_B, GAUSS_B, GAUSS_A, GAUSS_B])
    Ng = np.stack([1 - ptx - pty - ptz, ptx, pty, ptz], axis=1)      # (gauss, node)
    S = np.einsum("ga,gb->ab", Ng, Ng)
    He = (1.0 / rho0) * arg_stiff / 6
    Qe = S[:, :, None] * det_ja * (1.0 / (rho0 * c0 ** 2)) / 24
    return He, Qe


def assemble_tet4(elem_vol, nos, c0, rho0, compat_complex64=False):
    He, Qe = tet4_element_matrices(elem_vol, nos, float(c0), float(rho0))
    if compat_complex64:
        He = He.astype(np.complex64)
        Qe = Qe.astype(np.complex64)
    N = len(nos)
    return scatter_csc(He, elem_vol, N), scatter_csc(Qe, elem_vol, N)


# --------------------------------------------------------------------------------------------------
# Tet10: FEM_3D.py:244-254 (N), 281-291 (grad N), 888-919 (4-point rule), 398-417 (scatter)
# --------------------------------------------------------------------------------------------------
def tet10_shape(qsi):
    t1, t2, t3 = qsi
    t4 = 1 - t1 - t2 - t3
    return np.array([t4 * (2 * t4 - 1), t1 * (2 * t1 - 1), t2 * (2 * t2 - 1), t3 * (2 * t3 - 1),
                     4 * t1 * t4, 4 * t1 * t2, 4 * t2 * t4, 4 * t3 * t4, 4 * t2 * t3, 4 * t3 * t1])


def tet10_grad(qsi):
    t1, t2, t3 = 4 * qsi[0], 4 * qsi[1], 4 * qsi[2]
    d = np.array([[t1 + t2 + t3 - 3, t1 + t2 + t3 - 3, t1 + t2 + t3 - 3], [t1 - 1, 0, 0], [0, t2 - 1, 0],
                  [0, 0, t3 - 1], [4 - t2 - t3 - 2 * t1, -t1, -t1], [t2, t1, 0], [-t2, 4 - 2 * t2 - t3 - t1, -t2],
                  [-t3, -t3, 4 - t2 - 2 * t3 - t1], [0, t3, t2], [t3, 0, t1]])
    return d.T  # (3, 10)


def tet10_element_matrices(elem_vol, nos, c0, rho0):
    conn = np.asarray(elem_vol, dtype=np.int64)
    X = np.asarray(nos, dtype=np.float64)[conn]                      # (E, 10, 3)
    E = len(conn)
    He = np.zeros((E, 10, 10))
    Qe = np.zeros((E, 10, 10))
    pts = [(GAUSS_A, GAUSS_B, GAUSS_B), (GAUSS_B, GAUSS_A, GAUSS_B), (GAUSS_B, GAUSS_B, GAUSS_A),
           (GAUSS_B, GAUSS_B, GAUSS_B)]
    w = 1 / 24
    for q in pts:
        Ni = tet10_shape(q)
        GNi = tet10_grad(q)
        Ja = np.einsum("ij,ejl->eil", GNi, X)
        det = np.linalg.det(Ja)
        B = np.einsum("eij,jl->eil", np.linalg.inv(Ja), GNi)
        He += w * (1 / rho0) * np.einsum("eia,eib->eab", B, B) * det[:, None, None]
        Qe += w * (1 / (rho0 * c0 ** 2)) * np.outer(Ni, Ni)[None] * det[:, None, None]
    return np.transpose(He, (1, 2, 0)), np.transpose(Qe, (1, 2, 0))  # (10, 10, E) like Hez/Qez


def assemble_tet10(elem_vol, nos, c0, rho0):
    He, Qe = tet10_element_matrices(elem_vol, nos, float(c0), float(rho0))
    N = len(nos)
    return scatter_csc(He, elem_vol, N), scatter_csc(Qe, elem_vol, N)


# --------------------------------------------------------------------------------------------------
# Surfaces: FEM_3D.py:722-771 (Heron), 440-523 (Tri3), 541-554 + 1053-1092 (Tri6)
# --------------------------------------------------------------------------------------------------
def heron_areas(nos, elem_surf):
    """Heron's formula on the first three (corner) nodes, same operation order as ``area_normal_ph``."""
    f = np.asarray(elem_surf, dtype=np.int64)
    P = np.asarray(nos, dtype=np.float64)
    p0, p1, p2 = P[f[:, 0]], P[f[:, 1]], P[f[:, 2]]

    def length(u, v):
        return np.sqrt((u[:, 0] - v[:, 0]) ** 2 + (u[:, 1] - v[:, 1]) ** 2 + (u[:, 2] - v[:, 2]) ** 2)

    a, b, c = length(p0, p1), length(p1, p2), length(p2, p0)
    p = (a + b + c) / 2
    with np.errstate(invalid="ignore"):
        return np.abs(np.sqrt(p * (p - a) * (p - b) * (p - c)))


def tri3_unit_matrix():
    """``_damp_elem`` of ``assemble_surface_matrices`` (FEM_3D.py:494-503 with 440-459): a tangled 3x3 tensor
    rule, NOT the consistent mass matrix."""
    ptx = np.array([1 / 6, 1 / 6, 2 / 3])
    pty = np.array([1 / 6, 2 / 3, 1 / 6])
    sf = np.array([np.broadcast_to(ptx[:, None], (3, 3)).T, np.broadcast_to(pty[:, None], (3, 3)),
                   1 - ptx - np.broadcast_to(pty[:, None], (3, 3))]).transpose(2, 0, 1)
    D = np.zeros((3, 3))
    for ix in range(3):
        D += np.dot(sf[ix], sf[ix].T) * (1 / 9)
    return D


def tri6_shape(qsi):
    L0 = -qsi[0] - qsi[1] + 1
    return np.array([L0 * (2 * L0 - 1), qsi[0] * (2 * qsi[0] - 1), qsi[1] * (2 * qsi[1] - 1),
                     4 * qsi[0] * qsi[1], 4 * qsi[1] * L0, 4 * qsi[0] * L0])


def tri6_unit_matrix():
    """Sum over the 3-point rule of weight * N N^T (FEM_3D.py:1069-1090); multiply by the Heron area."""
    ptx = np.array([1 / 6, 1 / 6, 2 / 3])
    pty = np.array([1 / 6, 2 / 3, 1 / 6])
    w = 1 / 6 * 2
    D = np.zeros((6, 6))
    for g in range(3):
        Ni = tri6_shape((ptx[g], pty[g]))
        D = D + w * np.outer(Ni, Ni)
    return D


def assemble_surface(elem_surf, nos, domain_index_surf, group_ids, order):
    """One csc (N, N) float64 per group, in the order of ``group_ids`` (the reference passes
    ``np.sort([*mu])``, FEM_3D.py:1283/1290)."""
    f = np.asarray(elem_surf, dtype=np.int64)
    N = len(nos)
    areas = heron_areas(nos, f)
    out = []
    if order == 1:
        D = tri3_unit_matrix()
        for bl in group_ids:
            idx = np.argwhere(np.asarray(domain_index_surf) == bl).ravel()
            damp = D[:, :, None] * areas[idx][None, None, :]
            out.append(scatter_csc(damp, f[idx], N))
    else:
        for bl in group_ids:
            idx = np.argwhere(np.asarray(domain_index_surf) == bl).ravel()
            # per element: Ae accumulated over gauss points with detJa = area (FEM_3D.py:1078-1090)
            ptx = np.array([1 / 6, 1 / 6, 2 / 3]); pty = np.array([1 / 6, 2 / 3, 1 / 6]); w = 1 / 6 * 2
            Ae = np.zeros((len(idx), 6, 6))
            for g in range(3):
                Ni = tri6_shape((ptx[g], pty[g]))
                Ae = Ae + w * (np.outer(Ni, Ni)[None] * areas[idx][:, None, None])
            A = scatter_csc(np.transpose(Ae, (1, 2, 0)), f[idx], N)
            A.eliminate_zeros()  # the reference goes through a dense N x N scratch, csc_matrix(dense) drops zeros
            out.append(A)
    return out


# --------------------------------------------------------------------------------------------------
# compute(): FEM_3D.py:1242-1319 (branch A) and 1423; evaluate(): 2042-2090 with 173-224; p2SPL: 240-242
# --------------------------------------------------------------------------------------------------
def closest_node(nos, coord):
    d = np.asarray(nos) - np.asarray(coord)
    return int(np.argmin(np.einsum("ij,ij->i", d, d)))


def source_vector(nos, src_coord, src_q):
    q = np.zeros(len(nos), dtype=np.complex128)
    src_coord = np.atleast_2d(src_coord)
    src_q = np.asarray(src_q).ravel()
    for i in range(len(src_coord)):
        q[closest_node(nos, src_coord[i])] = src_q[i]  # assignment, not accumulation (FEM_3D.py:1300)
    return q


def system_matrix(H, Q, A_list, mu_list, w):
    """G = H + j w sum_b mu_b A_b - w^2 Q  (FEM_3D.py:1307-1312)."""
    Ag = 0
    for A, mu in zip(A_list, mu_list):
        Ag = Ag + A * mu
    return (H + 1j * w * Ag - (w ** 2) * Q).tocsc().astype(np.complex128)


def velocity_vector(V_list, v_by_group, n, N):
    """``Vn`` of the velocity-BC branch (FEM_3D.py:1326-1339), restated literally: ``V`` is assigned group by group
    (later groups overwrite shared nodes) and ``Vn += V_i @ V`` after each assignment."""
    V = np.zeros(N, dtype=np.complex128)
    Vn = np.zeros(N, dtype=np.complex128)
    for Vm, bl in zip(V_list, np.sort([*v_by_group])):
        idx = np.unique(Vm.tocoo().col)          # nodes of the group's triangles
        V[idx] = np.asarray(v_by_group[bl]).ravel()[n]
        Vn += Vm @ V
    return Vn


def solve_sweep(H, Q, A_list, mu_by_group, group_ids, freq, q, V_list=None, v_by_group=None):
    """Direct solve per frequency: returns pN (n_f, N) complex128."""
    freq = np.asarray(freq, dtype=np.float64)
    w = 2.0 * np.pi * freq
    pN = np.empty((len(freq), H.shape[0]), dtype=np.complex128)
    for n in range(len(freq)):
        mus = [np.asarray(mu_by_group[g]).ravel()[n] for g in group_ids]
        G = system_matrix(H, Q, A_list, mus, w[n])
        if V_list is not None:
            b = -1j * w[n] * velocity_vector(V_list, v_by_group, n, H.shape[0])   # FEM_3D.py:1341
        else:
            b = -1j * w[n] * q
        pN[n] = splu(G).solve(b)
    return pN


def compute(grid, freq, mu_by_group, src_coord, src_q, c0=343.0, rho0=1.21, compat_complex64=False, v_by_group=None):
    """Branch A of ``FEM3D.compute()`` (and its velocity-BC variant); returns dict(H, Q, A, q, pN, group_ids)."""
    group_ids = np.sort([*mu_by_group])
    if grid.order == 1:
        H, Q = assemble_tet4(grid.elem_vol, grid.nos, c0, rho0, compat_complex64)
    else:
        H, Q = assemble_tet10(grid.elem_vol, grid.nos, c0, rho0)
    A = assemble_surface(grid.elem_surf, grid.nos, grid.domain_index_surf, group_ids, grid.order)
    q = source_vector(grid.nos, src_coord, src_q)
    V = None
    if v_by_group:
        V = assemble_surface(grid.elem_surf, grid.nos, grid.domain_index_surf, np.sort([*v_by_group]), grid.order)
    pN = solve_sweep(H, Q, A, mu_by_group, group_ids, freq, q, V, v_by_group)
    return dict(H=H, Q=Q, A=A, q=q, pN=pN, group_ids=group_ids, V=V)


# --------------------------------------------------------------------------------------------------
# Equivalent-fluid branch: FEM_3D.py:1367-1421 with int_tetra_4gauss (773-805), assemble_Q_H_4_FAST_equifluid (353-374),
# int_tri_impedance_3gauss (953-990) and assemble_A_3_FAST (525-539).  Per frequency H and Q are re-assembled with the
# complex rho(w), c(w) of each element's volume domain.
# --------------------------------------------------------------------------------------------------
def tet4_fluid_geometry(elem_vol, nos, float32_shape=True):
    """K_e = B^T B det / 6 and M_e = sum_g N N^T det / 24 per element (4, 4, E): H_e = K_e / rho, Q_e = M_e / (rho c^2).
    ``float32_shape``: ``int_tetra_4gauss`` builds the shape-function vector ``Ni`` as complex64 (FEM_3D.py:800), so ``N N^T`` is
    evaluated in single precision before it meets the complex128 material factor; False gives the FP64 restatement."""
    det_ja, arg_stiff = tet4_precompute(elem_vol, nos)
    ptx = np.array([GAUSS_A, GAUSS_B, GAUSS_B, GAUSS_B])
    pty = np.array([GAUSS_B, GAUSS_A, GAUSS_B, GAUSS_B])
    ptz = np.array([GAUSS_B, GAUSS_B, GAUSS_A, GAUSS_B])
    Ng = np.stack([1 - ptx - pty - ptz, ptx, pty, ptz], axis=1)      # (gauss, node)
    if float32_shape:
        Ng32 = Ng.astype(np.float32)
        S = np.zeros((4, 4), dtype=np.float64)
        for gp in range(4):
            S = S + np.outer(Ng32[gp], Ng32[gp]).astype(np.float64)   # each N N^T rounded to float32, summed after the cast
        # (the reference adds weight * (1 / (rho c^2)) * (N N^T) * det per Gauss point in complex128: the float32 products are exact
        #  inputs to that sum)
    else:
        S = np.einsum("ga,gb->ab", Ng, Ng)
    return arg_stiff / 6.0, S[:, :, None] * det_ja / 24.0


def assemble_fluid(elem_vol, nos, c_by_domain, rho_by_domain, domain_index_vol, fi, float32_shape=True):
    """``assemble_Q_H_4_FAST_equifluid`` for frequency index ``fi``: H, Q csc complex128."""
    Ke, Me = tet4_fluid_geometry(elem_vol, nos, float32_shape)
    dom = np.asarray(domain_index_vol)
    rho = np.array([complex(np.asarray(rho_by_domain[d]).ravel()[fi]) for d in dom])
    c = np.array([complex(np.asarray(c_by_domain[d]).ravel()[fi]) for d in dom])
    He = Ke * (1.0 / rho)[None, None, :]
    Qe = Me * (1.0 / (rho * c ** 2))[None, None, :]
    N = len(nos)
    return scatter_csc(He, elem_vol, N), scatter_csc(Qe, elem_vol, N)


def tri3_impedance_unit_matrix():
    """``int_tri_impedance_3gauss`` without the area: sum over the 3 x 3 tensor rule of w_i w_j N N^T with
    N = [qsi1, qsi2, 1 - qsi1 - qsi2] (FEM_3D.py:969-988)."""
    ptx = np.array([1 / 6, 1 / 6, 2 / 3])
    pty = np.array([1 / 6, 2 / 3, 1 / 6])
    wt = np.array([1 / 6, 1 / 6, 1 / 6]) * 2
    D = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            Ni = np.array([ptx[i], pty[j], 1 - ptx[i] - pty[j]])
            D = D + wt[i] * wt[j] * np.outer(Ni, Ni)
    return D


def assemble_surface_fluid_branch(elem_surf, nos, domain_index_surf, number_ID_faces):
    """``assemble_A_3_FAST``: one csc per entry of ``number_ID_faces`` (in that order), built through a dense N x N scratch, so
    exact zeros are dropped."""
    f = np.asarray(elem_surf, dtype=np.int64)
    N = len(nos)
    areas = heron_areas(nos, f)
    D = tri3_impedance_unit_matrix()
    out = []
    for bl in number_ID_faces:
        idx = np.argwhere(np.asarray(domain_index_surf) == bl).ravel()
        A = scatter_csc(D[:, :, None] * areas[idx][None, None, :], f[idx], N)
        A.eliminate_zeros()
        out.append(A)
    return out


def compute_fluid(grid, freq, mu_by_group, c_by_domain, rho_by_domain, src_coord, src_q, float32_shape=True):
    """The equivalent-fluid branch of ``FEM3D.compute()`` (FEM_3D.py:1367-1399, Tet4): returns dict(pN, A, q)."""
    freq = np.asarray(freq, dtype=np.float64)
    w = 2.0 * np.pi * freq
    ids = np.asarray(grid.number_ID_faces)
    A = assemble_surface_fluid_branch(grid.elem_surf, grid.nos, grid.domain_index_surf, ids)
    q = source_vector(grid.nos, src_coord, src_q)
    pN = np.empty((len(freq), len(grid.nos)), dtype=np.complex128)
    for n in range(len(freq)):
        H, Q = assemble_fluid(grid.elem_vol, grid.nos, c_by_domain, rho_by_domain, grid.domain_index_vol, n, float32_shape)
        Ag = 0
        for Ai, bl in zip(A, ids):
            Ag = Ag + Ai * np.asarray(mu_by_group[bl]).ravel()[n]
        G = (H + 1j * w[n] * Ag - (w[n] ** 2) * Q).tocsc().astype(np.complex128)
        pN[n] = splu(G).solve(-1j * w[n] * q)
    return dict(pN=pN, A=A, q=q)


def tet4_interpolate(nos, elem_vol, coord, pN):
    """``coord_interpolation`` (FEM_3D.py:173-218): candidate tets contain the closest node; pick by the
    barycentric test of ``which_tetra`` (last hit wins); linear shape functions."""
    nos = np.asarray(nos, dtype=np.float64)
    ev = np.asarray(elem_vol, dtype=np.int64)[:, :4]
    coord = np.asarray(coord, dtype=np.float64)
    cl = closest_node(nos, coord)
    cand = np.where((ev == cl).any(axis=1))[0]
    hit = -1
    for t, e in enumerate(cand):
        o = nos[ev[e, 0]]
        M = np.stack([nos[ev[e, 1]] - o, nos[ev[e, 2]] - o, nos[ev[e, 3]] - o], axis=1)
        lam = np.linalg.solve(M, coord - o)
        if np.all(lam >= 0) and np.all(lam <= 1) and lam.sum() <= 1:
            hit = t
    e = cand[hit]  # hit == -1 reproduces the reference's sentinel indexing of the last candidate
    con = ev[e]
    X = nos[con]
    Ja = np.stack([X[1] - X[0], X[2] - X[0], X[3] - X[0]], axis=1)
    qsi = np.linalg.solve(Ja, coord - X[0])
    Ni = np.array([1 - qsi.sum(), qsi[0], qsi[1], qsi[2]])
    return pN[:, con] @ Ni


def evaluate(nos, elem_vol, pN, rcv_coord, tol=1e-3):
    rcv_coord = np.atleast_2d(rcv_coord)
    out = []
    for i in range(len(rcv_coord)):
        n = closest_node(nos, rcv_coord[i])
        if np.linalg.norm(rcv_coord[i] - nos[n]) < tol:
            out.append(pN[:, n])
        else:
            out.append(tet4_interpolate(nos, elem_vol, rcv_coord[i], pN))
    return np.asarray(out).T.reshape(pN.shape[0], len(rcv_coord))


def p2SPL(p):
    return np.real(10 * np.log10(0.5 * p * np.conj(p) / (2e-5) ** 2))
