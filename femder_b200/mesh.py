"""Mesh containers and synthetic / file inputs for the FEM3D hot path.

A :class:`Grid` carries exactly the attributes the reference's ``GridImport3D``
leaves behind for ``FEM3D`` to read (reference ``femder/grid_importing.py:87-329``,
consumed at ``femder/FEM_3D.py:1177-1193``): ``nos``, ``elem_vol``, ``elem_surf``,
``domain_index_surf``, ``domain_index_vol``, ``number_ID_faces``, ``number_ID_vol``,
``NumNosC``, ``NumElemC``, ``order``, ``path_to_geo``, ``path_to_geo_unrolled``.

gmsh and meshio are not available, so the producers here are

* :func:`shoebox_tet4` - the synthetic Kuhn-tetrahedra shoebox of SURVEY.md 8(d)
  (all tets positively oriented, wall tags 2..7 = x0, x1, y0, y1, z0, z1),
* :func:`to_tet10` - mid-edge nodes appended in gmsh-native edge order
  (``FEM_3D.py:244-254``: 4:(0,1) 5:(1,2) 6:(0,2) 7:(0,3) 8:(2,3) 9:(1,3)),
* :func:`read_msh41` - a reader for gmsh 4.1 ASCII ``.msh`` files.
"""
from __future__ import annotations

import itertools

import numpy as np

# edge (a, b) whose midpoint is Tet10 node 4..9 / Tri6 node 3..5, gmsh-native order
TET10_EDGES = ((0, 1), (1, 2), (0, 2), (0, 3), (2, 3), (1, 3))
TRI6_EDGES = ((0, 1), (1, 2), (0, 2))


class Grid:
    """Plain holder with the attribute names ``FEM3D`` reads from ``GridImport3D``."""

    def __init__(self, nos, elem_vol, elem_surf, domain_index_surf, domain_index_vol=None, order=1,
                 path_to_geo=None):
        self.nos = np.ascontiguousarray(nos, dtype=np.float64)
        self.elem_vol = np.ascontiguousarray(elem_vol)
        self.elem_surf = np.ascontiguousarray(elem_surf)
        self.domain_index_surf = np.asarray(domain_index_surf)
        if domain_index_vol is None:
            domain_index_vol = np.ones(len(self.elem_vol), dtype=np.int64)
        self.domain_index_vol = np.asarray(domain_index_vol)
        self.order = int(order)
        self.path_to_geo = path_to_geo
        self.path_to_geo_unrolled = None
        # grid_importing.py:317-323
        self.number_ID_faces = np.unique(self.domain_index_surf)
        self.number_ID_vol = np.unique(self.domain_index_vol)
        self.NumNosC = len(self.nos)
        self.NumElemC = len(self.elem_vol)

    def closest_node(self, coord):
        d = self.nos - np.asarray(coord, dtype=np.float64)
        return int(np.argmin(np.einsum("ij,ij->i", d, d)))


def _kuhn_local_tets():
    """The 6 Kuhn tets of the unit cube as corner offsets, each with det(J) > 0."""
    tets = []
    for perm in itertools.permutations(range(3)):
        v = np.zeros(3, dtype=np.int64)
        corners = [v.copy()]
        for ax in perm:
            v[ax] += 1
            corners.append(v.copy())
        corners = np.array(corners)
        edges = corners[1:] - corners[0]
        if np.linalg.det(edges.astype(np.float64)) < 0:
            corners[[2, 3]] = corners[[3, 2]]
        tets.append(corners)
    return np.array(tets)  # (6, 4, 3)


def shoebox_tet4(nx, ny, nz, L=(3.0, 4.0, 2.5), jitter=0.0, seed=1234, index_dtype=np.uint32):
    """Shoebox ``L`` split into ``nx*ny*nz`` hexahedra, each into 6 Kuhn tets.

    Node id = ``(i*(ny+1) + j)*(nz+1) + k``.  ``jitter`` displaces interior nodes by
    ``U(-jitter*h, jitter*h)`` per axis so the pattern carries no exact zeros.
    """
    L = np.asarray(L, dtype=np.float64)
    gx, gy, gz = nx + 1, ny + 1, nz + 1
    xs = np.linspace(0.0, L[0], gx)
    ys = np.linspace(0.0, L[1], gy)
    zs = np.linspace(0.0, L[2], gz)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    nos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if jitter:
        rng = np.random.default_rng(seed)
        h = L / np.array([nx, ny, nz])
        I, J, K = np.meshgrid(np.arange(gx), np.arange(gy), np.arange(gz), indexing="ij")
        interior = ((I > 0) & (I < nx) & (J > 0) & (J < ny) & (K > 0) & (K < nz)).ravel()
        d = rng.uniform(-jitter, jitter, size=nos.shape) * h
        nos[interior] += d[interior]

    def nid(i, j, k):
        return (i * gy + j) * gz + k

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    local = _kuhn_local_tets()
    elem = np.empty((len(I), 6, 4), dtype=np.int64)
    for t in range(6):
        for a in range(4):
            o = local[t, a]
            elem[:, t, a] = nid(I + o[0], J + o[1], K + o[2])
    elem = elem.reshape(-1, 4)

    # boundary triangles: every wall quad is cut along the Kuhn diagonal (0,0)-(1,1)
    surf, tags = [], []

    def wall(axis, side, tag):
        n = (nx, ny, nz)
        a1, a2 = [a for a in range(3) if a != axis]
        U, V = np.meshgrid(np.arange(n[a1]), np.arange(n[a2]), indexing="ij")
        U, V = U.ravel(), V.ravel()
        fixed = 0 if side == 0 else n[axis]

        def node(du, dv):
            c = [None, None, None]
            c[axis] = np.full_like(U, fixed)
            c[a1] = U + du
            c[a2] = V + dv
            return nid(c[0], c[1], c[2])

        t1 = np.stack([node(0, 0), node(1, 0), node(1, 1)], axis=1)
        t2 = np.stack([node(0, 0), node(1, 1), node(0, 1)], axis=1)
        tri = np.concatenate([t1, t2], axis=0)
        surf.append(tri)
        tags.append(np.full(len(tri), tag, dtype=np.int64))

    for tag, (axis, side) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)], start=2):
        wall(axis, side, tag)
    elem_surf = np.concatenate(surf, axis=0)
    dom_surf = np.concatenate(tags)
    g = Grid(nos, elem.astype(index_dtype), elem_surf.astype(index_dtype), dom_surf,
             np.ones(len(elem), dtype=np.int64), order=1, path_to_geo="synthetic:shoebox")
    g.shape_cells = (nx, ny, nz)
    g.L = tuple(L)
    return g


def to_tet10(grid, index_dtype=np.uint32):
    """Quadratic version of a Tet4 grid: straight edges, midpoints in gmsh-native order."""
    ev = np.asarray(grid.elem_vol, dtype=np.int64)
    es = np.asarray(grid.elem_surf, dtype=np.int64)
    n0 = len(grid.nos)
    pairs = []
    for a, b in TET10_EDGES:
        pairs.append(np.stack([ev[:, a], ev[:, b]], axis=1))
    for a, b in TRI6_EDGES:
        pairs.append(np.stack([es[:, a], es[:, b]], axis=1))
    allp = np.concatenate(pairs, axis=0)
    lo = np.minimum(allp[:, 0], allp[:, 1])
    hi = np.maximum(allp[:, 0], allp[:, 1])
    key = lo * n0 + hi
    ukey, inv = np.unique(key, return_inverse=True)
    mid = 0.5 * (grid.nos[ukey // n0] + grid.nos[ukey % n0])
    nos = np.concatenate([grid.nos, mid], axis=0)
    E, T = len(ev), len(es)
    mids_v = (inv[: 6 * E].reshape(6, E).T + n0)
    mids_s = (inv[6 * E:].reshape(3, T).T + n0)
    ev10 = np.concatenate([ev, mids_v], axis=1)
    es6 = np.concatenate([es, mids_s], axis=1)
    g = Grid(nos, ev10.astype(index_dtype), es6.astype(index_dtype), grid.domain_index_surf,
             grid.domain_index_vol, order=2, path_to_geo=grid.path_to_geo)
    return g


def scale_grid(grid, factors):
    """Same connectivity, node coordinates scaled per axis (config 5: perturbed rooms)."""
    g = Grid(grid.nos * np.asarray(factors, dtype=np.float64), grid.elem_vol, grid.elem_surf,
             grid.domain_index_surf, grid.domain_index_vol, order=grid.order, path_to_geo=grid.path_to_geo)
    return g


_GMSH_NODES_PER = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8, 8: 3, 9: 6, 11: 10, 15: 1}


def read_msh41(path, order=None, scale=1.0, index_dtype=np.uint32):
    """Read a gmsh 4.1 ASCII mesh the way the reference consumes it through meshio
    (``grid_importing.py:279-308``): points renumbered 0-based in file order, triangles and
    tetrahedra with their ``gmsh:physical`` tag (0 for untagged entities)."""
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    pos = 0

    def find(tag):
        nonlocal pos
        while lines[pos].strip() != tag:
            pos += 1
        pos += 1

    find("$MeshFormat")
    ver = lines[pos].split()
    if not ver[0].startswith("4") or int(ver[1]) != 0:
        raise ValueError("read_msh41: only gmsh 4.x ASCII is supported")
    # entity -> first physical tag
    phys = {0: {}, 1: {}, 2: {}, 3: {}}
    pos = 0
    try:
        find("$Entities")
        np_, nc, ns, nv = (int(t) for t in lines[pos].split())
        pos += 1
        for _ in range(np_):
            t = lines[pos].split(); pos += 1
            nph = int(t[4])
            phys[0][int(t[0])] = int(t[5]) if nph else 0
        for dim, cnt in ((1, nc), (2, ns), (3, nv)):
            for _ in range(cnt):
                t = lines[pos].split(); pos += 1
                nph = int(t[7])
                phys[dim][int(t[0])] = int(t[8]) if nph else 0
    except IndexError:
        pos = 0
    find("$Nodes")
    nblocks, nnodes, _, _ = (int(t) for t in lines[pos].split())
    pos += 1
    tags = np.empty(nnodes, dtype=np.int64)
    xyz = np.empty((nnodes, 3), dtype=np.float64)
    at = 0
    for _ in range(nblocks):
        _, _, parametric, nb = (int(t) for t in lines[pos].split())
        pos += 1
        for i in range(nb):
            tags[at + i] = int(lines[pos + i])
        pos += nb
        for i in range(nb):
            xyz[at + i] = [float(t) for t in lines[pos + i].split()[:3]]
        pos += nb
        at += nb
    remap = np.full(tags.max() + 1, -1, dtype=np.int64)
    remap[tags] = np.arange(nnodes)
    find("$Elements")
    nblocks = int(lines[pos].split()[0])
    pos += 1
    cells = {}
    for _ in range(nblocks):
        dim, etag, etype, nb = (int(t) for t in lines[pos].split())
        pos += 1
        nper = _GMSH_NODES_PER.get(etype)
        if nper is None:
            raise ValueError(f"read_msh41: unsupported gmsh element type {etype}")
        blk = np.array([[int(t) for t in lines[pos + i].split()[1:1 + nper]] for i in range(nb)],
                       dtype=np.int64).reshape(nb, nper)
        pos += nb
        c = cells.setdefault(etype, ([], []))
        c[0].append(remap[blk])
        c[1].append(np.full(nb, phys[dim].get(etag, 0), dtype=np.int64))
    have2 = 11 in cells
    if order is None:
        order = 2 if have2 else 1
    vt, st = (11, 9) if order == 2 else (4, 2)
    if vt not in cells or st not in cells:
        raise ValueError(f"read_msh41: no order-{order} tetrahedra/triangles in {path}")
    ev = np.concatenate(cells[vt][0]); dv = np.concatenate(cells[vt][1])
    es = np.concatenate(cells[st][0]); ds = np.concatenate(cells[st][1])
    return Grid(xyz / scale, ev.astype(index_dtype), es.astype(index_dtype), ds, dv, order=order, path_to_geo=path)
