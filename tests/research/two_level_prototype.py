"""Research note (CPU, numpy/scipy; uses the oracle, so it lives under tests/): what a coarse space over the tiles would add
to the block-Jacobi preconditioner of the sweep.  Additive two-level Schwarz: M^-1 = sum_t R_t^T B_t R_t + P Gc(w)^-1 P^T
with P = piecewise constant per 128-row tile (one coarse unknown per tile) and Gc = P^T (H - w^2 Q) P solved directly.

    python tests/research/two_level_prototype.py [nx ny nz] [f1,f2,...]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests/research")
from femder_b200.mesh import shoebox_tet4  # noqa: E402
from oracle import femder_oracle as O  # noqa: E402

nx, ny, nz = [int(t) for t in sys.argv[1:4]] if len(sys.argv) >= 4 else (40, 53, 33)
freqs = [float(t) for t in (sys.argv[4].split(",") if len(sys.argv) >= 5 else "50,100,200".split(","))]
TILE = 128
g = shoebox_tet4(nx, ny, nz, jitter=0.0)
c0, rho0 = 343.0, 1.21
H, Q = O.assemble_tet4(g.elem_vol, g.nos, c0, rho0)
H = H.tocsr().real.astype(np.float64)
Q = Q.tocsr().real.astype(np.float64)
N = g.NumNosC
h = 3.0 / nx
src = g.closest_node([1.0, 1.0, 1.2])


def blob_order(indptr, indices, n, tile):
    order = np.empty(n, dtype=np.int64)
    mark = -np.ones(n, dtype=np.int64)
    done = np.zeros(n, dtype=bool)
    frontier, fpos, next_seed, cnt, blob = [], 0, 0, 0, 0
    while cnt < n:
        seed = -1
        while fpos < len(frontier):
            c = frontier[fpos]; fpos += 1
            if not done[c]:
                seed = c; break
        if seed < 0:
            while done[next_seed]:
                next_seed += 1
            seed = next_seed
        q = [seed]; mark[seed] = blob; qh = 0; got = 0
        want = tile - cnt % tile
        while got < want and qh < len(q):
            u = q[qh]; qh += 1
            order[cnt] = u; cnt += 1; done[u] = True; got += 1
            for v in indices[indptr[u]:indptr[u + 1]]:
                if not done[v] and mark[v] != blob:
                    mark[v] = blob; q.append(v)
        frontier.extend(q[qh:])
        blob += 1
    return order


order = blob_order(H.indptr, H.indices, N, TILE)
Pm = sp.csr_matrix((np.ones(N), (np.arange(N), order)), shape=(N, N))
Hp = (Pm @ H @ Pm.T).tocsr()
Qp = (Pm @ Q @ Pm.T).tocsr()
srcp = int(np.where(order == src)[0][0])
n_tiles = (N + TILE - 1) // TILE
tile_of = np.arange(N) // TILE
P = sp.csr_matrix((np.ones(N), (np.arange(N), tile_of)), shape=(N, n_tiles))
Hc = (P.T @ Hp @ P).tocsc()
Qc = (P.T @ Qp @ P).tocsc()
shift = (2 * np.pi * 50.0) ** 2
Bs = []
for t in range(n_tiles):
    a, b = t * TILE, min(N, (t + 1) * TILE)
    Bs.append(np.linalg.inv((Hp[a:b, a:b] + shift * Qp[a:b, a:b]).toarray()).astype(np.float32))


def apply_blocks(r):
    z = np.empty_like(r)
    for t, B in enumerate(Bs):
        a, b = t * TILE, min(N, (t + 1) * TILE)
        z[a:b] = B @ r[a:b].astype(np.float32)
    return z


def cocg(G, b, prec, tol=1e-8, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rho = r @ z
    bb = b @ b
    for it in range(1, maxit + 1):
        q = G @ p
        alpha = rho / (p @ q)
        x += alpha * p; r -= alpha * q
        if r @ r <= tol * tol * bb:
            return x, it
        z = prec(r); rho_new = r @ z
        p = z + (rho_new / rho) * p; rho = rho_new
    return x, maxit


print(f"N={N} h={h:.4f} tiles={n_tiles} (coarse unknowns), tile diameter ~{(TILE ** (1 / 3)) * h:.3f} m")
print(f"{'f':>6} {'k h':>6} {'k H_tile':>9} {'jacobi':>8} {'block':>8} {'block+coarse':>13}")
for f in freqs:
    w = 2 * np.pi * f
    G = (Hp - w * w * Qp).tocsr()
    b = np.zeros(N); b[srcp] = -w * 1e-4
    d = 1.0 / G.diagonal()
    _, it_j = cocg(G, b, lambda r: d * r)
    _, it_b = cocg(G, b, apply_blocks)
    lu = splu((Hc - w * w * Qc).tocsc())
    _, it_2 = cocg(G, b, lambda r: apply_blocks(r) + P @ lu.solve(P.T @ r))
    print(f"{f:6.0f} {w / c0 * h:6.3f} {w / c0 * (TILE ** (1 / 3)) * h:9.3f} {it_j:8d} {it_b:8d} {it_2:13d}", flush=True)
