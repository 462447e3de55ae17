"""Research note (CPU, numpy/scipy; uses the oracle, so it lives under tests/): block-Jacobi preconditioner for the COCG sweep
with the blocks = the 128-row tiles of the sweep's blob order, iteration counts against point Jacobi.

    python tests/research/block_jacobi_prototype.py [nx ny nz]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, "/root/repo")
from femder_b200.mesh import shoebox_tet4  # noqa: E402
from oracle import femder_oracle as O  # noqa: E402

nx, ny, nz = [int(t) for t in sys.argv[1:4]] if len(sys.argv) >= 4 else (40, 53, 33)
TILE = int(sys.argv[4]) if len(sys.argv) >= 5 else 128
g = shoebox_tet4(nx, ny, nz, jitter=0.0)
c0, rho0 = 343.0, 1.21
H, Q = O.assemble_tet4(g.elem_vol, g.nos, c0, rho0)
H = H.tocsr().real.astype(np.float64)
Q = Q.tocsr().real.astype(np.float64)
N = g.NumNosC
h = 3.0 / nx
src = g.closest_node([1.0, 1.0, 1.2])
print(f"N={N} h={h:.4f} tile={TILE}")


def blob_order(indptr, indices, n, tile):
    """Same breadth-first blobs as femder_b200/csrc/reorder.cu (restated)."""
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


t0 = time.time()
order = blob_order(H.indptr, H.indices, N, TILE)
print(f"blob order {time.time() - t0:.1f}s")
Pm = sp.csr_matrix((np.ones(N), (np.arange(N), order)), shape=(N, N))   # new <- old
Hp = (Pm @ H @ Pm.T).tocsr()
Qp = (Pm @ Q @ Pm.T).tocsr()
srcp = int(np.where(order == src)[0][0])
n_tiles = (N + TILE - 1) // TILE


def block_inverses(M, dtype=np.float64):
    out = []
    for t in range(n_tiles):
        a, b = t * TILE, min(N, (t + 1) * TILE)
        blk = M[a:b, a:b].toarray()
        out.append(np.linalg.inv(blk).astype(dtype))
    return out


def apply_blocks(Bs, r):
    z = np.empty_like(r)
    for t, B in enumerate(Bs):
        a, b = t * TILE, min(N, (t + 1) * TILE)
        z[a:b] = (B @ r[a:b].astype(B.dtype)).astype(r.dtype)
    return z


def cocg(G, b, prec, tol=1e-8, maxit=60000):
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


Bs_H = block_inverses(Hp)                       # frequency-independent, SPD
Bs_H32 = [B.astype(np.float32) for B in Bs_H]
print(f"{'f':>6} {'kh':>6} {'jacobi':>8} {'blk exact':>10} {'blk H^-1':>9} {'blk H^-1 f32':>13} {'blk mid-shift':>14}")
fmid = float(sys.argv[5]) if len(sys.argv) >= 6 else None
for f in [float(t) for t in (sys.argv[6].split(",") if len(sys.argv) >= 7 else "40,80,120,160,200".split(","))]:
    w = 2 * np.pi * f
    G = (Hp - w * w * Qp).tocsr()
    b = np.zeros(N); b[srcp] = -w * 1e-4
    d = 1.0 / G.diagonal()
    _, it_j = cocg(G, b, lambda r: d * r)
    Bs_ex = block_inverses(G)
    _, it_e = cocg(G, b, lambda r: apply_blocks(Bs_ex, r))
    _, it_h = cocg(G, b, lambda r: apply_blocks(Bs_H, r))
    _, it_h32 = cocg(G, b, lambda r: apply_blocks(Bs_H32, r))
    it_m = -1
    if fmid:
        wm = 2 * np.pi * fmid
        Bs_m = block_inverses((Hp - wm * wm * Qp).tocsr())
        _, it_m = cocg(G, b, lambda r: apply_blocks(Bs_m, r))
    print(f"{f:6.0f} {w / c0 * h:6.3f} {it_j:8d} {it_e:10d} {it_h:9d} {it_h32:13d} {it_m:14d}", flush=True)
