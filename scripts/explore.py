"""Exploration on the GPU box: build times, SpMV bandwidth, COCG iteration counts at a given size."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="99,133,74")
ap.add_argument("--freqs", default="20,100,250,500")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--tol", type=float, default=1e-8)
ap.add_argument("--max-iter", type=int, default=40000)
ap.add_argument("--admittance", type=float, default=0.0)
ap.add_argument("--jitter", type=float, default=0.0)
ap.add_argument("--warm", type=int, default=0)
ap.add_argument("--spmv", type=int, default=1)
a = ap.parse_args()

shape = tuple(int(t) for t in a.shape.split(","))
t = time.time(); grid = fd.shoebox_tet4(*shape, jitter=a.jitter); print(f"mesh {shape}: N={grid.NumNosC} E={grid.NumElemC} T={len(grid.elem_surf)} host gen {time.time()-t:.1f}s", flush=True)
s = fd.System(0)
t = time.time(); s.set_volume_mesh(grid.nos, grid.elem_vol); print(f"pattern+maps: nnz={s.nnz} ({s.nnz/grid.NumNosC:.2f}/row) {time.time()-t:.3f}s", flush=True)
t = time.time(); s.assemble_volume(1.21, 343.0); s.synchronize(); print(f"assemble_volume {time.time()-t:.4f}s", flush=True)
t = time.time(); s.set_surface_mesh(grid.elem_surf, grid.domain_index_surf, np.arange(2, 8)); s.assemble_surface(); print(f"surface {time.time()-t:.3f}s", flush=True)
if a.spmv:
    for F in (1, 2, 4, 8, 16):
        ms, nb = s.spmv_bench(F, 20)
        print(f"spmv F={F:2d}: {ms*1e3:8.1f} us  {nb/1e6:8.1f} MB  {nb/ms/1e6:7.1f} GB/s", flush=True)
freq = np.array([float(f) for f in a.freqs.split(",")])
mu = np.zeros((6, len(freq)), dtype=complex)
if a.admittance:
    mu[5] = a.admittance / (343.0 * 1.21)
src = grid.closest_node([1.0, 1.0, 1.2])
rcv = grid.closest_node([2.0, 3.0, 1.25])
t = time.time()
_, pR, st = s.sweep(2 * np.pi * freq, mu, [src], [1e-4], np.array([[rcv] * 4]), np.array([[1.0, 0, 0, 0]]), want_pN=False,
                    tol=a.tol, max_iter=a.max_iter, batch=a.batch, warm_start=bool(a.warm))
print(f"sweep wall {time.time()-t:.2f}s device {st.seconds:.2f}s iters {st.iters.tolist()} relres {st.relres.tolist()} unconverged {st.n_unconverged}", flush=True)
print(f"  spmv launches {st.n_spmv}, {st.seconds/max(1,st.n_spmv)*1e6:.1f} us per iteration (batch {a.batch}), SPL {fd.p2SPL(pR).ravel().tolist()}")
