import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd
g1 = fd.shoebox_tet4(26, 34, 22)
s = fd.System(0)
s.set_volume_mesh(g1.nos, g1.elem_vol); s.assemble_volume(1.21, 343.0)
s.set_surface_mesh(g1.elem_surf, g1.domain_index_surf, np.arange(2, 8)); s.assemble_surface()
src = g1.closest_node([1.0, 1.0, 1.2])
nf = 64
freq = np.linspace(20, 500, nf)
for name, muval, fc in (("rigid-forced-complex", 0.0, True), ("admittance", 0.02 / (343 * 1.21), False)):
    mu = np.zeros((6, nf), dtype=complex); mu[5] = muval
    for batch in (64,):
        _, pR, st = s.sweep(2 * np.pi * freq, mu, [src], [1e-4], np.array([[src] * 4]), np.array([[1.0, 0, 0, 0]]), want_pN=False, tol=1e-8, batch=batch, check_every=64, max_iter=20000, force_complex=fc)
        bad = np.where(~(st.relres < 1e-7))[0]
        print(f"{name} batch={batch}: real={st.real_arithmetic} unconv {st.n_unconverged} bad idx {bad[:12].tolist()} iters {st.iters[bad][:6].tolist()} relres {st.relres[bad][:4].tolist()}", flush=True)
