"""Which frequencies of the config-4 modal sweep need many iterations, against their distance to the nearest room mode."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd
grid = fd.shoebox_tet4(99, 133, 74)
s = fd.System(0)
s.set_volume_mesh(grid.nos, grid.elem_vol); s.assemble_volume(1.21, 343.0)
s.modal_compute(float(np.hypot(500.0, 300.0)), f_acc=500.0)
lam, _ = s.modal_get(vectors=False)
res = s.modal_info()["residuals"]
fm = np.sqrt(np.maximum(lam, 0)) / 2 / np.pi
src = grid.closest_node([1.0, 1.0, 1.2]); rcv = grid.closest_node([2.0, 3.0, 1.25])
rn = np.array([[rcv] * 4], dtype=np.int32); rw = np.array([[1.0, 0, 0, 0]])
f = np.arange(20.0, 501.0, 1.0)
_, _, st = s.sweep(2 * np.pi * f, None, [src], [1e-4], rn, rw, want_pN=False, tol=1e-8, precond="modal", check_every=16)
print(st)
k = np.array([np.argmin(np.abs(fm - x)) for x in f])
d = np.abs(fm[k] - f)
o = np.argsort(-st.iters)
print("slowest 25: f, iters, |f - nearest mode| Hz, that mode's rel residual")
for i in o[:25]:
    print(f"  {f[i]:6.1f} {st.iters[i]:5d} {d[i]:8.4f} {res[k[i]] / max(lam[k[i]], 1.0):.2e}")
print("median iters for |df| < 0.05:", np.median(st.iters[d < 0.05]) if (d < 0.05).any() else None, "count", int((d < 0.05).sum()))
print("median iters for |df| > 0.2:", np.median(st.iters[d > 0.2]))
print("corr(iters, -log d):", np.corrcoef(st.iters, -np.log(d + 1e-6))[0, 1])
