"""Debug aid: convergence history of fdb_modal_compute (FDB_MODAL_VERBOSE=1)."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import femder_b200 as fd
C0, RHO0 = 343.0, 1.21
shape = [int(t) for t in sys.argv[1:4]] if len(sys.argv) >= 4 else (18, 24, 15)
f_acc = float(sys.argv[4]) if len(sys.argv) >= 5 else 200.0
delta = float(sys.argv[5]) if len(sys.argv) >= 6 else 300.0
g = fd.shoebox_tet4(*shape, jitter=0.2)
s = fd.System(0)
s.set_volume_mesh(g.nos, g.elem_vol)
s.assemble_volume(RHO0, C0)
f_cut = np.sqrt(f_acc ** 2 + delta ** 2)
x = f_cut / C0
hint = int(4 * np.pi / 3 * 30.0 * x ** 3 + np.pi / 4 * 59.0 * x ** 2 + 38.0 / 8 * x) + 1
t = time.time()
m = s.modal_compute(f_cut, n_hint=hint, f_acc=f_acc)
s.synchronize()
print(f"N={g.NumNosC} f_cut={f_cut:.1f} hint={hint} -> {m} modes in {time.time() - t:.2f}s", {k: v for k, v in s.modal_info().items() if k != "residuals"})
