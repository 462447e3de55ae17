"""1M-DOF probe of the modal path: set-up time, sweep time, iterations (FDB_MODAL_VERBOSE=1 for the set-up history)."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import femder_b200 as fd
C0, RHO0 = 343.0, 1.21
shape = [int(t) for t in sys.argv[1:4]] if len(sys.argv) >= 4 else (99, 133, 74)
delta = float(sys.argv[4]) if len(sys.argv) >= 5 else 300.0
stride = int(sys.argv[5]) if len(sys.argv) >= 6 else 5
g = fd.shoebox_tet4(*shape)
s = fd.System(0)
t = time.time(); s.set_volume_mesh(g.nos, g.elem_vol); s.assemble_volume(RHO0, C0); s.synchronize()
print(f"N={g.NumNosC} mesh+assembly {time.time() - t:.2f}s", flush=True)
f_hi = 500.0
f_cut = np.sqrt(f_hi ** 2 + delta ** 2)
x = f_cut / C0
hint = int(4 * np.pi / 3 * 30.0 * x ** 3 + np.pi / 4 * 59.0 * x ** 2 + 38.0 / 8 * x) + 1
t = time.time(); m = s.modal_compute(f_cut, n_hint=hint, f_acc=f_hi); s.synchronize(); t_modal = time.time() - t
info = s.modal_info()
print(f"modal_compute: f_cut={f_cut:.1f} hint={hint} -> {m} modes in {t_modal:.2f}s", {k: v for k, v in info.items() if k != "residuals"}, flush=True)
src = g.closest_node([1.0, 1.0, 1.2]); rcv = g.closest_node([2.0, 3.0, 1.25])
rn = np.array([[rcv] * 4], dtype=np.int32); rw = np.array([[1.0, 0, 0, 0]])
f = np.arange(20.0, 501.0, 1.0)[::stride]
for ce in (8, 16):
    t = time.time()
    _, pR, st = s.sweep(2 * np.pi * f, None, [src], [1e-4], rn, rw, want_pN=False, tol=1e-8, precond="modal", check_every=ce, modal_delta=delta, max_iter=4000)
    dt = time.time() - t
    print(f"modal sweep check_every={ce}: {len(f)} freqs in {dt:.3f}s wall ({st.seconds:.3f}s device) = {len(f) / dt:.1f} freq/s", st, flush=True)
print("iters by freq:", st.iters)
if "--block" in sys.argv:
    t = time.time()
    _, pR2, sb = s.sweep(2 * np.pi * f[::4], None, [src], [1e-4], rn, rw, want_pN=False, tol=1e-8, precond="block", batch=16, check_every=64)
    dt = time.time() - t
    print(f"block sweep: {len(f[::4])} freqs in {dt:.3f}s = {len(f[::4]) / dt:.2f} freq/s", sb, flush=True)
    d = np.abs(pR[::4] - pR2) / np.abs(pR2)
    print("max rel diff modal vs block at the receiver:", d.max())
