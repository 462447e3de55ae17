"""Config-1-size sweep (21 735 DOF, 481 frequencies) through FEM3D.compute for several batch widths."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd
import room_setup as rs  # noqa: E402
rigid = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g1 = fd.shoebox_tet4(26, 34, 22)
AP = rs.AirProperties()
AC = rs.AlgControls(AP, 20, 500, 1)
BC = rs.BC(AC, AP)
for t in range(2, 8):
    BC.rigid(t)
if not rigid:
    BC.normalized_admittance(7, 0.02)
S = rs.Source(coord=g1.nos[g1.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
R = rs.Receiver(coord=g1.nos[g1.closest_node([2.0, 3.0, 1.25])])
ref = None
for batch in ([16, 32, 64, "auto"] if rigid else [8, 16, 32, 64, "auto"]):
    o = fd.FEM3D(g1, S, R, AP, AC, BC, tol=1e-8, batch=batch, check_every=64)
    o.compute(timeit=False, keep_pN=False)
    o.H = None
    t = time.perf_counter(); o.compute(timeit=False, keep_pN=False); t = time.perf_counter() - t
    if ref is None: ref = o.pR.copy()
    print(f"rigid={rigid} batch={batch}: {len(AC.freq)/t:7.1f} freq/s wall {t:.2f}s sweep {o.stats.seconds:.2f}s iters mean {o.iters.mean():.0f} unconv {o.stats.n_unconverged} real {o.stats.real_arithmetic} maxdiff_pR {np.abs(o.pR-ref).max()/np.abs(ref).max():.1e}", flush=True)
