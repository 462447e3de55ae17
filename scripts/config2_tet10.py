"""BASELINE config 2: shoebox with quadratic Tet10 elements (~20k DOF) and a normalized-admittance wall, 20-200 Hz at 1 Hz."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd
import room_setup as rs  # noqa: E402
grid = fd.to_tet10(fd.shoebox_tet4(13, 17, 11, jitter=0.1))
AP = rs.AirProperties()
AC = rs.AlgControls(AP, 20, 200, 1)
BC = rs.BC(AC, AP)
for t in range(2, 7):
    BC.rigid(t)
BC.normalized_admittance(7, 0.02)
S = rs.Source(coord=grid.nos[grid.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
R = rs.Receiver(coord=grid.nos[grid.closest_node([2.0, 3.0, 1.25])])
obj = fd.FEM3D(grid, S, R, AP, AC, BC)
obj.compute(timeit=False)
obj.H = None
t = time.perf_counter(); obj.compute(timeit=False); t = time.perf_counter() - t
obj.evaluate(R)
print(f"config 2: N={grid.NumNosC} E={grid.NumElemC} nnz={obj.H.nnz} ({obj.H.nnz/grid.NumNosC:.1f}/row): {len(AC.freq)} frequencies in {t:.2f} s = {len(AC.freq)/t:.1f} freq/s, "
      f"sweep {obj.stats.seconds:.2f} s, iters mean {obj.iters.mean():.0f} max {obj.iters.max()}, unconverged {obj.stats.n_unconverged}, SPL@100Hz {obj.spl[80,0]:.2f} dB")
if "--check" in sys.argv:
    from oracle import femder_oracle as O
    sel = [0, 60, 120, 180]
    ref = O.compute(grid, AC.freq[sel], {k: v[sel] for k, v in BC.mu.items()}, S.coord, S.q.ravel(), 343.0, 1.21)
    rel = np.linalg.norm(obj.pN[sel] - ref["pN"], axis=1) / np.linalg.norm(ref["pN"], axis=1)
    print("   vs oracle direct solve at 4 frequencies: rel L2", rel.max())
