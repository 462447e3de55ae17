"""End-to-end FEM3D.compute() time on the config-4 mesh against the modal margin modal_delta (set-up grows with it, the sweep shrinks)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402
import room_setup as rs  # noqa: E402

grid = fd.shoebox_tet4(99, 133, 74)
AP = rs.AirProperties(c0=343.0, rho0=1.21)
AC = rs.AlgControls(AP, freq_vec=np.arange(20.0, 501.0, 1.0))
BC = rs.BC(AC, AP)
for t in range(2, 8):
    BC.rigid(t)
S = rs.Source(coord=grid.nos[grid.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
R = rs.Receiver(coord=grid.nos[grid.closest_node([2.0, 3.0, 1.25])])
prev = None
for delta in [300.0] + [float(t) for t in sys.argv[1:]]:
    obj = fd.FEM3D(grid, S, R, AP, AC, BC, modal_delta=delta)
    if prev is not None:
        obj._sys = prev._sys
    t = time.perf_counter()
    obj.compute(timeit=False, keep_pN=False)
    dt = time.perf_counter() - t
    info = obj._sys.modal_info()
    print(f"delta {delta:5.0f}: compute() {dt:.3f} s (sweep {obj.stats.seconds:.3f} s), {obj._sys.modal_count()} modes, set-up outer {info['outer']}, "
          f"iters mean {obj.iters.mean():.0f} max {obj.iters.max()}, resumed {obj.stats.n_resumed}, unconverged {obj.n_unconverged}", flush=True)
    prev = obj
