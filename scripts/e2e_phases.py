"""Where an end-to-end FEM3D.compute() call on the config-4 mesh spends its time (host wall clock, cProfile by cumulative time)."""
import cProfile
import os
import pstats
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
KW = dict(modal_delta=float(os.environ["FDB_SETUP_DELTA"])) if "FDB_SETUP_DELTA" in os.environ else {}
obj = fd.FEM3D(grid, S, R, AP, AC, BC, **KW)
obj.compute(timeit=False, keep_pN=False)          # warm-up: library load, handles, graphs
obj2 = fd.FEM3D(grid, S, R, AP, AC, BC, **KW)
obj2._sys = obj._sys
pr = cProfile.Profile()
t = time.perf_counter()
pr.enable()
obj2.compute(timeit=False, keep_pN=False)
pr.disable()
print(f"compute(): {time.perf_counter() - t:.3f} s; sweep device {obj2.stats.seconds:.3f} s; {obj2.stats}")
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
