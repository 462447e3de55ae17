"""Per-kernel times of one solver iteration on the config-4 mesh (the numbers bench.py's roofline block reports), for A/B runs."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="99,133,74")
ap.add_argument("--compress", type=int, default=1)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
grid = fd.shoebox_tet4(*[int(t) for t in a.shape.split(",")])
s = fd.System(0)
s.set_volume_mesh(grid.nos, grid.elem_vol)
s.assemble_volume(1.21, 343.0)
f_cut = float(np.hypot(500.0, 300.0))
t = time.perf_counter()
m = s.modal_compute(f_cut, f_acc=500.0)
print(f"modes {m} in {time.perf_counter() - t:.2f} s")
s.modal_compression(bool(a.compress))
kw = dict(real=True, block_jacobi=True, modal=True)
tot = 0.0
for which, name in ((0, "K5 spmv"), (1, "K6a fused"), (3, "projection"), (5, "coefficients"), (4, "expansion"), (6, "finish z += Phi u"), (2, "K6b")):
    if which == 6 and not a.compress:
        continue
    s.iteration_enqueue(16, 1, -1, **kw)
    s.iteration_enqueue(16, 3, which, **kw)
    s.synchronize()
    t = time.perf_counter()
    nb = s.iteration_enqueue(16, a.reps, which, **kw)
    s.synchronize()
    t = (time.perf_counter() - t) / a.reps
    if which != 3 or a.compress:
        tot += t
    print(f"{name:20s} {t * 1e6:8.1f} us  {nb / 1e6:9.1f} MB  {nb / t / 1e9:7.0f} GB/s")
print(f"iteration (fused path) {tot * 1e6:.1f} us")
