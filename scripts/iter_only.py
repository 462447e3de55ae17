"""Build the config-4 system on the GPU and launch one kernel of a solver iteration: target for ncu.
which: 0 K5 SpMV, 1 K6a (with --modal: k_precond_fused), 2 K6b, 4 modal expansion, 5 modal coefficients, 3 stand-alone projection."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="99,133,74")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--which", type=int, default=1)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--jacobi", type=int, default=0)
ap.add_argument("--modal", type=int, default=1)
a = ap.parse_args()
grid = fd.shoebox_tet4(*[int(t) for t in a.shape.split(",")])
s = fd.System(0)
s.set_volume_mesh(grid.nos, grid.elem_vol)
s.assemble_volume(1.21, 343.0)
kw = dict(real=True, block_jacobi=not a.jacobi, modal=bool(a.modal))
if a.modal:
    f_cut = float(np.hypot(500.0, 300.0))
    x = f_cut / 343.0
    m = s.modal_compute(f_cut, n_hint=int(4 * np.pi / 3 * 30.0 * x ** 3 + np.pi / 4 * 59.0 * x ** 2 + 4.7 * x) + 1, f_acc=500.0)
    print("modes", m)
s.iteration_enqueue(a.batch, 1, -1, **kw)
s.iteration_enqueue(a.batch, 3, a.which, **kw)
s.synchronize()
t = time.perf_counter()
nb = s.iteration_enqueue(a.batch, a.reps, a.which, **kw)
s.synchronize()
t = (time.perf_counter() - t) / a.reps
print(f"kernel {a.which} x {a.reps}: {nb / 1e6:.1f} MB algorithmic per launch, {t * 1e6:.1f} us per launch (host clock around {a.reps} launches), {nb / t / 1e9:.0f} GB/s")
