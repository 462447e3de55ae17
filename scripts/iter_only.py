"""Build the config-4 system on the GPU and launch one kernel of a solver iteration (K5 / K6a / K6b): target for ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="99,133,74")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--which", type=int, default=1)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--jacobi", type=int, default=0)
a = ap.parse_args()
grid = fd.shoebox_tet4(*[int(t) for t in a.shape.split(",")])
s = fd.System(0)
s.set_volume_mesh(grid.nos, grid.elem_vol)
s.assemble_volume(1.21, 343.0)
import time  # noqa: E402
s.iteration_enqueue(a.batch, 1, -1, real=True, block_jacobi=not a.jacobi)
s.iteration_enqueue(a.batch, 3, a.which, real=True, block_jacobi=not a.jacobi)
s.synchronize()
t = time.perf_counter()
nb = s.iteration_enqueue(a.batch, a.reps, a.which, real=True, block_jacobi=not a.jacobi)
s.synchronize()
t = (time.perf_counter() - t) / a.reps
print(f"kernel {a.which} x {a.reps}: {nb / 1e6:.1f} MB algorithmic per launch, {t * 1e6:.1f} us per launch (host clock around {a.reps} launches), {nb / t / 1e9:.0f} GB/s")
