"""BASELINE config 5 in the small: the room-optimizer loop (reference: FEM_3D.py:1633-1648, GeometryGenerate.py) -
perturbed room geometries on ONE connectivity, re-assembly + 20-200 Hz sweep each.  With torchrun the geometries
are dealt to the ranks (one process per GPU); each rank reuses its device pattern and only re-uploads coordinates.

    python scripts/config5_optimizer_loop.py [--geometries 64] [--shape 26,34,22]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402
from femder_b200.mesh import scale_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--geometries", type=int, default=64)
ap.add_argument("--shape", default="26,34,22")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = int(os.environ.get("LOCAL_RANK", "0"))
base = fd.shoebox_tet4(*[int(t) for t in a.shape.split(",")])
AP = fd.AirProperties()
AC = fd.AlgControls(AP, 20, 200, 1)
BC = fd.BC(AC, AP)
for t in range(2, 8):
    BC.rigid(t)
rng = np.random.default_rng(0)
factors = rng.uniform(0.95, 1.05, size=(a.geometries, 3))
t0 = time.perf_counter()
nfreq = 0
obj = None
for gi in range(rank, a.geometries, world):
    grid = scale_grid(base, factors[gi])
    S = fd.Source(coord=grid.nos[grid.closest_node(np.array([1.0, 1.0, 1.2]) * factors[gi])], q=[1e-4])
    R = fd.Receiver(coord=grid.nos[grid.closest_node(np.array([2.0, 3.0, 1.25]) * factors[gi])])
    new = fd.FEM3D(grid, S, R, AP, AC, BC, device=dev, shard=False)
    if obj is not None:
        new._sys = obj._sys          # keep the device allocation; the new coordinates force re-assembly
    obj = new
    obj.compute(timeit=False, keep_pN=False)
    nfreq += len(AC.freq)
    assert obj.stats.n_unconverged == 0
dt = time.perf_counter() - t0
print(f"rank {rank}/{world}: {a.geometries // world} geometries x {len(AC.freq)} frequencies in {dt:.2f} s = {nfreq / dt:.1f} frequencies/s "
      f"(N = {base.NumNosC}), last receiver SPL at 100 Hz = {obj.spl[80, 0]:.2f} dB")
