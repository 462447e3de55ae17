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
import room_setup as rs  # noqa: E402
from femder_b200.mesh import scale_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--geometries", type=int, default=64)
ap.add_argument("--shape", default="26,34,22")
ap.add_argument("--candidates", type=int, default=0, help="source / receiver candidate pairs per geometry (FEM_3D.py:1633-1648), batched as "
                                                            "right-hand sides of one sweep with FEM3D.compute_candidates; 0 = one pair, compute()")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = int(os.environ.get("LOCAL_RANK", "0"))
base = fd.shoebox_tet4(*[int(t) for t in a.shape.split(",")])
AP = rs.AirProperties()
AC = rs.AlgControls(AP, 20, 200, 1)
BC = rs.BC(AC, AP)
for t in range(2, 8):
    BC.rigid(t)
rng = np.random.default_rng(0)
factors = rng.uniform(0.95, 1.05, size=(a.geometries, 3))
t0 = time.perf_counter()
nfreq = 0
obj = None
for gi in range(rank, a.geometries, world):
    grid = scale_grid(base, factors[gi])
    S = rs.Source(coord=grid.nos[grid.closest_node(np.array([1.0, 1.0, 1.2]) * factors[gi])], q=[1e-4])
    R = rs.Receiver(coord=grid.nos[grid.closest_node(np.array([2.0, 3.0, 1.25]) * factors[gi])])
    new = fd.FEM3D(grid, S, R, AP, AC, BC, device=dev, shard=False)
    if obj is not None:
        new._sys = obj._sys          # keep the device allocation; the new coordinates force re-assembly
    obj = new
    if a.candidates > 0:
        crng = np.random.default_rng(gi)
        lo, hi = grid.nos.min(axis=0) + 0.3, grid.nos.max(axis=0) - 0.3
        sC = [rs.Source(coord=grid.nos[grid.closest_node(crng.uniform(lo, hi))], q=[1e-4]) for _ in range(a.candidates)]
        rC = [rs.Receiver(coord=crng.uniform(lo, hi)) for _ in range(a.candidates)]
        obj.compute_candidates(sC, rC)
        nfreq += len(AC.freq) * a.candidates
    else:
        obj.compute(timeit=False, keep_pN=False)
        nfreq += len(AC.freq)
    assert obj.n_unconverged == 0
dt = time.perf_counter() - t0
print(f"rank {rank}/{world}: {a.geometries // world} geometries x {max(1, a.candidates)} candidates x {len(AC.freq)} frequencies in {dt:.2f} s = "
      f"{nfreq / dt:.1f} systems/s (N = {base.NumNosC}), last receiver SPL at 100 Hz = {obj.spl[80, 0]:.2f} dB")
