"""Build the config-4 system on the GPU and launch only the fused SpMV (K5): target for ncu."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="99,133,74")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--rigid", type=int, default=1)
ap.add_argument("--real", type=int, default=0)
a = ap.parse_args()
grid = fd.shoebox_tet4(*[int(t) for t in a.shape.split(",")])
s = fd.System(0)
s.set_volume_mesh(grid.nos, grid.elem_vol)
s.assemble_volume(1.21, 343.0)
if not a.rigid:
    s.set_surface_mesh(grid.elem_surf, grid.domain_index_surf, np.arange(2, 8))
    s.assemble_surface()
ms, nb = s.spmv_bench(a.batch, a.reps, with_boundary=not a.rigid, real=bool(a.real))
print(f"spmv F={a.batch} real={a.real}: {ms*1e3:.1f} us, {nb/1e6:.1f} MB algorithmic, {nb/ms/1e6:.1f} GB/s")
