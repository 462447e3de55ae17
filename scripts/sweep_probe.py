"""One sweep through FEM3D.compute on a shoebox of a given size: target for `ncu --metrics gpu__time_duration.sum`
launch lists at sizes other than the benchmark's (per-kernel time of an iteration on small, L2-resident meshes)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402
import room_setup as rs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="26,34,22")
ap.add_argument("--batch", default="auto")
ap.add_argument("--f0", type=float, default=300.0)
ap.add_argument("--nf", type=int, default=64)
ap.add_argument("--rigid", type=int, default=1)
ap.add_argument("--tet10", type=int, default=0)
ap.add_argument("--check-every", type=int, default=64)
a = ap.parse_args()
g = fd.shoebox_tet4(*[int(t) for t in a.shape.split(",")])
if a.tet10:
    g = fd.to_tet10(g)
AP = rs.AirProperties()
AC = rs.AlgControls(AP, freq_vec=a.f0 + np.arange(a.nf))
BC = rs.BC(AC, AP)
for t in range(2, 8):
    BC.rigid(t)
if not a.rigid:
    BC.normalized_admittance(7, 0.02)
S = rs.Source(coord=g.nos[g.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
R = rs.Receiver(coord=g.nos[g.closest_node([2.0, 3.0, 1.25])])
batch = a.batch if a.batch == "auto" else int(a.batch)
o = fd.FEM3D(g, S, R, AP, AC, BC, tol=1e-8, batch=batch, check_every=a.check_every)
t = time.perf_counter()
o.compute(timeit=False, keep_pN=False)
t = time.perf_counter() - t
st = o.stats
print(f"N={g.NumNosC} batch={a.batch} rigid={a.rigid} tet10={a.tet10}: wall {t:.2f}s sweep {st.seconds:.3f}s n_spmv {st.n_spmv} "
      f"-> {st.seconds / max(1, st.n_spmv) * 1e6:.1f} us per batched iteration, iters mean {o.iters.mean():.0f}, "
      f"unconv {st.n_unconverged}, real {st.real_arithmetic}")
