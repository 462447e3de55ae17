"""A short modal sweep on the config-4 mesh inside an NVTX range ("sweep"), so `ncu --nvtx --nvtx-include "sweep/"` lists the
launches of the sweep only (the modal set-up before it launches thousands of SpMVs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

grid = fd.shoebox_tet4(99, 133, 74)
s = fd.System(0)
s.set_volume_mesh(grid.nos, grid.elem_vol)
s.assemble_volume(1.21, 343.0)
f_cut = float(np.hypot(500.0, 300.0))
x = f_cut / 343.0
s.modal_compute(f_cut, n_hint=int(4 * np.pi / 3 * 30.0 * x ** 3 + np.pi / 4 * 59.0 * x ** 2 + 4.7 * x) + 1, f_acc=500.0)
src = grid.closest_node([1.0, 1.0, 1.2]); rcv = grid.closest_node([2.0, 3.0, 1.25])
rn = np.array([[rcv] * 4], dtype=np.int32); rw = np.array([[1.0, 0, 0, 0]])
f = np.arange(20.0, 501.0, 1.0)[::15]
s.sweep(2 * np.pi * f[:16], None, [src], [1e-4], rn, rw, want_pN=False, tol=1e-8, precond="modal", check_every=16)   # warm-up
s.synchronize()
torch.cuda.nvtx.range_push("sweep")
_, _, st = s.sweep(2 * np.pi * f, None, [src], [1e-4], rn, rw, want_pN=False, tol=1e-8, precond="modal", check_every=16)
s.synchronize()
torch.cuda.nvtx.range_pop()
print(st)
