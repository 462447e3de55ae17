"""Time of the volume assembly on the config-4 mesh (Tet4) and on a Tet10 mesh, host clock around synchronised calls."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

for name, grid in (("Tet4 99x133x74", fd.shoebox_tet4(99, 133, 74)), ("Tet10 of 43x57x36", fd.to_tet10(fd.shoebox_tet4(43, 57, 36)))):
    s = fd.System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.assemble_volume(1.21, 343.0)
    s.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        s.assemble_volume(1.21, 343.0)
    s.synchronize()
    t = (time.perf_counter() - t) / 5
    ev = grid.elem_vol
    b = 4 * ev.size + 24 * len(grid.nos) + 4 * ev.shape[1] * ev.size + 16 * s.nnz
    print(f"{name}: N = {grid.NumNosC}, E = {len(ev)}, nnz = {s.nnz}: assemble_volume {t * 1e3:.2f} ms, B_asm = {b / 1e6:.0f} MB -> {b / t / 1e9:.0f} GB/s")
