"""Timing experiments on k_spmv_blob at config-4 size (FDB_BLOB_DBG switches off parts of the kernel; results are wrong
on purpose): what bounds the kernel - the fill, the shared-memory compute, or their overlap?"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402

grid = fd.shoebox_tet4(99, 133, 74)
names = {0: "full kernel", 1: "no compute (fill + y store)", 2: "no x fill (matrix stream + compute)", 3: "matrix stream + y store only",
         4: "x fill + compute on a stale matrix", 5: "x fill + y store only", 7: "y store only"}
TILES = [int(t) for t in os.environ.get("EXP_TILES", "128,64").split(",")]
LPRS = [int(t) for t in os.environ.get("EXP_LPRS", "2,4").split(",")]
DBGS = [int(t) for t in os.environ.get("EXP_DBGS", "0,1,2,3,5,7").split(",")]
for tile in TILES:
    os.environ["FDB_TILE_ROWS"] = str(tile)   # read when the internal order is built
    s = fd.System(0)
    s.set_volume_mesh(grid.nos, grid.elem_vol)
    s.assemble_volume(1.21, 343.0)
    for lpr in LPRS:
        os.environ["FDB_BLOB_LPR"] = str(lpr)
        for dbg in DBGS:   # 4 alone (compute on a stale matrix stream) reads wild positions: not run
            os.environ["FDB_BLOB_DBG"] = str(dbg)
            out = []
            for batch, real in ((16, True), (8, False)):   # alternating widths forces a fresh work set, which re-reads the environment
                ms, nb = s.spmv_bench(batch, 20, with_boundary=False, real=real)
                out.append(f"F={batch} {'real' if real else 'cpx '} {ms*1e3:6.1f} us")
            print(f"tile={tile} LPR={lpr} dbg={dbg} {names[dbg]:40s}: " + " | ".join(out), flush=True)
    del s
