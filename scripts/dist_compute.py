"""torchrun target: FEM3D.compute() sharded over the ranks (one process per GPU); rank 0 checks against a single-rank run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402
import room_setup as rs  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
grid = fd.shoebox_tet4(12, 16, 10, jitter=0.2)
AP = rs.AirProperties()
AC = rs.AlgControls(AP, 20, 120, 3)
BC = rs.BC(AC, AP)
for t in range(2, 8):
    BC.rigid(t)
BC.normalized_admittance(7, 0.05)
S = rs.Source(coord=[1.0, 1.0, 1.2], q=[1e-4])
R = rs.Receiver(coord=[[2.0, 3.0, 1.25], [0.41, 3.6, 0.77]])
obj = fd.FEM3D(grid, S, R, AP, AC, BC, tol=1e-10)
obj.compute(timeit=False)
pR = obj.evaluate(R)
single = fd.FEM3D(grid, S, R, AP, AC, BC, tol=1e-10, shard=False, device=local_rank)
single.compute(timeit=False)
err = np.linalg.norm(obj.pN - single.pN) / np.linalg.norm(single.pN)
print(f"rank {dist.get_rank()}/{dist.get_world_size()}: pN {obj.pN.shape} sharded-vs-single rel {err:.2e}, iters {obj.iters.sum()}", flush=True)
assert obj.pN.shape == (len(AC.freq), grid.NumNosC) and err < 1e-8
dist.barrier()
dist.destroy_process_group()
