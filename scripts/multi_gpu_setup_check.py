"""torchrun check of the shared modal set-up: every rank must end with the modes a single process computes, and the sharded
FEM3D.compute() must return what the single-process call returns.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_setup_check.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd  # noqa: E402
import room_setup as rs  # noqa: E402
from femder_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shape = tuple(int(t) for t in (sys.argv[1] if len(sys.argv) > 1 else "43,57,36").split(","))
grid = fd.shoebox_tet4(*shape)
f_cut = float(np.hypot(500.0, 300.0)) if shape[0] > 60 else 330.0

alone = fd.System(local)
alone.set_volume_mesh(grid.nos, grid.elem_vol)
alone.assemble_volume(1.21, 343.0)
alone.modal_compute(f_cut, f_acc=f_cut)          # warm-up (handles, allocations)
alone.synchronize()
t = time.perf_counter()
m1 = alone.modal_compute(f_cut, f_acc=f_cut)
alone.synchronize()
t1 = time.perf_counter() - t
lam1, _ = alone.modal_get(vectors=False)

sharding.share_modal_setup(alone)
alone.modal_compute(f_cut, f_acc=f_cut)
dist.barrier()
torch.cuda.synchronize()
t = time.perf_counter()
m2 = alone.modal_compute(f_cut, f_acc=f_cut)
alone.synchronize()
dist.barrier()
t2 = time.perf_counter() - t
lam2, V2 = alone.modal_get()
info = alone.modal_info()
# all ranks hold the same modes
chk = torch.tensor([float(np.sum(lam2)), float(np.abs(V2).sum())], dtype=torch.float64, device="cuda")
parts = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(parts, chk)
same = all(torch.equal(parts[0], p) for p in parts)
n = min(len(lam1), len(lam2))
rel = np.abs(np.sqrt(np.abs(lam2[:n])) - np.sqrt(np.abs(lam1[:n]))).max() / (2 * np.pi)
if rank == 0:
    print(f"N = {grid.NumNosC}, world {world}: alone {m1} modes in {t1:.3f} s; shared {m2} modes in {t2:.3f} s ({info['outer']} outer iterations, "
          f"worst residual {info['worst_residual']:.2e}); identical on all ranks: {same}; max |f_shared - f_alone| = {rel:.2e} Hz", flush=True)
assert same and abs(m1 - m2) <= max(2, m1 // 50) and rel < 0.05

# sharded FEM3D.compute() against the single-process one
AP = rs.AirProperties(c0=343.0, rho0=1.21)
AC = rs.AlgControls(AP, freq_vec=np.arange(20.0, 301.0, 4.0))
BC = rs.BC(AC, AP)
for g in range(2, 8):
    BC.rigid(g)
S = rs.Source(coord=grid.nos[grid.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
R = rs.Receiver(coord=grid.nos[grid.closest_node([2.0, 3.0, 1.25])])
o1 = fd.FEM3D(grid, S, R, AP, AC, BC, device=local, shard=False)
o1.compute(timeit=False, keep_pN=False)
o2 = fd.FEM3D(grid, S, R, AP, AC, BC, device=local, shard=True)
o2.compute(timeit=False, keep_pN=False)
err = np.abs(o2.pR - o1.pR).max() / np.abs(o1.pR).max()
if rank == 0:
    print(f"sharded compute vs alone: max relative difference {err:.2e}, unconverged {o2.n_unconverged}, precond {o2.stats.precond_used}", flush=True)
assert err < 1e-6 and o2.n_unconverged == 0
dist.barrier()
dist.destroy_process_group()
