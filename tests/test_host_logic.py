"""CPU tests of the host-side logic: meshes, receiver stencils, frequency sharding (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from oracle import femder_oracle as O


def test_shoebox_counts_and_orientation():
    from femder_b200.mesh import shoebox_tet4
    g = shoebox_tet4(4, 5, 3, jitter=0.2)
    assert g.NumNosC == 5 * 6 * 4 and g.NumElemC == 6 * 4 * 5 * 3 and len(g.elem_surf) == 2 * 2 * (4 * 5 + 4 * 3 + 5 * 3)
    det, _ = O.tet4_precompute(g.elem_vol, g.nos)
    assert (det > 0).all() and abs(det.sum() / 6 - 30.0) < 1e-12
    assert list(g.number_ID_faces) == [2, 3, 4, 5, 6, 7]
    # every boundary triangle is a face of some tetrahedron
    faces = set()
    for t in np.asarray(g.elem_vol, dtype=np.int64):
        for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)):
            faces.add(tuple(sorted(t[list(f)])))
    assert all(tuple(sorted(t)) in faces for t in np.asarray(g.elem_surf, dtype=np.int64))


def test_config_sizes():
    from femder_b200.mesh import shoebox_tet4
    g = shoebox_tet4(26, 34, 22)
    assert (g.NumNosC, g.NumElemC, len(g.elem_surf)) == (21735, 116688, 8816)


def test_tet10_midpoints():
    from femder_b200.mesh import TET10_EDGES, shoebox_tet4, to_tet10
    g4 = shoebox_tet4(2, 2, 2, jitter=0.1)
    g = to_tet10(g4)
    assert g.order == 2 and g.elem_vol.shape[1] == 10 and g.elem_surf.shape[1] == 6
    ev = np.asarray(g.elem_vol, dtype=np.int64)
    for k, (a, b) in enumerate(TET10_EDGES):
        assert np.allclose(g.nos[ev[:, 4 + k]], 0.5 * (g.nos[ev[:, a]] + g.nos[ev[:, b]]))
    H, Q = O.assemble_tet10(g.elem_vol, g.nos, 343.0, 1.21)
    assert np.abs(H @ np.ones(g.NumNosC)).max() < 1e-11


def test_receiver_stencil_matches_oracle_evaluate():
    from femder_b200 import receiver_stencil
    from femder_b200.mesh import shoebox_tet4
    g = shoebox_tet4(4, 4, 3, jitter=0.2)
    rng = np.random.default_rng(0)
    pN = rng.standard_normal((3, g.NumNosC)) + 1j * rng.standard_normal((3, g.NumNosC))
    pts = np.vstack([g.nos[17], rng.uniform([0.1, 0.1, 0.1], [2.9, 3.9, 2.4], size=(12, 3))])
    ref = O.evaluate(g.nos, g.elem_vol, pN, pts)
    for i, c in enumerate(pts):
        nodes, w = receiver_stencil(g.nos, g.elem_vol, c)
        assert abs(w.sum() - 1) < 1e-12
        assert np.abs(pN[:, nodes] @ w - ref[:, i]).max() < 1e-12


def test_owned_frequencies_partition():
    from femder_b200.sharding import owned_frequencies
    for nf, batch, world in ((481, 8, 8), (181, 8, 2), (5, 8, 4), (0, 4, 2), (17, 1, 3)):
        parts = [owned_frequencies(nf, batch, r, world) for r in range(world)]
        allf = np.sort(np.concatenate(parts)) if parts else np.array([])
        assert np.array_equal(allf, np.arange(nf))
        for p in parts:  # whole batches, ascending
            assert np.all(np.diff(p) > 0) if len(p) > 1 else True
    assert np.array_equal(owned_frequencies(20, 4, 1, 2), [4, 5, 6, 7, 12, 13, 14, 15])


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from femder_b200.sharding import owned_frequencies, gather_rows, dist_info
dist.init_process_group("gloo")
rank, world, _ = dist_info()
nf, batch = 21, 4
owned = owned_frequencies(nf, batch, rank, world)
freq = np.arange(nf) + 20.0
local = np.stack([np.exp(1j * freq[owned]), freq[owned] * (1 + 2j)], axis=1)   # stands in for the solver's pR rows
full = gather_rows(local, owned, nf)
ref = np.stack([np.exp(1j * freq), freq * (1 + 2j)], axis=1)
assert full.shape == (nf, 2) and np.array_equal(full, ref), rank
assert gather_rows(None, owned, nf) is None
dist.barrier()
dist.destroy_process_group()
sys.stdout.write(f"rank {rank} ok\n"); sys.stdout.flush()
'''


def test_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_deal_by_cost_spreads_the_expensive_items():
    from femder_b200 import sharding
    cost = np.ones(37)
    cost[[3, 11, 12, 30, 31, 36]] = 10.0
    seen = []
    for r in range(4):
        idx = sharding.deal_by_cost(cost, r, 4)
        assert np.all(np.diff(idx) > 0)
        assert 1 <= (cost[idx] > 1).sum() <= 2          # 6 expensive items over 4 ranks
        seen.extend(idx.tolist())
    assert sorted(seen) == list(range(37))
    assert np.array_equal(sharding.deal_by_cost(cost, 0, 1), np.arange(37))


_QUEUE_WORKER = r'''
import ctypes, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from femder_b200 import sharding
dist.init_process_group("gloo")
rank, world, _ = sharding.dist_info()
assert sharding.same_node()
q = sharding.counters()
assert q is sharding.counters()
k = q.fresh()
assert k == 0 and q.values[k] == 0
# every rank claims items from the common queue the way fdb_sweep does (atomic fetch-and-add on the shared counter)
import fcntl
n_items, mine = 1000, []
add = ctypes.c_int64.from_address(q.address(k))
while True:
    # the library claims with __atomic_fetch_add; two Python processes emulate the atomicity with an advisory file lock
    with open(sys.argv[2], "a") as lf:
        fcntl.flock(lf, fcntl.LOCK_EX)
        i = add.value
        add.value = i + 1
        fcntl.flock(lf, fcntl.LOCK_UN)
    if i >= n_items:
        break
    mine.append(i)
parts = [None] * world
dist.all_gather_object(parts, mine)
allc = sorted(sum(parts, []))
assert allc == list(range(n_items)), (len(allc), rank)          # every item exactly once
assert q.values[k] >= n_items
k2 = q.fresh()
assert k2 == 1 and q.values[k2] == 0
# the rows of a dynamically claimed sweep are gathered like statically dealt ones
rows = np.asarray(mine, dtype=np.float64)[:, None] * (1 + 1j)
full = sharding.gather_rows(rows, np.asarray(mine, dtype=np.int64), n_items)
assert np.array_equal(full[:, 0], np.arange(n_items) * (1 + 1j))
dist.barrier()
dist.destroy_process_group()
sys.stdout.write(f"rank {rank} ok {len(mine)}\n"); sys.stdout.flush()
'''


def test_shared_queue_world_size_2_gloo(tmp_path):
    """The common work queue of a sharded sweep (sharding.SharedCounters: a counter in POSIX shared memory that fdb_sweep claims
    systems from): both ranks see the same memory, every item is claimed exactly once, counters are handed out in step."""
    script = tmp_path / "queue_worker.py"
    script.write_text(_QUEUE_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29618", str(script), ROOT, str(tmp_path / "lock")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
