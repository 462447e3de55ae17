"""TEST / BENCH INFRASTRUCTURE ONLY - times the reference's CPU algorithm for ``bench.py``'s ``cpu_baseline`` leg
and its ``--impl reference`` arm.

One frequency is solved exactly as the reference does it (``femder/FEM_3D.py:1312-1318``): build
``G = H - w^2 Q`` (rigid walls), factorise it from scratch, solve, discard the factors.  scipy SuperLU stands in
for MKL PARDISO (``pyMKL`` is un-vendored and not installable here).  SuperLU is single-threaded, so to use every
host core a step solves ``cores`` frequencies at once, one per worker process; the reference's PARDISO would
instead thread one factorisation.

The 1M-DOF factorisation does not fit host memory (3-D fill), so the sample mesh is config 1's 21 735-DOF shoebox.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_STATE = {}


def _setup(shape, src, rcv, c0, rho0):
    from femder_b200.mesh import shoebox_tet4
    from oracle import femder_oracle as O
    g = shoebox_tet4(*shape)
    H, Q = O.assemble_tet4(g.elem_vol, g.nos, c0, rho0)
    q = O.source_vector(g.nos, [src], [1e-4])
    _STATE.update(H=H, Q=Q, q=q, rcv=g.closest_node(rcv), N=g.NumNosC)


def _ready(_):
    return _STATE["N"]


def _solve(freq):
    from scipy.sparse.linalg import splu

    from oracle import femder_oracle as O
    w = 2 * np.pi * freq
    G = O.system_matrix(_STATE["H"], _STATE["Q"], [], [], w)
    p = splu(G).solve(-1j * w * _STATE["q"])
    return complex(p[_STATE["rcv"]])


def run(steps, warmup, shape=(26, 34, 22), cores=None, f_lo=20.0, f_hi=500.0, src=(1.0, 1.0, 1.2), rcv=(2.0, 3.0, 1.25),
        c0=343.0, rho0=1.21, workload_dof=None):
    cores = cores or os.cpu_count() or 1
    total_steps = steps + warmup
    rng = np.random.default_rng(0)
    freqs = rng.permutation(np.linspace(f_lo, f_hi, total_steps * cores)).reshape(total_steps, cores)
    ctx = mp.get_context("fork")   # callers must not hold a CUDA context (bench.py runs this in a CUDA-free subprocess)
    times = []
    with ctx.Pool(cores, initializer=_setup, initargs=(shape, list(src), list(rcv), c0, rho0)) as pool:
        n_dof = pool.map(_ready, range(cores), chunksize=1)[0]  # workers have assembled (untimed)
        for s in range(total_steps):
            t = time.perf_counter()
            pool.map(_solve, list(freqs[s]), chunksize=1)
            dt = time.perf_counter() - t
            if s >= warmup:
                times.append(dt)
    total = float(np.sum(times))
    measured = steps * cores / total
    scale = (n_dof / workload_dof) if workload_dof else 1.0
    sample = (f"oracle port of the reference algorithm (FP64 assembly, then per frequency: build G, scipy SuperLU factorise + solve, "
              f"PARDISO stand-in) on the config-1 shoebox {shape[0]}x{shape[1]}x{shape[2]} cells = {n_dof} DOF; {steps} steps x {cores} "
              f"frequencies (one per core) drawn from {f_lo:g}-{f_hi:g} Hz: MEASURED {measured:.4g} frequencies/s at {n_dof} DOF. "
              f"A 1M-DOF sparse LU does not fit this host's RAM, so the reference's direct solve cannot run the workload at all")
    if workload_dof:
        sample += (f"; `value` is that measurement scaled LINEARLY in DOF to the {workload_dof}-DOF workload (x {scale:.5f}), an "
                   f"extrapolation that favours the reference (sparse LU cost grows faster than linearly, ~N^2 in 3-D)")
    return {"value": measured * scale, "measured_at_sample": measured, "seconds": total, "cores": cores, "N": n_dof, "sample": sample}
