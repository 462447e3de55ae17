"""TEST / BENCH INFRASTRUCTURE ONLY - times the reference's CPU algorithm for ``bench.py``'s ``cpu_baseline`` leg
and its ``--impl reference`` arm.

One frequency is solved exactly as the reference does it (``femder/FEM_3D.py:1312-1318``): build
``G = H - w^2 Q`` (rigid walls), factorise it from scratch, solve, discard the factors.  scipy SuperLU stands in
for MKL PARDISO (``pyMKL`` is un-vendored and not installable here).  SuperLU is single-threaded, so to use the
host's cores a step solves several frequencies at once, one per worker process; the reference's PARDISO would
instead thread one factorisation.

The 1M-DOF factorisation does not fit host memory (3-D fill), so the path is MEASURED on a ladder of smaller shoebox
meshes of the same room (``LADDER``) and the throughput at the benchmark's size is an extrapolation of the fitted power
law ``seconds per factorise-and-solve = a * DOF^p`` - labelled as such wherever it is printed.
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


def measure(shape, steps, warmup, workers, f_lo=20.0, f_hi=500.0, src=(1.0, 1.0, 1.2), rcv=(2.0, 3.0, 1.25), c0=343.0, rho0=1.21):
    """``steps`` timed steps of ``workers`` concurrent factorise-and-solve calls on the shoebox ``shape``; returns
    dict(N, frequencies_per_s, seconds, workers, steps)."""
    total_steps = steps + warmup
    rng = np.random.default_rng(0)
    freqs = rng.permutation(np.linspace(f_lo, f_hi, total_steps * workers)).reshape(total_steps, workers)
    ctx = mp.get_context("fork")   # callers must not hold a CUDA context (bench.py runs this in a CUDA-free subprocess)
    times = []
    with ctx.Pool(workers, initializer=_setup, initargs=(shape, list(src), list(rcv), c0, rho0)) as pool:
        n_dof = pool.map(_ready, range(workers), chunksize=1)[0]  # workers have assembled (untimed)
        for s in range(total_steps):
            t = time.perf_counter()
            pool.map(_solve, list(freqs[s]), chunksize=1)
            dt = time.perf_counter() - t
            if s >= warmup:
                times.append(dt)
    total = float(np.sum(times))
    return {"N": int(n_dof), "shape": list(shape), "frequencies_per_s": steps * workers / total, "seconds": total, "workers": int(workers),
            "steps": int(steps), "seconds_per_solve": total / steps}


# the room of the benchmark (3.0 x 4.0 x 2.5 m) at growing resolution: 21 735, 47 250 and 94 424 DOF
LADDER = ((26, 34, 22), (34, 44, 29), (43, 57, 36))


def fit_power_law(points):
    """Least-squares fit of log(seconds per factorise-and-solve) = log a + p log N; returns (a, p)."""
    n = np.log([p["N"] for p in points])
    y = np.log([p["seconds_per_solve"] for p in points])
    if len(points) == 1:
        return float(np.exp(y[0]) / np.exp(n[0])), 1.0          # one size only: linear in DOF (favours the reference)
    A = np.vstack([np.ones_like(n), n]).T
    (la, p), *_ = np.linalg.lstsq(A, y, rcond=None)
    return float(np.exp(la)), float(p)


def run(steps, warmup, sizes=2, cores=None, f_lo=20.0, f_hi=500.0, src=(1.0, 1.0, 1.2), rcv=(2.0, 3.0, 1.25),
        c0=343.0, rho0=1.21, workload_dof=None):
    """Measure the first ``sizes`` rungs of the ladder (every host core busy, fewer workers on the largest meshes whose factors
    are gigabytes each), fit the power law, and report the throughput extrapolated to ``workload_dof``."""
    cores = cores or os.cpu_count() or 1
    points = []
    for i, shape in enumerate(LADDER[:max(1, sizes)]):
        workers = max(1, min(cores, (cores, max(3, cores // 2), max(3, cores // 4))[i]))
        st = steps if i == 0 else 1
        wu = warmup if i == 0 else 0
        points.append(measure(shape, st, wu, workers, f_lo, f_hi, src, rcv, c0, rho0))
    a, p = fit_power_law(points)
    measured = points[0]["frequencies_per_s"]
    # one factorisation per core, as on the smallest mesh - at the workload's size not even one fits in memory: favours the reference
    value = cores / (a * float(workload_dof) ** p) if workload_dof else measured
    pts = "; ".join(f"{q['N']} DOF: {q['frequencies_per_s']:.4g} frequencies/s ({q['workers']} concurrent factorisations, "
                    f"{q['steps'] * q['workers']} frequencies in {q['seconds']:.1f} s)" for q in points)
    sample = (f"oracle port of the reference algorithm (FP64 assembly, then per frequency: build G, scipy SuperLU factorise + solve, "
              f"PARDISO stand-in), rigid shoebox 3.0x4.0x2.5 m, frequencies drawn from {f_lo:g}-{f_hi:g} Hz, MEASURED at {pts}. "
              f"A 1M-DOF sparse LU does not fit this host's RAM, so the reference's direct solve cannot run the workload at all")
    if workload_dof:
        how = (f"{cores} cores / (fitted seconds per factorise-and-solve = {a:.4g} * DOF^{p:.3f} over the measured sizes), i.e. as if every "
               f"core could hold a factorisation of that size (the larger measured meshes already cannot: fewer concurrent ones)" if len(points) > 1
               else "ONE measured size scaled linearly in DOF (favours the reference: sparse LU cost grows ~DOF^2 in 3-D)")
        sample += f"; `value` is an EXTRAPOLATION to the {workload_dof}-DOF workload: {how}"
    return {"value": value, "measured_at_sample": measured, "seconds": float(sum(q["seconds"] for q in points)), "cores": cores,
            "N": points[0]["N"], "points": points, "fit": {"a": a, "p": p}, "sample": sample}
