#!/usr/bin/env python
"""bench.py - frequencies/sec of the 1M-DOF Tet4 sweep (BASELINE.json metric, config 4) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (CUDA, through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]    # the reference's CPU algorithm (oracle port)

Workload: synthetic shoebox 3.0 x 4.0 x 2.5 m, 99 x 133 x 74 cells -> 1 005 000 nodes, 5 846 148 Tet4, rigid walls,
point source q = 1e-4 near (1.0, 1.0, 1.2), receiver near (2.0, 3.0, 1.25), sweep 20..500 Hz at 1 Hz (481
frequencies), every frequency converged to 1e-8 relative residual.

A "step" is one call of the hot path on one batch of synthetic input: 96 (97 for the first) frequencies taken
every `--stride` (5) Hz across the sweep, f = 20 + s + 5 j Hz for step s, so every step samples the whole 20..500 Hz
range and 5 consecutive steps are exactly the 481-frequency sweep (a user's call solves all 481 at once; a fifth of
it keeps the default run within minutes while the tail of a call - the last, hardest systems running in a
narrowing batch - weighs about what it does in the full sweep).  Inside the call the frequencies run through
`--batch` (16) solver slots of the fused multi-frequency SpMV; a slot whose frequency converged is reloaded with
the next one.  With N ranks, rank r takes steps r, r+N, ... (frequency sharding, no collective on the solve
path; weak scaling: per-GPU work per step is fixed whatever N is).

Timing: W >= 3 untimed warm-up steps, then exactly K steps bracketed by a barrier + torch.cuda.synchronize() on
both sides and by CUDA events on the stream the library launches on; the reported time is the MAX over ranks.
Working set per step (matrix 0.3 GB + 5 vectors x 128 MB) exceeds the 126 MB L2, so no L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C0, RHO0 = 343.0, 1.21
SHAPE = (99, 133, 74)
F_LO, F_HI = 20.0, 500.0
SRC, RCV = [1.0, 1.0, 1.2], [2.0, 3.0, 1.25]
METRIC = "frequencies/sec for 1M-DOF Tet4 sweep"
WORKLOAD = ("synthetic shoebox 3.0x4.0x2.5 m, 99x133x74 cells, 1005000 DOF Tet4, rigid walls, 1 source, 1 receiver, "
            "20-500 Hz at 1 Hz, tol 1e-8")


STEP_STRIDE = 5  # Hz between the frequencies of one step (--stride); STEP_STRIDE steps tile the sweep


def step_frequencies(i):
    """Frequencies of step i: every STEP_STRIDE-th frequency of the sweep starting at 20 + (i mod STEP_STRIDE) Hz."""
    return np.arange(F_LO, F_HI + 1.0, 1.0)[(i % STEP_STRIDE)::STEP_STRIDE]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's algorithm (assembly + sparse direct solve per frequency)
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, shape=(26, 34, 22), cores=None):
    from oracle import cpu_baseline
    workload_dof = (SHAPE[0] + 1) * (SHAPE[1] + 1) * (SHAPE[2] + 1)
    return cpu_baseline.run(steps, warmup, shape=shape, cores=cores, f_lo=F_LO, f_hi=F_HI, src=SRC, rcv=RCV, c0=C0, rho0=RHO0,
                            workload_dof=workload_dof)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if os.environ.get("FDB_REF_BLAS1") != "1":
        # One SuperLU per core: BLAS must be single-threaded in every worker, or the spinning BLAS pools of
        # concurrent factorisations starve each other (measured: 2 concurrent solves never finish).  The thread
        # count is read when numpy loads, so restart this process with the environment set.
        env = dict(os.environ, FDB_REF_BLAS1="1", OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        os.execve(sys.executable, [sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env)
    r = cpu_reference_run(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frequencies/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_sample": r["sample"]},
            "measured_at_sample": r["measured_at_sample"], "sample_dof": r["N"],
            "cpu_baseline": {"value": r["value"], "unit": "frequencies/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                             "measured_at_sample": r["measured_at_sample"], "sample_dof": r["N"]},
            "e2e": {"value": r["value"], "unit": "frequencies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import femder_b200 as fd
    from femder_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if _lib.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device - the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream(device=local_rank)

    shape = tuple(int(t) for t in args.shape.split(","))
    F = args.batch

    # ---- inputs: host mesh arrays in pinned memory (the e2e leg copies them every step) ----
    grid = fd.shoebox_tet4(*shape)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t

    keep = []
    nos, t0 = pinned(grid.nos); keep.append(t0)
    ev, t1 = pinned(np.asarray(grid.elem_vol, dtype=np.int32)); keep.append(t1)
    es, t2 = pinned(np.asarray(grid.elem_surf, dtype=np.int32)); keep.append(t2)
    grid.nos, grid.elem_vol, grid.elem_surf = nos, ev, es
    src_node = grid.closest_node(SRC)
    rcv_node = grid.closest_node(RCV)
    group_ids = np.arange(2, 8)

    # ---- resident system: assemble once on the GPU ----
    sysm = fd.System(local_rank, stream.cuda_stream)
    t_setup = time.perf_counter()
    sysm.set_volume_mesh(nos, ev)
    sysm.assemble_volume(RHO0, C0)
    sysm.set_surface_mesh(es, grid.domain_index_surf, group_ids)
    sysm.assemble_surface()
    sysm.synchronize()
    t_setup = time.perf_counter() - t_setup
    mu_zero = np.zeros((len(group_ids), 512), dtype=np.complex128)
    rcv_nodes = np.array([[rcv_node] * 4], dtype=np.int32)
    rcv_w = np.array([[1.0, 0.0, 0.0, 0.0]])

    batch_freqs = step_frequencies

    def step_resident(i):
        f = batch_freqs(i)
        _, pR, st = sysm.sweep(2 * np.pi * f, mu_zero[:, :len(f)], [src_node], [1e-4], rcv_nodes, rcv_w, want_pN=False,
                               tol=args.tol, max_iter=args.max_iter, batch=F, check_every=args.check_every)
        return len(f), st, pR

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, idx0):
        """K steps of fn bracketed by barrier+sync and CUDA events on the library's stream; MAX over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nfreq, nk, its, unconv, extra = 0, 0, [], 0, []
        with torch.cuda.stream(stream):
            e0.record(stream)
            for s in range(args.steps):
                n, st, x = fn(idx0 + s * world + rank)
                nfreq += n; nk += st.n_kernels; its.extend(st.iters.tolist()); unconv += st.n_unconverged; extra.append(x)
            e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local_rank}")
        tot = torch.tensor([float(nfreq), float(nk), float(unconv), float(np.sum(its))], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return float(ms.item()), tot.tolist(), its

    # ---- warm-up (untimed) then the timed region ----
    idx_w = args.steps * world  # warm-up uses batches after the timed ones in the visiting order
    for s in range(args.warmup):
        step_resident(idx_w + s * world + rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, tot, its = timed(step_resident, 0)
    clocks = sampler.stop() if rank == 0 else None
    value = tot[0] / (ms * 1e-3)

    # ---- e2e: the user's call, FEM3D.compute(), from pinned HOST mesh buffers to host receiver pressures ----
    AP = fd.AirProperties(c0=C0, rho0=RHO0)
    S = fd.Source(coord=grid.nos[src_node], q=[1e-4])
    R = fd.Receiver(coord=grid.nos[rcv_node])
    h2d = nos.nbytes + ev.nbytes + es.nbytes + 4 * len(es) + 16 * (481 // STEP_STRIDE + 1) * (1 + len(group_ids))
    d2h_holder = [0]

    def step_e2e(i):
        f = batch_freqs(i)
        AC = fd.AlgControls(AP, freq_vec=f)
        BC = fd.BC(AC, AP)
        for g in group_ids:
            BC.rigid(int(g))
        obj = fd.FEM3D(grid, S, R, AP, AC, BC, device=local_rank, stream=stream.cuda_stream, tol=args.tol,
                       max_iter=args.max_iter, batch=F, check_every=args.check_every, shard=False)  # every rank runs its own step
        obj._sys = sysm            # reuse the device allocation; everything below is recomputed from the host buffers
        obj._sys_key = None        # force mesh upload + pattern build + assembly (a fresh object is not assembled)
        obj.compute(timeit=False, keep_pN=False)
        # read back: Heron areas, receiver pressures, iteration counts, residuals (H, Q, A stay on the GPU until asked for)
        d2h_holder[0] = obj.areas.nbytes // 2 + obj.pR.nbytes + obj.iters.nbytes + obj.relres.nbytes
        return len(f), obj.stats, obj.pR

    e2e = None
    if not args.no_e2e:
        step_e2e(idx_w + rank)  # one warm-up of the e2e path
        ms_e, tot_e, _ = timed(step_e2e, 0)
        e2e = {"value": tot_e[0] / (ms_e * 1e-3), "unit": "frequencies/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h_holder[0]), "ms_per_step": ms_e / args.steps,
               "call": "FEM3D.compute(keep_pN=False): mesh upload, pattern + assembly, receiver location, sweep, receiver read-back"}

    # ---- rooflines of the three kernels of a solver iteration, each measured live with CUDA events on the launching
    #      stream (50 back-to-back launches on resident vectors, working set > L2); `roofline` = the dominant one ----
    peaks, peak_src = measured_peaks()
    reps = 50
    real = F >= 4  # the rigid workload runs in real arithmetic whenever the batch width allows it
    bj = real and F <= 16
    traffic_file = {}
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.isfile(tp):
        try:
            traffic_file = json.load(open(tp))
        except Exception:
            traffic_file = {}
    names = {0: ("K5", "k_spmv_blob (fused multi-frequency SpMV, shared-memory x tiles)"),
             1: ("K6a", "k_update_bj (residual update + block-Jacobi z = B r out of shared memory + rho, ||r||^2)" if bj
                 else "k_update (residual update + Jacobi + rho, ||r||^2)"),
             2: ("K6b", "k_pupdate_bj (x += alpha p, p = z + beta p)" if bj else "k_pupdate (x += alpha p, p = z + beta p)")}
    kernels = []
    for which in (0, 1, 2):
        sysm.iteration_enqueue(F, 1, -1, real=real, block_jacobi=bj)      # fresh vectors and scalars, every slot active
        sysm.iteration_enqueue(F, 5, which, real=real, block_jacobi=bj)   # warm-up
        barrier()
        with torch.cuda.stream(stream):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            nbytes = sysm.iteration_enqueue(F, reps, which, real=real, block_jacobi=bj)
            a1.record(stream)
        torch.cuda.synchronize()
        us = a0.elapsed_time(a1) / reps * 1e3
        achieved = nbytes / (us * 1e-6) / 1e9
        kernels.append({"id": names[which][0], "kernel": names[which][1] + ", batch %d, %s arithmetic" % (F, "real" if real else "complex"),
                        "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                        "traffic": traffic_file.get(names[which][0]), "algorithmic_bytes_per_launch": int(nbytes), "us_per_launch": us})
    dom = max(kernels, key=lambda k: k["us_per_launch"])
    roofline = dict(dom)
    roofline.update({"peak_source": peak_src,
                     "how": f"{reps} back-to-back launches on resident vectors (matrix + vectors > L2), torch CUDA events on the launching stream",
                     "iteration_kernels": kernels,
                     "iteration": {"us": sum(k["us_per_launch"] for k in kernels),
                                   "algorithmic_bytes": sum(k["algorithmic_bytes_per_launch"] for k in kernels),
                                   "frac": sum(k["algorithmic_bytes_per_launch"] for k in kernels) / (sum(k["us_per_launch"] for k in kernels) * 1e-6) / 1e9 / peaks["hbm_gbs"]}})

    # ---- the CPU baseline's own mesh (config 1, 21 735 DOF) through the same public call: the two arms' `value`s are on
    #      different problem sizes (the reference cannot factorise 1M DOF), this one is the like-for-like comparison
    same_mesh = None
    if rank == 0 and not args.no_e2e:
        g1 = fd.shoebox_tet4(26, 34, 22)
        f1 = np.arange(F_LO, F_HI + 1.0, 1.0)
        AC1 = fd.AlgControls(AP, freq_vec=f1)
        BC1 = fd.BC(AC1, AP)
        for g in group_ids:
            BC1.rigid(int(g))
        S1 = fd.Source(coord=g1.nos[g1.closest_node(SRC)], q=[1e-4])
        R1 = fd.Receiver(coord=g1.nos[g1.closest_node(RCV)])
        o1 = fd.FEM3D(g1, S1, R1, AP, AC1, BC1, device=local_rank, tol=args.tol, max_iter=args.max_iter, check_every=args.check_every,
                      shard=False)
        o1.compute(timeit=False, keep_pN=False)      # warm-up (graphs, allocations)
        o1.H = None
        t1 = time.perf_counter()
        o1.compute(timeit=False, keep_pN=False)
        t1 = time.perf_counter() - t1
        same_mesh = {"value": len(f1) / t1, "unit": "frequencies/s", "N": int(g1.NumNosC), "frequencies": int(len(f1)), "seconds": t1,
                     "iters_mean": float(np.mean(o1.iters)), "unconverged": int(o1.stats.n_unconverged),
                     "what": "FEM3D.compute(keep_pN=False) on the cpu_baseline mesh (config-1 shoebox 26x34x22, rigid), whole 20-500 Hz sweep, "
                             "host wall clock incl. upload, pattern, assembly, sweep"}

    line = None
    if rank == 0:
        its_a = np.asarray(its)
        line = {"metric": METRIC, "value": value, "unit": "frequencies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "step": f"one sweep call over the {481 // STEP_STRIDE}-{481 // STEP_STRIDE + 1} frequencies 20+s+{STEP_STRIDE}j Hz (COCG to {args.tol:g}, {F} solver slots, "
                           f"continuous refill); {STEP_STRIDE} steps = the whole sweep", "batch": F, "parallelism": f"frequency-sharded x{world}",
                           "l2": "working set per step 0.9 GB > 126 MB L2 (no flush needed)", "N": int(grid.NumNosC), "nnz": int(sysm.nnz),
                           "setup_seconds_once": t_setup},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot[1]),
                "solver": {"iters_mean": float(its_a.mean()), "iters_min": int(its_a.min()), "iters_max": int(its_a.max()),
                           "unconverged": int(tot[2]), "frequencies": int(tot[0])},
                "roofline": roofline, "same_mesh_as_cpu_baseline": same_mesh}
        if world == 1 and not args.no_cpu:
            # CUDA-free child process: the CPU baseline forks worker processes, which must not inherit a CUDA context
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_steps),
                                      "--warmup", "1"], capture_output=True, text=True, timeout=900)
                ref = json.loads(out.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
            except Exception as e:  # the bench line must still be printed
                line["cpu_baseline"] = {"value": None, "unit": "frequencies/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"cpu baseline failed: {e!r}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    global STEP_STRIDE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--stride", type=int, default=STEP_STRIDE, help="Hz between the frequencies of one step: 481/stride frequencies per step")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="solver slots (16: real arithmetic for the rigid workload)")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-iter", type=int, default=200000)
    ap.add_argument("--check-every", type=int, default=64)
    ap.add_argument("--shape", default=",".join(str(s) for s in SHAPE))
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    STEP_STRIDE = max(1, args.stride)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
