"""Robustness checks at scale on the GPU box: Tet10 with long rows, jittered 1M-DOF mesh with an admittance wall (complex
arithmetic + overlay), many receivers.  Every solve is verified by the library's true-residual check (relres)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femder_b200 as fd
import room_setup as rs  # noqa: E402

def run(name, grid, freqs, admittance, n_rcv=3, tol=1e-8):
    AP = rs.AirProperties()
    AC = rs.AlgControls(AP, freq_vec=np.asarray(freqs, dtype=float))
    BC = rs.BC(AC, AP)
    for t in range(2, 8):
        BC.rigid(t)
    if admittance:
        BC.normalized_admittance(7, admittance)
    rng = np.random.default_rng(1)
    S = rs.Source(coord=grid.nos[grid.closest_node([1.0, 1.0, 1.2])], q=[1e-4])
    R = rs.Receiver(coord=rng.uniform([0.1, 0.1, 0.1], [2.9, 3.9, 2.4], size=(n_rcv, 3)))
    obj = fd.FEM3D(grid, S, R, AP, AC, BC, tol=tol)
    t = time.perf_counter(); obj.compute(timeit=False, keep_pN=False); t = time.perf_counter() - t
    st = obj.stats
    print(f"{name}: N={grid.NumNosC} nnz/row={obj.H.nnz/grid.NumNosC:.1f} freqs={len(freqs)} wall={t:.2f}s sweep={st.seconds:.2f}s real={st.real_arithmetic} "
          f"iters[min/mean/max]={obj.iters.min()}/{obj.iters.mean():.0f}/{obj.iters.max()} max_relres={obj.relres.max():.2e} unconverged={st.n_unconverged} "
          f"pR finite={np.isfinite(obj.pR).all()} shape={obj.pR.shape}", flush=True)
    assert st.n_unconverged == 0 and np.isfinite(obj.pR).all()

run("tet10 admittance", fd.to_tet10(fd.shoebox_tet4(24, 30, 20, jitter=0.1)), np.arange(20, 201, 10), 0.02)
run("tet10 rigid", fd.to_tet10(fd.shoebox_tet4(24, 30, 20, jitter=0.1)), np.arange(20, 201, 10), 0.0)
run("tet4 1M jitter admittance", fd.shoebox_tet4(99, 133, 74, jitter=0.2), [20, 60, 100, 140, 180, 220, 260, 300], 0.05, n_rcv=1000)
run("tet4 1M jitter rigid", fd.shoebox_tet4(99, 133, 74, jitter=0.2), [30, 90, 150, 210], 0.0, n_rcv=5)
