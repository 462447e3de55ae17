"""Turn the raw ncu output under gpurun_out/ into the committed summaries under profiles/ (run in the build container)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs(OUT, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches():
    p = os.path.join(GO, f"launches_{R}.csv")
    if not os.path.isfile(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = r[ki].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{R}_launches.md"), "w") as f:
        f.write(f"# {R}: kernel launch list of the sweep in `scripts/sweep_for_ncu.py` (33 frequencies of 20-500 Hz on the 1M-DOF mesh, modal path; "
                "the NVTX range keeps the modal set-up's SpMVs out)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include sweep/` (cold-cache, serialised launches: compare SHARES).\n\n"
                "| kernel | launches | mean us | share of captured time |\n|---|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{n}` | {c} | {t / c / 1000:.1f} | {t / tot:.3f} |\n")
    with open(os.path.join(OUT, f"{R}_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "gpu__time_duration_ns"])
        for r in rows[1:]:
            w.writerow([r[ki].split("(")[0], r[vi]])


def capture(name, title):
    rep = os.path.join(GO, f"{name}_{R}.ncu-rep")
    rawf = os.path.join(GO, f"{name}_{R}.raw.csv")
    if os.path.isfile(rawf) and os.path.getsize(rawf):
        raw = open(rawf).read()
    elif os.path.isfile(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return None
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, d = rows[0], rows[1], rows[2]
    vals = {}
    with open(os.path.join(OUT, f"{R}_{name}.md"), "w") as f:
        f.write(f"# {R}: {title}\n\n`ncu --set full --clock-control none --import-source on`, one launch after warm-up, 1M-DOF config-4 mesh.\n\n"
                f"kernel: `{d[hdr.index('Kernel Name')]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {d[i]} | {units[i]} |\n")
                vals[k] = (d[i], units[i])
    return vals


def to_bytes(v):
    x, u = float(v[0].replace(",", "")), v[1].lower()
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


launches()
traffic = {}
for name, kid, title in (("spmv_real16", "K5", "K5 fused multi-frequency SpMV k_spmv_blob, batch 16, real arithmetic (the variant the rigid benchmark sweep launches)"),
                         ("precond_fused", "K6a", "K6a k_precond_fused (tile-compressed modes), batch 16 real: r -= alpha q, z = B r (bf16 block inverse) and t = Phi^T r on tcgen05"),
                         ("coarse_project", "K8a", "K8a k_modal_project on the coarse vector, batch 16 real, 747 modes: c = C^T t on tcgen05"),
                         ("coarse_expand", "K8c", "K8c k_modal_expand on the coarse vector, batch 16 real, 747 modes: u = C e on tcgen05"),
                         ("coarse_finish", "K8d", "K8d k_coarse_finish, batch 16 real: z += Phi u, rho' = r^T z, ||r||^2, iteration scalars"),
                         ("pupdate_tiles_real16", "K6b", "K6b k_pupdate_tiles, batch 16 real: x += alpha p, p = z + beta p (bulk copies both ways)")):
    v = capture(name, title)
    if v and kid:
        traffic[kid] = to_bytes(v["dram__bytes_read.sum"]) + to_bytes(v["dram__bytes_write.sum"])
if traffic:
    traffic["source"] = f"profiles/{R}_spmv_real16.md, {R}_precond_fused.md, {R}_coarse_project.md, {R}_coarse_expand.md, {R}_coarse_finish.md, {R}_pupdate_tiles_real16.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(traffic, open(os.path.join(OUT, "kernel_traffic.json"), "w"))
for fn in ("bench_%s.json" % R,):
    src = os.path.join(GO, fn)
    if os.path.isfile(src) and os.path.getsize(src):
        open(os.path.join(OUT, f"{R}_bench.json"), "w").write(open(src).read())
print(os.listdir(OUT))
