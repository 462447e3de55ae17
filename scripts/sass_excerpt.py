"""profiles/<round>_sass_excerpt.md: mnemonic counts of the solver iteration's kernels and the first occurrence of every tensor /
TMA / TMEM instruction, from `cuobjdump -sass` of the built library (run in the build container, no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
LIB = os.path.join(ROOT, "femder_b200", "libfemder_b200.so")
WANT = [("k_spmv_blob", "ILb0ELb1ELb0ELi2ELi128ELb0ELb0EE"), ("k_precond_fused", "ILb0ELb0ELb1EE"), ("k_modal_project", ""), ("k_modal_coef", "ILb0EE"),
        ("k_modal_expand", "ILb0ELb0ELb1EE"), ("k_coarse_finish", "ILb0ELb0EE"), ("k_pupdate_tiles", "ILb0EE"), ("k_tile_basis", ""),
        ("k_precond_fused", "ILb0ELb0ELb0EE"), ("k_modal_expand", "ILb0ELb0ELb0EE")]
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "DFMA", "FFMA", "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR", "F2F", "SHFL", "HMMA", "IMMA"]
SHOW = ["UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = {}
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
    elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[name].append(line.rstrip())
out = [f"# {R}: SASS evidence (`cuobjdump -sass femder_b200/libfemder_b200.so`, sm_100a; `scripts/sass_excerpt.py`)",
       "Mnemonic counts per kernel of the solver iteration and of the modal set-up's tile-basis kernel (the template instantiation the rigid "
       "benchmark sweep launches; the last two are the full-resolution-tile variants), and the first occurrence of every tensor / TMA / TMEM "
       "instruction with its neighbours.",
       "`UTCHMMA` = `tcgen05.mma` (kind::f16, bf16 operands), `LDTM` = `tcgen05.ld`, `UTCBAR` = `tcgen05.commit`, `UTCATOMSWS` = TMEM allocation, "
       "`UBLKCP` = `cp.async.bulk` (both directions), `SYNCS` = mbarrier operations.", ""]
for kern, inst in WANT:
    cands = [n for n in funcs if f"{len(kern)}{kern}" in n and inst in n]
    if not cands:
        out.append(f"## `{kern}{inst}`: not found\n")
        continue
    n = sorted(cands, key=len)[0]
    body = funcs[n]
    cnt = collections.Counter()
    for l in body:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            op = m.group(1)
            for k in KEY:
                if op.startswith(k):
                    cnt[k] += 1
                    break
    out.append(f"## `{kern}` ({n})\n")
    out.append(f"instructions: {len(body)}; " + ", ".join(f"{k} {v}" for k, v in sorted(cnt.items())) + "\n")
    out.append("```")
    seen = set()
    for i, l in enumerate(body):
        for k in SHOW:
            if k in l and k not in seen:
                seen.add(k)
                out.extend(x[:130] for x in body[max(0, i - 1):i + 2])
                out.append("...")
    out.append("```\n")
open(os.path.join(ROOT, "profiles", f"{R}_sass_excerpt.md"), "w").write("\n".join(out) + "\n")
print("\n".join(o for o in out if o.startswith("##") or o.startswith("instructions")))
