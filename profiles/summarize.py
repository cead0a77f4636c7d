#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/<tag>_summary.md.
usage: python profiles/summarize.py <tag>   (expects gpurun_out/<tag>_launches.csv, <tag>_scatter.ncu-rep, <tag>_join.ncu-rep)"""
import collections, csv, io, os, re, subprocess, sys

tag = sys.argv[1]
G = "gpurun_out"
out = [f"# ncu summary `{tag}`  (command: python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu)\n"]

p = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        m = re.search(r"(\w+)(<[^(]*>)?\(", r[ki]); name = (m.group(1) + (m.group(2) or "")) if m else r[ki][:50]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out.append("## launch list (gpu__time_duration.sum, serialised, cold cache: compare shares)\n")
    out.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {v / 1e6:.3f} | {v / tot * 100:.1f}% |")
    out.append("")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum"]
for kern in ("scatter", "join"):
    rep = os.path.join(G, f"{tag}_{kern}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        out.append(f"## `{r[h.index('Kernel Name')][:80]}` (ncu --set full, {os.path.basename(rep)})\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in h:
                out.append(f"| {w} | {r[h.index(w)]} | {units[h.index(w)]} |")
        st = []
        for i, n in enumerate(h):
            if "pcsamp_warps_issue_stalled" in n and not n.endswith("_not_issued") and r[i]:
                try: st.append((float(r[i].replace(",", "")), n.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError: pass
        st.sort(reverse=True); tot = sum(v for v, _ in st) or 1
        out.append("\nstall samples: " + ", ".join(f"{n} {v / tot * 100:.0f}%" for v, n in st[:8]) + "\n")
open(os.path.join("profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
