"""Summarise ncu outputs from gpurun_out/ into profiles/ (tracked).  Usage: summarize_ncu.py <round-tag>"""
import csv, sys, os, subprocess, collections, json
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles"); os.makedirs(out, exist_ok=True)
src = os.path.join(root, "gpurun_out", "launches.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: t = float(r[vi].replace(",", ""))
        except ValueError: continue
        name = r[ki].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out, f"{tag}_launch_summary.csv"), "w") as f:
        f.write("kernel,launches,total_ns,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.0f},{t / tot:.4f}\n")
    print("launch summary:", len(agg), "kernels, total", tot / 1e6, "ms over", sum(a[0] for a in agg.values()), "launches")
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'launch__shared_mem_per_block_dynamic']
res = {}
for f in sorted(os.listdir(os.path.join(root, "gpurun_out"))):
    if not f.endswith(".ncu-rep"): continue
    p = subprocess.run(["ncu", "-i", os.path.join(root, "gpurun_out", f), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(p.stdout.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:100]}
        for w in want:
            if w in hdr: d[w + (f" [{units[hdr.index(w)]}]" if units[hdr.index(w)] else "")] = r[hdr.index(w)]
        res.setdefault(f, []).append(d)
json.dump(res, open(os.path.join(out, f"{tag}_ncu_full_top_kernels.json"), "w"), indent=1)
print("wrote", os.listdir(out))
