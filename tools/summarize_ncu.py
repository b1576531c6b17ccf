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

# ---- DRAM traffic per launch of each class's representative capture: what bench.py reports as roofline.traffic --------
CLASS_OF = (("gemm_tc_kernel", "gemm_tc"), ("knn_tc_kernel", "knn_tc"), ("knn_xyz_kernel", "knn_simt"), ("edge_gather_smem", "edge_gather_fwd"),
            ("edge_bwd_route", "edge_bwd_pre"), ("edge_bwd_main_smem", "edge_bwd_main"), ("bn_act_kernel", "bn_act"))
SHAPE = {"prof_gemm_tc.ncu-rep": "gemm_tc 65536x512x128 (conv4 a|b GEMM: algorithmic 4(MK+NK+MN) = 168.0 MB)",
         "prof_gemm_tc_k512.ncu-rep": "gemm_tc 65536x512x512 (conv5: algorithmic 269.5 MB, 34.4 GF)",
         "prof_knn_tc.ncu-rep": "kNN B=64 N=1024 C=64 k=20 (algorithmic 22.0 MB, 8.59 GF)",
         "prof_knn_xyz.ncu-rep": "kNN B=64 N=1024 C=3 k=20 (algorithmic 6.0 MB)",
         "prof_edge.ncu-rep": "EdgeConv 64->128, B=64 N=1024 k=20"}
traffic = {}
for f, launches in res.items():
    for d in launches:
        cls = next((c for pat, c in CLASS_OF if pat in d["kernel"]), None)
        if cls is None or (cls in traffic and f != "prof_gemm_tc.ncu-rep"):
            continue
        def num(key):
            k = next((x for x in d if x.startswith(key)), None)
            if k is None:
                return 0.0
            v = float(d[k].replace(",", ""))
            unit = k[k.index("[") + 1:-1] if "[" in k else "byte"
            return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        tkey = next(x for x in d if x.startswith("gpu__time_duration.sum"))
        traffic[cls] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                        "kernel_time_us_under_ncu": float(d[tkey].replace(",", "")) * (1e-3 if "nsecond" in tkey else 1.0),
                        "kernel": d["kernel"], "capture": f"profiles/{tag}_ncu_full_top_kernels.json <- gpurun_out/{f}",
                        "shape": SHAPE.get(f, f)}
json.dump(traffic, open(os.path.join(out, "ncu_traffic.json"), "w"), indent=1)
print("ncu_traffic.json:", {k: round(v["dram_bytes_per_launch"] / 1e6, 1) for k, v in traffic.items()}, "MB")
