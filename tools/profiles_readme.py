"""Write profiles/README.md for a round from the files tools/summarize_ncu.py produced: python tools/profiles_readme.py r02"""
import json, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(root, "profiles")
bench = json.loads(open(os.path.join(P, f"{tag}_bench_line.json")).read().strip().splitlines()[-1])
ref = json.loads(open(os.path.join(P, f"{tag}_bench_reference_arm.json")).read().strip().splitlines()[-1])
ncu = json.load(open(os.path.join(P, f"{tag}_ncu_full_top_kernels.json")))
traffic = json.load(open(os.path.join(P, "ncu_traffic.json")))


def g(d, key):
    k = next((x for x in d if x.startswith(key)), None)
    return d[k] if k else "-"


L = []
L.append(f"# profiles/ -- measured on B200 (sm_100a) through `gpurun`, round {tag[1:].lstrip('0')}\n")
L.append("Produced by `tools/final_profile.sh` on the GPU box, condensed by `tools/summarize_ncu.py` and `tools/profiles_readme.py`;"
         " nothing here is edited by hand.  Round 1's page is `r01_README.md`.\n")
L.append("| file | what |\n|---|---|")
L.append(f"| `{tag}_bench_line.json` | the JSON line of `python bench.py --steps 20 --warmup 5` (N = 1, CUDA-graph step; with `roofline`, `roofline_by_class`, `peaks`, `torch_gpu_reference`, `cpu_baseline`) |")
L.append(f"| `{tag}_bench_reference_arm.json` | `python bench.py --impl reference --steps 3 --warmup 1` (the reference algorithm on the host cores, same config) |")
L.append(f"| `{tag}_launch_summary.csv` | per-kernel launch list of eager steps (`ncu --metrics gpu__time_duration.sum --clock-control none -c 4000`) |")
L.append(f"| `{tag}_ncu_full_top_kernels.json`, `ncu_traffic.json` | `ncu --set full --clock-control none --import-source on` of one launch of each top kernel at the step's shapes; DRAM bytes per launch per class (what `bench.py` reports as `roofline.traffic`) |")
L.append(f"| `{tag}_microbench.md` | BASELINE.json configs[3] sweep (N, k, C) + configs[4] LiDAR-scale inference |")
L.append(f"| `{tag}_scaling.md`, `{tag}_dp_parity_2gpu.log`, `{tag}_nccl_graph_probe_2gpu.log` | 1 / 2 / 4 / 8 GPUs launched like the driver does (both MMD scopes); N-rank step vs single-process step; NCCL inside a CUDA graph |")
L.append(f"| `{tag}_mmd_gradient_bound.md` | which parameters the reference's fp32 MMD-gradient noise touches and by how much (64+64 step) |")
L.append(f"| `{tag}_sass_opcodes.md` | SASS opcode histogram per translation unit (UTCHMMA / UTMALDG / LDTM / STTM / FMNMX3 ...) |")
L.append(f"| `{tag}_sanitizer.md` | compute-sanitizer is closed on this pool; what stands in for it |\n")
L.append(f"## Headline (N = 1, `{tag}_bench_line.json`)\n")
r = bench["roofline"]
tg = bench.get("torch_gpu_reference", {})
L.append(f"* {bench['value']:.0f} clouds/s, {bench['ms_per_step']:.2f} ms per SUG step; end to end from pinned host memory with the loss read back: "
         f"{bench['e2e']['value']:.0f} clouds/s ({bench['e2e']['ms_per_step']:.2f} ms; {bench['e2e']['h2d_bytes_per_step']} B in, {bench['e2e']['d2h_bytes_per_step']} B out per step); "
         f"{bench['gpu_launches'] // bench['steps']} kernel launches per step; clocks {bench['clocks']}.")
if "tf32_default" in tg:
    L.append(f"* The reference ALGORITHM as plain PyTorch ops on the same GPU, same run: {tg['tf32_default']['ms_per_step']:.1f} ms (cuDNN TF32, PyTorch's default) / "
             f"{tg['fp32']['ms_per_step']:.1f} ms (fp32) per step -> this library is {tg['speedup_of_this_library']['vs_tf32_default']:.1f}x / {tg['speedup_of_this_library']['vs_fp32']:.1f}x faster (north_star target: >= 3x).")
L.append(f"* CPU: `--impl reference` {ref['value']:.1f} clouds/s ({ref['cpu_baseline']['sample']}); in-line `cpu_baseline` {bench['cpu_baseline']['value']:.1f} clouds/s ({bench['cpu_baseline']['sample']}).")
L.append(f"* Dominant class `{r['kernel']}`: {r['launches_timed']} launches, {r['avg_launch_us']:.1f} us average, {r['share_of_step'] * 100:.1f} % of the step; bound = {r['bound']}: "
         f"{r['achieved']:.1f} {r['unit']} of {r['peak']:.1f} = **{r['frac']:.3f}** ({r['peak_source']}); {r['gbps']:.0f} GB/s = {r['frac_of_hbm']:.3f} of the HBM copy peak; "
         f"DRAM traffic of the captured launch {r['traffic'] / 1e6 if r['traffic'] else float('nan'):.1f} MB ({(r.get('traffic_capture') or {}).get('shape', '')}).")
L.append(f"* Peaks: {bench['peaks']}.\n")
L.append("## Every kernel class of the step (`roofline_by_class`: instrumented eager step, algorithmic bytes / flops)\n")
L.append("| class | launches / step | avg us | ms / step | bound | achieved | fraction of that roofline | of HBM copy peak | of 3xTF32 ceiling |\n|---|---|---|---|---|---|---|---|---|")
for k, v in sorted(bench["roofline_by_class"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    L.append(f"| {k} | {v['launches_per_step']} | {v['avg_launch_us']:.1f} | {v['ms_per_step']:.3f} | {v['bound']} | {v['achieved']:.1f} {v['unit']} | {v['frac']:.3f} | {v['frac_of_hbm']:.3f} | "
             f"{v.get('frac_of_3xtf32_ceiling', '-')} |")
L.append("\n## ncu --set full, one launch of each top kernel\n")
L.append("| capture | kernel | time us | DRAM read + write MB | tensor pipe % | issue slots % | warps active % | regs | smem KB | L2 hit % |\n|---|---|---|---|---|---|---|---|---|---|")
for f, launches in ncu.items():
    for d in launches:
        t = g(d, "gpu__time_duration.sum")
        L.append(f"| {f} | `{d['kernel'][:60]}` | {t} | {g(d, 'dram__bytes_read.sum')} + {g(d, 'dram__bytes_write.sum')} | {g(d, 'sm__pipe_tensor_cycles_active')} | "
                 f"{g(d, 'smsp__issue_active')} | {g(d, 'sm__warps_active')} | {g(d, 'launch__registers_per_thread')} | {g(d, 'launch__shared_mem_per_block_dynamic')} | {g(d, 'lts__t_sector_hit_rate')} |")
L.append("\nDRAM bytes per launch used for `roofline.traffic`: " + ", ".join(f"{k} {v['dram_bytes_per_launch'] / 1e6:.1f} MB" for k, v in traffic.items()) + ".\n")
extra = os.path.join(P, f"{tag}_notes.md")
if os.path.exists(extra):
    L.append(open(extra).read())
open(os.path.join(P, "README.md"), "w").write("\n".join(L) + "\n")
print("wrote profiles/README.md,", len(L), "blocks")
