"""profiles/<tag>_scaling.md from the gpurun_out/scale_n*.log lines of tools/scale_run.sh: python tools/scaling_md.py r02"""
import json, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(root, "gpurun_out")


def rd(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        return None
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        return None


L = [f"# Data-parallel scaling, round {tag[1:].lstrip('0')} (one box, launched like the driver: `python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
     "--master-addr 127.0.0.1 ... bench.py --gpus N --steps 30 --warmup 5`)\n",
     "B = 64 + 64 clouds per GPU (weak scaling); the step is ONE CUDA graph per rank with the NCCL collectives inside; efficiency = "
     "clouds/s at N divided by N x the N = 1 run on the SAME box in the same `gpurun` call (`tools/scale_run.sh`).\n",
     "| N | MMD scope | ms / step | clouds/s | e2e clouds/s | efficiency | same-box N = 1 ms | replicas_in_sync |", "|---|---|---|---|---|---|---|---|"]
for n in (2, 4, 8):
    b = rd(f"scale_n1_on{n}.log")
    for scope in ("local", "global"):
        d = rd(f"scale_n{n}_{scope}.log")
        if d and b:
            L.append(f"| {n} | {scope} | {d['ms_per_step']:.3f} | {d['value']:.0f} | {d['e2e']['value']:.0f} | {d['value'] / (n * b['value']):.4f} | {b['ms_per_step']:.3f} | {d.get('replicas_in_sync')} |")
L.append("\n`local` = the reference's DDP intent (every rank evaluates the MMD on its own shard, train_dg.py:357-368): the gradient all-reduce is "
         "the only collective, issued group by group from gradient hooks and overlapped with the backward.  `global` = north_star's variant: "
         "one packed all-gather of the sub-domain features per MMD call, every rank evaluates the m = 64 N MMD.")
open(os.path.join(root, "profiles", f"{tag}_scaling.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L[3:]))
