#!/bin/bash
# Scaling runs on ONE box, launched exactly like the driver does: bash tools/scale_run.sh <N> [extra bench args]
# (also runs --gpus 1 on the same box first, so that the efficiency has a same-box denominator).
N=$1; shift
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-torch-reference --no-serial-roofline > gpurun_out/scale_n1_on${N}.log 2> gpurun_out/scale_n1_on${N}.err
for scope in local global; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --mmd-scope $scope "$@" > gpurun_out/scale_n${N}_${scope}.log 2> gpurun_out/scale_n${N}_${scope}.err
  echo "N=$N scope=$scope rc=$?"
done
python - <<PY
import json
def rd(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
b = rd("gpurun_out/scale_n1_on${N}.log")
print("N=1 on this box:", b and (round(b["ms_per_step"], 3), round(b["value"], 1)))
for s in ("local", "global"):
    d = rd("gpurun_out/scale_n${N}_%s.log" % s)
    if d and b:
        print("N=${N}", s, round(d["ms_per_step"], 3), "ms", round(d["value"], 1), "clouds/s  efficiency vs same-box N=1:", round(d["value"] / (${N} * b["value"]), 4),
              "in_sync", d.get("replicas_in_sync"), d["execution"]["mode"], "e2e", round(d["e2e"]["value"], 1))
    else:
        print("N=${N}", s, "no line")
PY
