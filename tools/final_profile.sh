#!/bin/bash
# Round-end measurement pass on the GPU box: tests, bench lines, ncu launch list, ncu --set full captures of
# the top kernels (each capture only after the same command has run without ncu), micro-benchmark sweep.
# Everything lands in gpurun_out/; tools/summarize_ncu.py turns it into profiles/<tag>_*.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
timeout 1500 bash tools/gpu_check.sh > gpurun_out/full_check.log 2>&1; tail -3 gpurun_out/full_check.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --share-trunk > gpurun_out/bench_share.log 2>&1; echo "share rc=$?"
timeout 200 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 100 python tools/gemm_only.py 65536 512 128 > gpurun_out/gemm_only.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -s 2 -o gpurun_out/prof_gemm_tc \
    python tools/gemm_only.py 65536 512 128 1 > gpurun_out/ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 100 python tools/knn_only.py 64 20 > gpurun_out/knn_only.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 1 -s 2 -o gpurun_out/prof_knn_tc \
    python tools/knn_only.py 64 20 1024 64 1 > gpurun_out/ncu_knn.log 2>&1; echo "knn rc=$?"
timeout 100 python tools/edge_only.py 64 128 > gpurun_out/edge_only.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"edge_gather_smem|edge_bwd_main|edge_bwd_pre" -c 3 -s 6 \
    -o gpurun_out/prof_edge python tools/edge_only.py 64 128 1 > gpurun_out/ncu_edge.log 2>&1; echo "edge rc=$?"
timeout 400 python tools/microbench.py > gpurun_out/microbench.md 2> gpurun_out/microbench.err; echo "microbench rc=$?"
cat gpurun_out/gemm_only.log gpurun_out/knn_only.log gpurun_out/edge_only.log
