#!/bin/bash
# Round-end measurement pass on the GPU box (run through gpurun): bench lines, ncu launch list, ncu --set full captures
# of the top kernels (each capture only after the same command has exited 0 without ncu), micro-benchmark sweep.
# Everything lands in gpurun_out/; tools/summarize_ncu.py <tag> turns it into profiles/<tag>_*.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log 2> gpurun_out/bench_ref_final.err; echo "reference arm rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-reference --share-trunk > gpurun_out/bench_share.log 2>&1; echo "share rc=$?"
timeout 200 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-torch-reference > gpurun_out/bench_eager.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-torch-reference > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 100 python tools/gemm_only.py 65536 512 128 > gpurun_out/gemm_only.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -s 2 -o gpurun_out/prof_gemm_tc \
    python tools/gemm_only.py 65536 512 128 1 > gpurun_out/ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 100 python tools/gemm_only.py 65536 512 512 > gpurun_out/gemm_only2.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -s 2 -o gpurun_out/prof_gemm_tc_k512 \
    python tools/gemm_only.py 65536 512 512 1 > gpurun_out/ncu_gemm2.log 2>&1; echo "gemm k512 rc=$?"
timeout 100 python tools/knn_only.py 64 20 > gpurun_out/knn_only.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 1 -s 2 -o gpurun_out/prof_knn_tc \
    python tools/knn_only.py 64 20 1024 64 1 > gpurun_out/ncu_knn.log 2>&1; echo "knn rc=$?"
timeout 100 python tools/knn_only.py 3 20 > gpurun_out/knn_xyz_only.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_xyz_kernel -c 1 -s 2 -o gpurun_out/prof_knn_xyz \
    python tools/knn_only.py 3 20 1024 64 1 > gpurun_out/ncu_knn_xyz.log 2>&1; echo "knn xyz rc=$?"
timeout 100 python tools/edge_only.py 64 128 > gpurun_out/edge_only.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"edge_gather_smem|edge_bwd_main_smem|edge_bwd_route|bn_act_kernel" -c 4 -s 8 \
    -o gpurun_out/prof_edge python tools/edge_only.py 64 128 1 > gpurun_out/ncu_edge.log 2>&1; echo "edge rc=$?"
timeout 500 python tools/microbench.py > gpurun_out/microbench.md 2> gpurun_out/microbench.err; echo "microbench rc=$?"
cat gpurun_out/gemm_only.log gpurun_out/gemm_only2.log gpurun_out/knn_only.log gpurun_out/knn_xyz_only.log gpurun_out/edge_only.log
