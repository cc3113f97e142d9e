#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -4 gpurun_out/pytest_gpu.txt
timeout 200 python tools/perf_probe.py 4096,16384 0 3 > gpurun_out/perf6.txt 2>&1; grep "^N=" gpurun_out/perf6.txt
GPX_NO_PERSIST=1 timeout 200 python tools/perf_probe.py 16384 0 3 > gpurun_out/perf6_nopersist.txt 2>&1; grep "^N=" gpurun_out/perf6_nopersist.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 300 python tools/optimize_probe.py 16384 32 25 > gpurun_out/optimize_m52_d32.json 2>&1; cat gpurun_out/optimize_m52_d32.json
timeout 300 python tools/sparse_blocks_probe.py > gpurun_out/sparse_blocks.json 2>&1; cat gpurun_out/sparse_blocks.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_update_persist_kernel -s 1 -c 1 -f -o gpurun_out/r01_update_persist python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_upd2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kbuild_kernel -c 1 -f -o gpurun_out/r01_kbuild2 python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_kb2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_lauum_kernel -c 1 -f -o gpurun_out/r01_lauum2 python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_lauum2.log 2>&1
ls -la gpurun_out/*.ncu-rep
