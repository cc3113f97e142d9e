#!/bin/bash
# One GPU-box session: tests, smoke, bench (both arms), ncu launch list and full captures. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/launches.csv python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_update_kernel -s 7 -c 1 -o gpurun_out/prof_update python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_update.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_lauum_kernel -c 1 -o gpurun_out/prof_lauum python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_lauum.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kbuild_kernel -c 1 -o gpurun_out/prof_kbuild python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_kbuild.log 2>&1
ls -la gpurun_out
