#!/bin/bash
# Round-end validation and evidence on ONE B200 (gpurun): GPU tests, smoke, bench (with the CPU leg), launch list of the
# bench command, dram traffic of every trailing-update launch of one evaluation, one --set full capture of the first.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.txt 2>&1; tail -3 gpurun_out/final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1; tail -2 gpurun_out/final_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_ours.json 2> gpurun_out/final_bench_ours.err; tail -c 3000 gpurun_out/final_bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/final_launch_list.csv python bench.py --no-cpu --steps 2 --warmup 1 > gpurun_out/final_launch_bench.log 2>&1
python tools/launch_summary.py gpurun_out/final_launch_list.csv > gpurun_out/final_launch_summary.txt 2>&1; cat gpurun_out/final_launch_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_update_kernel --csv --log-file gpurun_out/final_update_traffic.csv python tools/perf_probe.py 16384 0 1 > gpurun_out/final_update_traffic.log 2>&1
python tools/traffic_summary.py gpurun_out/final_update_traffic.csv > gpurun_out/final_update_traffic.txt 2>&1; cat gpurun_out/final_update_traffic.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_update_kernel -s 8 -c 1 -f -o gpurun_out/r01_final_update python tools/perf_probe.py 16384 0 1 > gpurun_out/final_ncu_update.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
