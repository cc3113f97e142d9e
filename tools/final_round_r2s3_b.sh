#!/bin/bash
# Round-2, third session, call B on ONE B200 (gpurun -- bash tools/final_round_r2s3_b.sh): the evidence of the final state —
# GPU tests, bench line with the CPU leg, ncu launch list of the bench command, one --set full capture of the dominant kernel and
# its per-launch DRAM traffic, configs[2] (full optimize() with the CPU oracle at the final theta), configs[4] (sparse line),
# the metric table, timelines. Ordered by priority: the call may be cut by the remaining GPU budget.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2s3_final_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2s3_final_pytest_gpu.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s3_final_bench_ours.json 2> gpurun_out/r2s3_final_bench_ours.err; tail -c 2200 gpurun_out/r2s3_final_bench_ours.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2s3_final_launch_list.csv python bench.py --no-cpu --steps 2 --warmup 1 > gpurun_out/r2s3_final_launch_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2s3_final_launch_list.csv > gpurun_out/r2s3_final_launch_summary.txt 2>&1; head -30 gpurun_out/r2s3_final_launch_summary.txt
timeout 300 python tools/optimize_probe.py 16384 32 40 iso cpu > gpurun_out/r2s3_final_optimize_config3.json 2> gpurun_out/r2s3_final_optimize_config3.err; tail -c 1500 gpurun_out/r2s3_final_optimize_config3.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:oz_gemm -s 49 -c 1 -f -o gpurun_out/r2s3_final_oz_gemm2 python tools/one_eval.py 16384 2 > gpurun_out/r2s3_final_ncu_oz.log 2>&1; tail -2 gpurun_out/r2s3_final_ncu_oz.log
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:oz_gemm --csv --log-file gpurun_out/r2s3_final_oz_traffic.csv python tools/one_eval.py 16384 2 > gpurun_out/r2s3_final_oz_traffic.log 2>&1
python tools/traffic_summary.py gpurun_out/r2s3_final_oz_traffic.csv > gpurun_out/r2s3_final_oz_traffic.txt 2>&1; head -8 gpurun_out/r2s3_final_oz_traffic.txt
timeout 200 python bench.py --workload sparse --no-cpu --steps 3 --warmup 2 > gpurun_out/r2s3_final_bench_sparse.json 2> gpurun_out/r2s3_final_bench_sparse.err; tail -c 1200 gpurun_out/r2s3_final_bench_sparse.json
timeout 240 python tools/metric_table.py 256,512,1024,2048,4096,8192,16384,32768,65536 512,4096 > gpurun_out/r2s3_final_metric_table.json 2> gpurun_out/r2s3_final_metric_table.err; cut -c1-260 gpurun_out/r2s3_final_metric_table.json
timeout 120 python tools/timeline.py 16384 gpurun_out/r2s3_final_timeline_16384.txt > /dev/null 2>&1
timeout 120 python tools/timeline.py 4096 gpurun_out/r2s3_final_timeline_4096.txt > /dev/null 2>&1
timeout 120 python tools/timeline.py 512 gpurun_out/r2s3_final_timeline_512.txt > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
