#!/bin/bash
# Round-2 (second session) end validation and evidence on ONE B200 (gpurun -- bash tools/final_round_r2s2.sh): GPU tests, smoke,
# bench with the CPU leg, ncu launch list of the bench command, per-launch DRAM traffic of the tcgen05 GEMM, one --set full
# capture of its largest launch and one of the fine DMMA GEMM, the sparse workload line, the metric table.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s2_final_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2s2_final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s2_final_smoke.txt 2>&1; tail -2 gpurun_out/r2s2_final_smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s2_final_bench_ours.json 2> gpurun_out/r2s2_final_bench_ours.err; tail -c 2500 gpurun_out/r2s2_final_bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2s2_final_launch_list.csv python bench.py --no-cpu --steps 2 --warmup 1 > gpurun_out/r2s2_final_launch_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2s2_final_launch_list.csv > gpurun_out/r2s2_final_launch_summary.txt 2>&1; cat gpurun_out/r2s2_final_launch_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:oz_gemm --csv --log-file gpurun_out/r2s2_final_oz_traffic.csv python tools/one_eval.py 16384 2 > gpurun_out/r2s2_final_oz_traffic.log 2>&1
python tools/traffic_summary.py gpurun_out/r2s2_final_oz_traffic.csv > gpurun_out/r2s2_final_oz_traffic.txt 2>&1; head -8 gpurun_out/r2s2_final_oz_traffic.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm -s 49 -c 1 -f -o gpurun_out/r2s2_final_oz_gemm2 python tools/one_eval.py 16384 2 > gpurun_out/r2s2_final_ncu_oz.log 2>&1; tail -2 gpurun_out/r2s2_final_ncu_oz.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_fine -s 60 -c 1 -f -o gpurun_out/r2s2_final_gemm_fine python tools/one_eval.py 4096 2 > gpurun_out/r2s2_final_ncu_fine.log 2>&1; tail -2 gpurun_out/r2s2_final_ncu_fine.log
timeout 600 python bench.py --workload sparse --steps 3 --warmup 2 > gpurun_out/r2s2_final_bench_sparse.json 2> gpurun_out/r2s2_final_bench_sparse.err; tail -c 1500 gpurun_out/r2s2_final_bench_sparse.json; tail -2 gpurun_out/r2s2_final_bench_sparse.err
timeout 900 python tools/metric_table.py 256,512,1024,2048,4096,8192,16384,32768,65536 512,4096 > gpurun_out/r2s2_final_metric_table.json 2> gpurun_out/r2s2_final_metric_table.err; cut -c1-300 gpurun_out/r2s2_final_metric_table.json
timeout 300 python tools/timeline.py 16384 gpurun_out/r2s2_final_timeline_16384.txt > /dev/null 2>&1
timeout 300 python tools/timeline.py 4096 gpurun_out/r2s2_final_timeline_4096.txt > /dev/null 2>&1
timeout 300 python tools/timeline.py 512 gpurun_out/r2s2_final_timeline_512.txt > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
