mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest_gpu.txt 2>&1; tail -3 gpurun_out/s2_pytest_gpu.txt
timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/s2_bench_nocpu.json 2> gpurun_out/s2_bench_nocpu.err; tail -c 1800 gpurun_out/s2_bench_nocpu.json
timeout 300 python tools/metric_table.py 512,4096,8192 512 > gpurun_out/s2_metric_table.json 2> gpurun_out/s2_metric_table.err; cut -c1-1500 gpurun_out/s2_metric_table.json
