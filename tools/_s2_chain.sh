mkdir -p gpurun_out
CHAIN_AB_SHORT=1 CHAIN_AB_PROF=1 timeout 900 python tools/chain_ab.py 256,512,1300,2048,4096,8192,16384 7 > gpurun_out/s2_chain_ab4.txt 2>&1; cat gpurun_out/s2_chain_ab4.txt | cut -c1-260
timeout 300 python tools/timeline.py 4096 gpurun_out/s2_timeline4_4096.txt > /dev/null 2>&1; sed -n 36,70p gpurun_out/s2_timeline4_4096.txt
timeout 300 python tools/timeline.py 512 gpurun_out/s2_timeline4_512.txt > /dev/null 2>&1; sed -n 28,70p gpurun_out/s2_timeline4_512.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest_gpu5.txt 2>&1; tail -5 gpurun_out/s2_pytest_gpu5.txt
timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/s2_bench_nocpu3.json 2> gpurun_out/s2_bench_nocpu3.err; cut -c1-400 gpurun_out/s2_bench_nocpu3.json
