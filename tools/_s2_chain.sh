mkdir -p gpurun_out
CHAIN_AB_SHORT=1 timeout 900 python tools/chain_ab.py 256,512,640 7 > gpurun_out/s2_chain_ab7.txt 2>&1; cut -c1-230 gpurun_out/s2_chain_ab7.txt
timeout 300 python tools/timeline.py 512 gpurun_out/s2_timeline7_512.txt > /dev/null 2>&1; sed -n 20,50p gpurun_out/s2_timeline7_512.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest_gpu8.txt 2>&1; tail -5 gpurun_out/s2_pytest_gpu8.txt
timeout 300 python tools/metric_table.py 512,4096,8192,16384,32768,65536 512,4096 > gpurun_out/s2_metric_table2.json 2> gpurun_out/s2_metric_table2.err; cut -c1-1200 gpurun_out/s2_metric_table2.json
