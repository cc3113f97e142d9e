mkdir -p gpurun_out
timeout 900 python tools/panel_ab.py 1300,4096,16384 5 > gpurun_out/s2_panel_ab2.txt 2>&1; cut -c1-230 gpurun_out/s2_panel_ab2.txt
timeout 300 python tools/timeline.py 16384 gpurun_out/s2_timeline9_16384.txt > /dev/null 2>&1; sed -n 1,22p gpurun_out/s2_timeline9_16384.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest_gpu10.txt 2>&1; tail -5 gpurun_out/s2_pytest_gpu10.txt
