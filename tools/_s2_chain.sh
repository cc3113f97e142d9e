mkdir -p gpurun_out
timeout 900 python tools/panel_ab.py 1300,4096,8192,16384,32768 5 > gpurun_out/s2_panel_ab.txt 2>&1; cut -c1-230 gpurun_out/s2_panel_ab.txt
timeout 300 python tools/timeline.py 16384 gpurun_out/s2_timeline8_16384.txt > /dev/null 2>&1; sed -n 1,34p gpurun_out/s2_timeline8_16384.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest_gpu9.txt 2>&1; tail -5 gpurun_out/s2_pytest_gpu9.txt
timeout 600 python tools/ozaki_check.py full > gpurun_out/s2_ozaki_check2.txt 2>&1; tail -8 gpurun_out/s2_ozaki_check2.txt | cut -c1-200
