#!/bin/bash
# ncu --set full captures of the four heavy kernels at N=16384 (one launch each), reports under gpurun_out/
set -x
mkdir -p gpurun_out
for spec in "gemm_update_kernel 14 upd" "gemm_lauum_kernel 0 lauum" "gemm_panel_kernel 8 panel" "base_sweep16_kernel 3 base" "kbuild_kernel 0 kbuild"; do
  set -- $spec
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/r01_$3 python tools/perf_probe.py 16384 0 1 > gpurun_out/ncu_$3.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
