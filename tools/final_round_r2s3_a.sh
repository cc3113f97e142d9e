#!/bin/bash
# Round-2, third session, call A on ONE B200 (gpurun -- bash tools/final_round_r2s3_a.sh): the heteroscedastic VarDTC diagnostic
# (every error norm printed), the whole GPU test suite, smoke, a short bench line without the CPU leg.
set -x
mkdir -p gpurun_out
timeout 300 python tools/sparse_het_check.py 262144 4096 > gpurun_out/r2s3_sparse_het_check.txt 2>&1; tail -12 gpurun_out/r2s3_sparse_het_check.txt
timeout 900 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2s3_pytest_gpu.txt 2>&1; tail -30 gpurun_out/r2s3_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s3_smoke.txt 2>&1; tail -2 gpurun_out/r2s3_smoke.txt
timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/r2s3_bench_nocpu.json 2> gpurun_out/r2s3_bench_nocpu.err; tail -c 1800 gpurun_out/r2s3_bench_nocpu.json
