#!/bin/bash
# gpurun --gpus G -- tools/multi_gpu_round.sh G    : multi-GPU evidence of the round on one box
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
# 0. the 2-GPU tests of the suite (sharded evaluation / prediction / gathered L and the row-sharded sparse model against the oracle)
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2_pytest_dist_${G}gpu.txt 2>&1; tail -3 gpurun_out/r2_pytest_dist_${G}gpu.txt
# 1. the bench line the driver's SCALE run produces (sharded N=16384), NCCL lines on stderr
NCCL_DEBUG=VERSION $TR --master-port 29551 bench.py --gpus $G --steps 3 --warmup 3 > gpurun_out/r2_bench_${G}gpu.json 2> gpurun_out/r2_bench_${G}gpu.err
cut -c1-600 gpurun_out/r2_bench_${G}gpu.json; grep -o '"parity_vs_1gpu": {[^}]*}' gpurun_out/r2_bench_${G}gpu.json
# 2. parity vs oracle (small) and vs the single-GPU engine (large), sharded predict and gathered L
SIZES="3000,16384"; if [ "$G" -ge 8 ]; then SIZES="3000,65536"; fi
GPX_DIST_VERIFY=1 $TR --master-port 29552 tools/dist_check.py $SIZES > gpurun_out/r2_dist_check_${G}gpu.txt 2>&1; grep "^G=" gpurun_out/r2_dist_check_${G}gpu.txt
# 3. beyond one GPU's workspace: N = 131072 (8 GPUs only): 2 x 17 GB per rank instead of 137 GB
if [ "$G" -ge 8 ]; then
  $TR --master-port 29553 tools/dist_check.py 131072 > gpurun_out/r2_dist_n131072_${G}gpu.txt 2>&1; grep "^G=" gpurun_out/r2_dist_n131072_${G}gpu.txt; nvidia-smi --query-gpu=index,memory.used --format=csv | head -3
fi
