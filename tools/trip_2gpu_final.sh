#!/bin/bash
# two ranks on hardware: the N-rank step equals the 1-rank step, then one N=2 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "two_rank" -s > gpurun_out/r2s_test_2gpu.log 2>&1; tail -4 gpurun_out/r2s_test_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_n2.log 2>&1
grep -a '^{' gpurun_out/r2s_n2.log | tail -1 | cut -c1-700
