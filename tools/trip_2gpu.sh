mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "two_rank" -s > gpurun_out/r2j_test_2gpu.log 2>&1; tail -6 gpurun_out/r2j_test_2gpu.log
run2() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_n2_$name.log 2>&1
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2j_n2_$name.log | head -1 | sed "s/^/$name /"; grep -o "NVLS[^\"]*" gpurun_out/r2j_n2_$name.log | head -2; }
run2 nvls X=1
run2 nccl W2V2_NVLS=0
run2 nvls32 W2V2_NVLS_CTAS=32
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_n1.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2j_n1.log | head -1
