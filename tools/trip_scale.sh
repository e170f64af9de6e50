# 8-GPU trip: train bench at N = 1 and N = 8, NVLS all-reduce kernel vs NCCL
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_n1.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_n1.log | head -1 | sed "s/^/n1 /"
run8() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_n8_$name.log 2>&1
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_n8_$name.log | head -1 | sed "s/^/$name /"; grep -o "dp8 ([^)]*)" gpurun_out/r2l_n8_$name.log | head -1
}
run8 nvls128 X=1
run8 nvls64 W2V2_NVLS_CTAS=64
run8 nvls256 W2V2_NVLS_CTAS=256
run8 nccl W2V2_NVLS=0
run8 nvls_uniform4 W2V2_AR_LAYERS=4
run8 nvls_sched1 W2V2_AR_LAYERS=1
env timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_n4_nvls.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_n4_nvls.log | head -1 | sed "s/^/n4 /"
env timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_n2_nvls.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_n2_nvls.log | head -1 | sed "s/^/n2 /"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/r2l_n8_cfg4.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_n8_cfg4.log | head -1 | sed "s/^/cfg4 /"
