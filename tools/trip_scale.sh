# 8-GPU trip: 2-rank hardware equality test, then train bench at N = 1 and N = 8 with all-reduce schedule variants
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r2h_gpus.log
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "two_rank" > gpurun_out/r2h_test_2gpu.log 2>&1; tail -4 gpurun_out/r2h_test_2gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_n1.log 2>&1; tail -c 200 gpurun_out/r2h_n1.log
run8() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_n8_$name.log 2>&1
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2h_n8_$name.log | head -1 | sed "s/^/$name /"
}
run8 default X=1
run8 ctas8 NCCL_MAX_CTAS=8
run8 ctas16 NCCL_MAX_CTAS=16
run8 uniform4 W2V2_AR_LAYERS=4
run8 sched W2V2_AR_SCHEDULE=3,3,3,2,1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/r2h_n8_cfg4.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2h_n8_cfg4.log | head -1 | sed "s/^/cfg4 /"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg3 > gpurun_out/r2h_n8_cfg3.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2h_n8_cfg3.log | head -1 | sed "s/^/cfg3 /"
