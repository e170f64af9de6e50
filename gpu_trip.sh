mkdir -p gpurun_out
T0=$(date +%s)
NG=$(nvidia-smi -L | wc -l)
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t46_train$NG.log 2>&1; tail -1 gpurun_out/t46_train$NG.log | cut -c1-330
echo "train$NG done $(( $(date +%s) - T0 )) s"
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $NG --mode forward --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t46_fwd$NG.log 2>&1; tail -1 gpurun_out/t46_fwd$NG.log | cut -c1-330
echo "all done $(( $(date +%s) - T0 )) s"
