mkdir -p gpurun_out
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t13_train1.log 2>&1; tail -1 gpurun_out/t13_train1.log | cut -c1-400
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t13_train2.log 2>&1; tail -1 gpurun_out/t13_train2.log | cut -c1-400
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reg > gpurun_out/t13_train1_noreg.log 2>&1; tail -1 gpurun_out/t13_train1_noreg.log | cut -c1-300
