mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
P=29530
for L in 1 3 6; do
P=$((P+1))
W2V2_AR_LAYERS=$L timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $P bench.py --gpus $NG --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/t44_train${NG}_L$L.log 2>&1; echo "AR_LAYERS=$L: $(tail -1 gpurun_out/t44_train${NG}_L$L.log | cut -c60-190)"
done
