mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 5 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_regularise.py tests/test_gpu_training.py -m gpu -x -q > gpurun_out/t18_tests.log 2>&1; tail -5 gpurun_out/t18_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t18_train1.log 2>&1; tail -1 gpurun_out/t18_train1.log | cut -c1-300
echo "bench done $(( $(date +%s) - T0 )) s"
timeout -k 5 300 ncu --set full --import-source on --clock-control none -k regex:attention_bwd -s 13 -c 1 -o gpurun_out/t18_attn_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t18_ncu1.log 2>&1; tail -2 gpurun_out/t18_ncu1.log | cut -c1-200
echo "all done $(( $(date +%s) - T0 )) s"
