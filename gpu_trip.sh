mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -5
timeout -k 5 300 python tools/time_ops.py 64 > gpurun_out/t6_time.log 2>&1; cat gpurun_out/t6_time.log
W2V2_POSCONV_U=2 timeout -k 5 100 python tools/op_bench.py posconv
W2V2_POSCONV_U=2 timeout -k 5 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k posconv 2>&1 | tail -2
timeout -k 5 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_modules.py -m gpu -q 2>&1 | tail -8
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -k 5 600 python bench.py --steps 20 --warmup 5 > gpurun_out/t6_bench.log 2>&1; tail -3 gpurun_out/t6_bench.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 0 -c 1 -o gpurun_out/prof6_conv0gemm python tools/op_bench.py conv0 > gpurun_out/t6_ncu1.log 2>&1; tail -1 gpurun_out/t6_ncu1.log
