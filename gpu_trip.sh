mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -5
timeout -k 5 200 python tools/gemm_bench.py > gpurun_out/t5_gemm.log 2>&1; cat gpurun_out/t5_gemm.log
timeout -k 5 300 python tools/time_ops.py 64 > gpurun_out/t5_time.log 2>&1; cat gpurun_out/t5_time.log
timeout -k 5 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_modules.py -m gpu -q 2>&1 | tail -8
ITERS=1 timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof5_ffn1 python tools/gemm_bench.py ffn1 > gpurun_out/t5_ncu1.log 2>&1; tail -1 gpurun_out/t5_ncu1.log
ITERS=1 timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof5_qkv python tools/gemm_bench.py qkv > gpurun_out/t5_ncu2.log 2>&1; tail -1 gpurun_out/t5_ncu2.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/prof5_attn python tools/op_bench.py attention > gpurun_out/t5_ncu3.log 2>&1; tail -1 gpurun_out/t5_ncu3.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:posconv_kernel -s 3 -c 1 -o gpurun_out/prof5_posconv python tools/op_bench.py posconv > gpurun_out/t5_ncu4.log 2>&1; tail -1 gpurun_out/t5_ncu4.log
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:conv0_apply -s 3 -c 1 -o gpurun_out/prof5_conv0 python tools/op_bench.py conv0 > gpurun_out/t5_ncu5.log 2>&1; tail -1 gpurun_out/t5_ncu5.log
