mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_modules.py tests/test_gpu_round2.py tests/test_gpu_ragged.py -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; tail -4 gpurun_out/r2n_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_train.log 2>&1; tail -c 420 gpurun_out/r2n_train.log
python bench.py --mode forward --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_fwd.log 2>&1; tail -c 200 gpurun_out/r2n_fwd.log
# memcheck of the small-shape kernel tests (race / memory checking)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backward.py tests/test_gpu_regularise.py -m gpu -x -q -k "not 9536 and not 64-149 and not persistent and not full_size and not large" > gpurun_out/r2n_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2n_sanitizer_memcheck.log
