mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ragged.py tests/test_gpu_round2.py tests/test_gpu_modules.py tests/test_gpu_engine.py -m gpu -q > gpurun_out/r2k_tests_new.log 2>&1; tail -12 gpurun_out/r2k_tests_new.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_ragged.py --deselect tests/test_gpu_round2.py > gpurun_out/r2k_tests_all.log 2>&1; tail -4 gpurun_out/r2k_tests_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -1 gpurun_out/r2k_smoke.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_train.log 2>&1; tail -c 300 gpurun_out/r2k_train.log
