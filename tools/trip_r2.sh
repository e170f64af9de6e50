mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests_all.log 2>&1; tail -4 gpurun_out/r2o_tests_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; tail -1 gpurun_out/r2o_smoke.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel|conv0_" -s 12 -c 12 -o gpurun_out/r2o_kernels python tools/ncu_kernels.py > gpurun_out/r2o_ncu_kernels.log 2>&1; tail -1 gpurun_out/r2o_ncu_kernels.log
# shared-memory race check of the persistent attention kernels and the small-shape kernel tests
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_attention_persist.py -m gpu -x -q -k "20-96 or 160-17 or 31-129" > gpurun_out/r2o_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2o_sanitizer_racecheck.log
