mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_engine.py -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; tail -3 gpurun_out/r2q_tests.log
b() { name=$1; shift; env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2q_$name.log 2>&1; python - "$name" <<'PY'
import json,sys
for line in open(f"gpurun_out/r2q_{sys.argv[1]}.log"):
    if line.startswith('{"metric'):
        d=json.loads(line); print(sys.argv[1], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'gemm', round(d['roofline']['ms_per_step'],3), 'conv0', round(d['roofline_hbm']['ms_per_step'],3))
PY
}
b default X=1
b plan_early W2V2_PLAN_EARLY=1
b no_tail_skip W2V2_GEMM_TAIL_SKIP=0
b default2 X=1
b both_old W2V2_PLAN_EARLY=1 W2V2_GEMM_TAIL_SKIP=0
