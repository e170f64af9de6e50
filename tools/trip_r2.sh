mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests_all.log 2>&1; tail -3 gpurun_out/r2r_tests_all.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2r_train.log 2>&1; python - <<'PY'
import json
for line in open("gpurun_out/r2r_train.log"):
    if line.startswith('{"metric'):
        d=json.loads(line); print('train', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'gemm', round(d['roofline']['ms_per_step'],3), round(d['roofline']['frac'],3), 'conv0', round(d['roofline_hbm']['ms_per_step'],3))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_reference_arm.log 2>&1; tail -c 600 gpurun_out/r2r_reference_arm.log
